#!/usr/bin/env python
"""BASELINE config 5 (reduced): every allowed block size x {complex fp32, complex fp64} on the 27-point block stencil
(n^3 block rows, 64 RHS columns), one GPU: iterations, time per iteration, which product kernel ran, and the product's
GFLOP/s / GB/s (formulas of SURVEY.md 8d).  fp32 at tol 1e-3 (sigma 8), fp64 at tol 1e-9 (sigma 1)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tfqmrgpu_b200 import api, synthetic, _lib as L

ALLOWED = [(4, 4), (4, 5), (4, 8), (4, 32), (8, 8), (8, 9), (8, 10), (8, 32), (8, 64), (16, 16), (16, 32), (16, 64), (32, 32), (32, 64), (64, 64)]
n = int(os.environ.get("N", "12"))
max_lm = int(os.environ.get("MAX_LM", "64"))      # e.g. MAX_LM=8: only the small-block kernel's sizes
rows = []
for lm, ln in ALLOWED:
    if lm > max_lm:
        continue
    ncols = max(1, 64//ln)
    for prec, dt, sigma, tol in (("c", np.float32, 8.0, 1e-3), ("z", np.float64, 1.0, 1e-9)):
        es = 4 if prec == "c" else 8
        sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=sigma, dtype=dt, device="cuda")
        h = api.Handle(); pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
        pl.buffer_size_for(lm, ln, prec); pl.set_buffer()
        pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix("B", sp.valB)
        info = pl.plan_info()
        for _ in range(2):
            st = pl.solve(tol, 200)
        pl.set_profiling(True)
        st = pl.solve(tol, 200); prof = pl.solve_profile(); res = pl.info()
        pl.set_profiling(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pl.solve(tol, 200); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        nP = info["nPairs"]
        sp_ms = prof["spmm_ms"]/max(prof["spmm_launches"], 1)
        flops = nP*8*lm*lm*ln
        nbytes = sp.nnzbA*2*lm*lm*es + 2*info["nnzbX"]*2*lm*ln*es + 8*nP + 4*(info["nnzbX"] + 1)
        rows.append(dict(lm=lm, ln=ln, prec=prec, kernel="dmma" if info["use_dmma"] else ("tcgen05" if info["use_tc"] else ("simt-small" if info.get("use_small") else "simt")),
                         status=int(st), iterations=res["iterations"], residual=res["residuum"], solve_ms=ms,
                         ms_per_iteration=ms/max(res["iterations"], 1), spmm_us=1e3*sp_ms,
                         spmm_gflops=flops/sp_ms*1e-6 if sp_ms else 0, spmm_gbs=nbytes/sp_ms*1e-6 if sp_ms else 0))
        pl.close(); h.close(); del sp
        torch.cuda.empty_cache()
print(json.dumps({"n": n, "block_rows": n**3, "rhs_columns": 64, "rows": rows}))
