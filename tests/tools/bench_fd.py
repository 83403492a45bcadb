#!/usr/bin/env python
"""BASELINE config 2: full tfQMR solve of the reference's FD_problem.xml (complex fp64, 171 block rows of 8x8, 1 RHS block
column) through the C-ABI: solve time and iterations, next to the reference's own CUDA build on the same GPU (oracle/_ref)
and the reference CPU build.  Launch-latency-bound: the roofline fraction is not meaningful here."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orclib as O
from tfqmrgpu_b200 import api, problems as P

prob = P.read_xml(os.path.join(ROOT, "tests", "golden", "FD_problem.xml"))
vA = P.interleave(prob.A.val, np.float64); vB = P.interleave(prob.B.val, np.float64)
h = api.Handle()
pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
pl.buffer_size_for(prob.lm, prob.ln, "z"); pl.set_buffer()
pl.set_matrix("A", vA, "t"); pl.set_matrix("B", vB, "t")
ts = []
for _ in range(20):
    t0 = time.perf_counter(); st = pl.solve(prob.tolerance, 2000); ts.append(time.perf_counter() - t0)
info = pl.info()
out = {"config": "FD_problem.xml, complex fp64, tol %g" % prob.tolerance,
       "ours": {"status": int(st), "iterations": info["iterations"], "residual": info["residuum"],
                "solve_ms_median": 1e3*float(np.median(ts[5:])), "solve_ms_min": 1e3*min(ts)}}
pl.close(); h.close()
for name, ref in (("reference_gpu", O.ref_gpu()), ("reference_cpu", O.ref_cpu())):
    if ref is None:
        continue
    tr = []
    for _ in range(5):
        with O.quiet_stdout():
            r = ref.solve(prob.mb, prob.lm, prob.ln, prob.A.rowptr, prob.A.colind, vA, prob.X.rowptr, prob.X.colind,
                          prob.B.rowptr, prob.B.colind, vB, prob.tolerance, 2000, "z", transA="t", trans_b="t")
        tr.append(r["t_solve"])
    out[name] = {"status": int(r["status"]), "iterations": r["iterations"], "residual": r["residuum"],
                 "solve_ms_median": 1e3*float(np.median(tr[1:])), "solve_ms_min": 1e3*min(tr)}
print(json.dumps(out))
