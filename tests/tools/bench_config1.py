#!/usr/bin/env python
"""BASELINE config 1: the reference's own multiplication plan test/multiplication/plan_unordered.14-287-16 (golden copy in
tests/golden), complex fp32 16x16 blocks: time of the bare block-sparse product Y = A*X (`bench_tfqmrgpu multi` role),
warm (working set 45 MB < L2) and cold (L2 flushed between launches), GB/s and GFLOP/s with the formulas of SURVEY.md 8d,
next to the oracle's restatement of the reference's OpenMP CPU check loop (bench_tfqmrgpu.cu:358-404) on the host cores."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import orclib as O
from tfqmrgpu_b200 import api, problems as P, _lib as L

g = np.load(os.path.join(ROOT, "tests", "golden", "plan_unordered.npz"))
starts, pairs = g["starts"], g["pairs"]
nY, nA, nX = [int(v) for v in g["nnz"]]
out = {"config": "plan_unordered.14-287-16, complex 16x16 blocks, %d Y blocks, %d pairs" % (nY, len(pairs))}
mb, rpA, ciA, rpX, ciX = P.bsr_from_multiplication_plan(starts, pairs, nA)
for prec, dt in (("c", np.float32), ("z", np.float64)):
    es = np.dtype(dt).itemsize
    h = api.Handle(); pl = api.BsrsvPlan(h, mb, rpA, ciA, rpX, ciX, rpX, ciX)
    pl.buffer_size_for(16, 16, prec); pl.set_buffer()
    A = O.fill_cos_sin(nA, 16, 16, dt); X = O.fill_cos_sin(nX, 16, 16, dt)
    pl.set_matrix("A", A, "t", L.LAYOUT_RRRRIIII); pl.set_matrix("X", X, "n", L.LAYOUT_RRRRIIII)
    used_a = len(np.unique(pairs[:, 0]))
    nbytes = used_a*2*256*es + 2*nX*2*256*es + 8*len(pairs) + 4*(nY + 1)
    flops = len(pairs)*8*16**3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pl.multiply(20); torch.cuda.synchronize()
    e0.record(); pl.multiply(200); e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1)/200
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cold = []
    for _ in range(20):
        flush.zero_(); e0.record(); pl.multiply(1); e1.record(); torch.cuda.synchronize(); cold.append(e0.elapsed_time(e1))
    cold = float(np.median(cold))
    Y = pl.get_vector("Y", "n", L.LAYOUT_RRRRIIII).reshape(nX, 2, 16, 16)
    t0 = time.perf_counter(); Yo = O.multiply(A, X, starts, pairs.reshape(-1), 16, 16, nthreads=os.cpu_count()); tcpu = time.perf_counter() - t0
    out[prec] = {"warm_us": 1e3*warm, "cold_us": 1e3*cold, "algorithmic_MB": nbytes*1e-6, "GFLOP": flops*1e-9,
                 "warm_TFLOPs": flops/warm*1e-9, "cold_GBs": nbytes/cold*1e-6, "cold_TFLOPs": flops/cold*1e-9,
                 "maxdev_vs_oracle": float(np.abs(Y - Yo).max()), "cpu_check_loop_ms": 1e3*tcpu, "cpu_threads": os.cpu_count(),
                 "kernel": "dmma" if pl.plan_info()["use_dmma"] else ("tc" if pl.plan_info()["use_tc"] else "simt")}
    pl.close(); h.close()
print(json.dumps(out))
