# dev-only: accuracy of the tensor-core product against fp64 for rows of 27 ... 64 entries (pass boundaries at 896/LM entries), cos/sin fill
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import orclib as O
from tfqmrgpu_b200 import api, problems as P, _lib as L
import test_gpu_parity as T
for lm, ln in ((16, 16), (32, 32), (32, 64)):
    for mb in (27, 28, 29, 40, 57, 64):
        prob = P.random_system(mb, lm, ln, ncols=3, pA=1.0, pX=1.0, seed=lm + ln, unsorted=True)
        A, X, Y, lists = T._spmm_case(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, "c")
        Y64 = O.multiply(A.astype(np.float64), X.astype(np.float64), lists["starts"], lists["pairs"], lm, ln, nthreads=8)
        Y32 = O.multiply(A, X, lists["starts"], lists["pairs"], lm, ln, nthreads=8)
        print(f"{lm}x{ln} entries {mb}: max|Y| {np.abs(Y64).max():.2f}  tc err {np.abs(Y-Y64).max():.3e} ({np.abs(Y-Y64).max()/np.abs(Y64).max():.2e} rel)  simt-order fp32 err {np.abs(Y32-Y64).max():.3e}", flush=True)
