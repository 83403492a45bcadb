# dev-only: accuracy of the fp16-pair tensor-core product against fp64 as a function of the accumulation segment (TFQMRGPU_TC_CHAIN),
# cos/sin fill (the reference harness's operands, pass bar 1e-4 absolute, bench_tfqmrgpu.cu:414)
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import orclib as O
from tfqmrgpu_b200 import api, problems as P, _lib as L
import test_gpu_parity as T
rp, ci = P.stencil27_pattern(4)
z = np.load(os.path.join('tests', 'golden', 'plan_unordered.npz'))
nY, nA, nX = [int(v) for v in z['nnz']]
pu = P.bsr_from_multiplication_plan(z['starts'], z['pairs'], nA)
cases = []
for ncol in (1, 2):
    rpX = (ncol*np.arange(65)).astype(np.int32); ciX = np.tile(np.arange(ncol, dtype=np.int32), 64)
    cases.append((f'stencil27 32x32 ncol {ncol}', 64, rp, ci, rpX, ciX, 32, 32, None))
    cases.append((f'stencil27 64x64 ncol {ncol}', 64, rp, ci, rpX, ciX, 64, 64, None))
for lm, ln in ((32, 32), (64, 64), (16, 16)):
    cases.append((f'plan_unordered {lm}x{ln}', pu[0], pu[1], pu[2], pu[3], pu[4], lm, ln, nA))
prob = P.random_system(64, 32, 32, ncols=3, pA=1.0, pX=1.0, seed=64, unsorted=True)
cases.append(('64 entries/row 32x32', prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, 32, 32, None))
for name, mb, rpA, ciA, rpX, ciX, lm, ln, na in cases:
    A, X, Y, lists = T._spmm_case(mb, rpA, ciA, rpX, ciX, lm, ln, "c", na)
    Y64 = O.multiply(A.astype(np.float64), X.astype(np.float64), lists["starts"], lists["pairs"], lm, ln, nthreads=8)
    print(f"chain {os.environ.get('TFQMRGPU_TC_CHAIN', 'default')} {name}: max|Y| {np.abs(Y64).max():.2f}  err {np.abs(Y-Y64).max():.3e}", flush=True)
