import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import orclib as O
from tfqmrgpu_b200 import problems as P, api, _lib as L
lm = ln = 32
# ONE block row, one A block (diagonal), two block columns -> one unit with a single entry
mb = 1; ncol = 2
rp = np.array([0, 1], np.int32); ci = np.array([0], np.int32)
rpX = np.array([0, ncol], np.int32); ciX = np.arange(ncol, dtype=np.int32)
h = api.Handle()
pl = api.BsrsvPlan(h, mb, rp, ci, rpX, ciX, rpX, ciX)
pl.buffer_size_for(lm, ln, 'c'); pl.set_buffer()
nA, nX = 1, ncol
ca, k, i = np.meshgrid(np.arange(2), np.arange(lm), np.arange(lm), indexing='ij')
A = (ca*1000 + k*32 + i).astype(np.float32)[None]          # A[ca][k][i] internal
b, c, k2, j = np.meshgrid(np.arange(nX), np.arange(2), np.arange(lm), np.arange(ln), indexing='ij')
X = (b*4000 + c*2000 + k2*32 + j).astype(np.float32)        # X[b][c][k][j]
pl.set_matrix('A', A, 't', L.LAYOUT_RRRRIIII); pl.set_matrix('X', X, 'n', L.LAYOUT_RRRRIIII)
pl.multiply(1)
Y = pl.get_vector('Y', 'n', L.LAYOUT_RRRRIIII).reshape(nX, 2, lm, ln)
np.set_printoptions(linewidth=250, suppress=True)
dbg = int(os.environ.get('TFQMRGPU_TC_DEBUG', '0'))
print('debug', dbg)
print('Y[0] plane0 [i=:6, j=:8]\n', Y[0, 0, :6, :8]); print('Y[0] plane1 [i=:6, j=:8]\n', Y[0, 1, :6, :8])
if dbg & 16:   # X == 1: D[m][(ca,i)] = sum_k A[ca][k][i]
    print('expect plane0 row i:', A[0, 0].sum(axis=0)[:6], ' plane1:', A[0, 1].sum(axis=0)[:6])
if dbg & 32:   # A == 1: D[(g,cx,j)][n] = sum_k X[g][cx][k][j]
    cx = (dbg >> 7) & 1
    print('expect (all i) over j:', X[0, cx].sum(axis=0)[:8])
