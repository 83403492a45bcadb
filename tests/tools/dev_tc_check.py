# dev-only: tensor-core SpMM vs oracle on small problems (run on the GPU box)
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import orclib as O
from tfqmrgpu_b200 import problems as P, api, _lib as L

def run(mb, rpA, ciA, rpX, ciX, lm, ln, label, fill='cos'):
    h = api.Handle()
    pl = api.BsrsvPlan(h, mb, rpA, ciA, rpX, ciX, rpX, ciX)
    pl.buffer_size_for(lm, ln, 'c'); pl.set_buffer()
    info = pl.plan_info()
    nA, nX = len(ciA), len(ciX)
    if fill == 'cos':
        A = O.fill_cos_sin(nA, lm, lm, np.float32); X = O.fill_cos_sin(nX, lm, ln, np.float32)
    else:
        rng = np.random.default_rng(1)
        A = rng.uniform(-1, 1, (nA, 2, lm, lm)).astype(np.float32); X = rng.uniform(-1, 1, (nX, 2, lm, ln)).astype(np.float32)
        if fill == 'tf32':   # operands exactly representable in TF32: isolates the tensor core's accumulation error
            A = (A.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32); X = (X.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    pl.set_matrix('A', A, 't', L.LAYOUT_RRRRIIII); pl.set_matrix('X', X, 'n', L.LAYOUT_RRRRIIII)
    pl.multiply(1)
    Y = pl.get_vector('Y', 'n', L.LAYOUT_RRRRIIII).reshape(nX, 2, lm, ln)
    lists = pl.plan_lists()
    Y64 = O.multiply(A.astype(np.float64), X.astype(np.float64), lists['starts'], lists['pairs'], lm, ln, nthreads=8)
    Y32 = O.multiply(A, X, lists['starts'], lists['pairs'], lm, ln, nthreads=8)
    err = np.abs(Y - Y64); e32 = np.abs(Y32 - Y64)
    print(f'{label}: use_tc {info["use_tc"]} gmax {info["gmax"]} units {info["nUnits"]}  max|Y| {np.abs(Y64).max():.3e}  '
          f'tc err max {err.max():.3e} rms {np.sqrt((err**2).mean()):.3e}   fp32 simt-order err max {e32.max():.3e} rms {np.sqrt((e32**2).mean()):.3e}')
    if err.max() > 1e-3*np.abs(Y64).max():
        b, c, i, j = np.unravel_index(np.argmax(err), err.shape)
        print('   worst at block', b, 'plane', c, 'i', i, 'j', j, 'got', Y[b, c, i, j], 'want', Y64[b, c, i, j])
        print('   per-plane max err', err.max(axis=(0, 2, 3)), ' per-i', np.round(err.max(axis=(0, 1, 3))[:8], 3), ' per-j', np.round(err.max(axis=(0, 1, 2))[:8], 3))
        print('   Y[0,0,:4,:4]\n', Y[0, 0, :4, :4], '\n   want\n', Y64[0, 0, :4, :4])
    pl.close(); h.close()

for (lm, ln) in [(32, 32), (32, 64)]:
    prob = P.random_system(12, lm, ln, seed=lm*100 + ln, unsorted=True)
    run(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, f'rand {lm}x{ln}')
    run(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, f'rand {lm}x{ln} uniform', fill='u')
rp, ci = P.stencil27_pattern(4)
for ncol in (1, 2, 3):
    rpX = (ncol*np.arange(65)).astype(np.int32); ciX = np.tile(np.arange(ncol, dtype=np.int32), 64)
    run(64, rp, ci, rpX, ciX, 32, 32, f'stencil4 32x32 ncol {ncol}', fill='u')
    run(64, rp, ci, rpX, ciX, 32, 32, f'stencil4 32x32 ncol {ncol} exact-tf32 operands', fill='tf32')
