# dev-only: the fp16-pair tensor-core product (spmm_tc16.cu) against the fp64 oracle on small problems (run on the GPU box)
#   python tests/tools/dev_tc16_check.py [quick]
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import orclib as O
from tfqmrgpu_b200 import problems as P, api, _lib as L


def run(mb, rpA, ciA, rpX, ciX, lm, ln, label, fill='cos', nA=None):
    h = api.Handle()
    pl = api.BsrsvPlan(h, mb, rpA, ciA, rpX, ciX, rpX, ciX)
    pl.buffer_size_for(lm, ln, 'c'); pl.set_buffer()
    info = pl.plan_info()
    nA, nX = nA or len(ciA), len(ciX)
    if fill == 'cos':
        A = O.fill_cos_sin(nA, lm, lm, np.float32); X = O.fill_cos_sin(nX, lm, ln, np.float32)
    else:
        rng = np.random.default_rng(1)
        A = rng.uniform(-1, 1, (nA, 2, lm, lm)).astype(np.float32); X = rng.uniform(-1, 1, (nX, 2, lm, ln)).astype(np.float32)
        if fill == 'wide':    # right-hand-side columns and block rows of very different magnitude
            X *= (10.0**rng.uniform(-12, 6, (1, 1, 1, ln))).astype(np.float32)
            A *= (10.0**rng.uniform(-3, 3, (nA, 1, 1, 1))).astype(np.float32)
    pl.set_matrix('A', A, 't', L.LAYOUT_RRRRIIII); pl.set_matrix('X', X, 'n', L.LAYOUT_RRRRIIII)
    pl.multiply(1)
    Y = pl.get_vector('Y', 'n', L.LAYOUT_RRRRIIII).reshape(nX, 2, lm, ln)
    lists = pl.plan_lists()
    Y64 = O.multiply(A.astype(np.float64), X.astype(np.float64), lists['starts'], lists['pairs'], lm, ln, nthreads=8)
    Y32 = O.multiply(A, X, lists['starts'], lists['pairs'], lm, ln, nthreads=8)
    scale = np.abs(Y64).max(axis=(0, 1, 2), keepdims=True) + 1e-300     # per right-hand-side lane j
    err = np.abs(Y - Y64); e32 = np.abs(Y32 - Y64)
    npairs = np.diff(lists['starts'].astype(np.int64))
    print(f'{label}: tc {info["use_tc"]} gmax {info["gmax"]} units {info["nUnits"]} entries {info["nEntries"]} pairs/row {npairs.min()}..{npairs.max()}  '
          f'max|Y| {np.abs(Y64).max():.3e}  err max {err.max():.3e} (rel/col {(err/scale).max():.2e}) rms {np.sqrt((err**2).mean()):.3e}   '
          f'fp32-order err max {e32.max():.3e} (rel/col {(e32/scale).max():.2e})', flush=True)
    if (err/scale).max() > 1e-4:
        b, c, i, j = np.unravel_index(np.argmax(err/scale), err.shape)
        print('   worst at block', b, 'plane', c, 'i', i, 'j', j, 'got', Y[b, c, i, j], 'want', Y64[b, c, i, j])
        print('   per-plane max err', err.max(axis=(0, 2, 3)), ' per-i', np.round(err.max(axis=(0, 1, 3))[:8], 4), ' per-j', np.round(err.max(axis=(0, 1, 2))[:8], 4))
        print('   Y[0,0,:4,:4]\n', Y[0, 0, :4, :4], '\n   want\n', Y64[0, 0, :4, :4])
        print('   ratio Y/Y64 [0,0,:4,:4]\n', Y[0, 0, :4, :4]/Y64[0, 0, :4, :4])
    pl.close(); h.close()
    return float(err.max())


quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
shapes = [(32, 32)] if quick else [(32, 32), (16, 16), (16, 32), (16, 64), (32, 64), (64, 64)]
for (lm, ln) in shapes:
    prob = P.random_system(12, lm, ln, seed=lm*100 + ln, unsorted=True)
    run(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, f'rand {lm}x{ln} cos')
    run(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, f'rand {lm}x{ln} uniform', fill='u')
    run(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, f'rand {lm}x{ln} wide', fill='wide')
if quick:
    sys.exit(0)
rp, ci = P.stencil27_pattern(4)
for (lm, ln) in [(32, 32), (64, 64), (16, 16)]:
    for ncol in (1, 2, 3, 5):
        rpX = (ncol*np.arange(65)).astype(np.int32); ciX = np.tile(np.arange(ncol, dtype=np.int32), 64)
        run(64, rp, ci, rpX, ciX, lm, ln, f'stencil4 {lm}x{ln} ncol {ncol} cos')
# long rows: 64 entries per row (several accumulation segments)
for lm, ln in ((16, 16), (32, 32), (32, 64), (64, 64)):
    prob = P.random_system(64, lm, ln, ncols=3, pA=1.0, pX=1.0, seed=lm + ln, unsorted=True)
    run(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, lm, ln, f'64 entries/row {lm}x{ln} cos')
# the reference's own multiplication plans at the harness's block sizes (pass bar 1e-4, bench_tfqmrgpu.cu:414)
for name in ('plan_unordered', 'plan_reordered'):
    z = np.load(os.path.join('tests', 'golden', name + '.npz'))
    starts, pairs = z['starts'], z['pairs']
    nY, nA, nX = [int(v) for v in z['nnz']]
    if name == 'plan_unordered':
        mb, rpA, ciA, rpX, ciX = P.bsr_from_multiplication_plan(starts, pairs, nA)
        for lm, ln in ((16, 16), (32, 32), (32, 64), (64, 64)):
            e = run(mb, rpA, ciA, rpX, ciX, lm, ln, f'{name} {lm}x{ln} cos', nA=nA)
            print('   ', 'PASS' if e <= 1e-4 else 'FAIL', 'reference bar 1e-4', flush=True)
