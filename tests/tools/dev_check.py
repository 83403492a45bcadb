# dev-only quick GPU check (not a test): plan parity, SpMM parity, solve parity on small problems
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import orclib as O
from tfqmrgpu_b200 import problems as P, api, _lib as L

def check_problem(prob, prec, tol, maxit, transA='n', transB='n', label=''):
    dt = np.float64 if prec == 'z' else np.float32
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
    h = api.Handle()
    pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    op = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    lists = pl.plan_lists()
    for k in ('starts', 'pairs', 'subset', 'colindx'):
        assert np.array_equal(lists[k], getattr(op, k)), (label, k)
    pl.buffer_size_for(prob.lm, prob.ln, prec)
    pl.set_buffer()
    v3 = pl.get_v3()
    pl.set_matrix('A', vA, transA); pl.set_matrix('B', vB, transB)
    # SpMM check: X := cos/sin fill
    Xf = O.fill_cos_sin(pl.nnzbX, prob.lm, prob.ln, dt)
    pl.set_matrix('X', Xf, 'n', L.LAYOUT_RRRRIIII)
    pl.multiply(1)
    Y = pl.get_vector('Y', 'n', L.LAYOUT_RRRRIIII).reshape(pl.nnzbX, 2, prob.lm, prob.ln)
    A_int = O.import_blocks(vA, prob.A.nnzb, prob.lm, prob.lm, trans=transA, var='A')
    B_int = O.import_blocks(vB, prob.B.nnzb, prob.lm, prob.ln, trans=transB, var='B')
    Yo = O.multiply(A_int, Xf, op.starts, op.pairs, prob.lm, prob.ln)
    print(f'{label}: spmm maxdev {np.abs(Y - Yo).max():.3e} (scale {np.abs(Yo).max():.3e})')
    Xback = pl.get_vector('X', 'n', L.LAYOUT_RRRRIIII).reshape(Xf.shape)
    assert np.array_equal(Xback, Xf), 'X roundtrip'
    t0 = time.time(); st = pl.solve(tol, maxit); t1 = time.time()
    info = pl.info(); stats = pl.solve_stats()
    X = pl.get_matrix('X', 'n', L.LAYOUT_RRRRIIII).reshape(pl.nnzbX, 2, prob.lm, prob.ln)
    o = O.solve(op, prob.lm, prob.ln, A_int, B_int, v3, tol, maxit)
    scale = np.abs(o['X']).max()
    print(f'   solve st {st} it {info["iterations"]} res {info["residuum"]:.3e} flops {info["flops"]:.6g} probes {stats["probes"]} host_ms {stats["host_ms"]:.2f}')
    print(f'   orc   st {o["status"]} it {o["iterations"]} res {o["residuum"]:.3e} flops {o["flops"]:.6g} probes {o["probes"]}  max|dX|/scale {np.abs(X - o["X"]).max()/scale:.3e}')
    pl.close(); h.close()

k = P.julia_kat()
check_problem(k, 'z', 1.2e-8, 210, label='julia z')
check_problem(k, 'c', 1.2e-5, 210, label='julia c')
fd = P.read_xml('tests/golden/FD_problem.xml')
check_problem(fd, 'z', fd.tolerance, 2000, 't', 't', label='FD z')
for w in range(3):
    check_problem(P.fortran_pattern(w), 'z', 1e-9, 200, label=f'fortran{w} z')
for (lm, ln) in [(4,4),(4,8),(4,32),(8,8),(8,9),(8,10),(8,32),(8,64),(16,16),(16,32),(16,64),(32,32),(32,64),(64,64)]:
    for prec, tol in (('z', 1e-9), ('c', 1e-4)):
        check_problem(P.random_system(12, lm, ln, seed=lm*100+ln, unsorted=True), prec, tol, 200, label=f'rand {lm}x{ln} {prec}')
