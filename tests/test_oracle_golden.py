"""CPU tests: the oracle (plain-C restatement) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py) and against the reference's own fixtures."""
import os

import numpy as np
import pytest

import orclib as O
from cases import golden_cases, CFG1_FINGERPRINT
from tfqmrgpu_b200 import problems as P

CASES = golden_cases()


def _operands(prob, prec, tA, tB):
    dt = np.float64 if prec == "z" else np.float32
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
    A_int = O.import_blocks(vA, prob.A.nnzb, prob.lm, prob.lm, trans=tA, var="A")
    B_int = O.import_blocks(vB, prob.B.nnzb, prob.lm, prob.ln, trans=tB, var="B")
    return A_int, B_int


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_plan_lists_bit_exact_vs_reference(case, golden):
    name, prob = case[0], case[1]
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    assert pl.status == 0
    for k in ("starts", "pairs", "subset", "colindx"):
        assert np.array_equal(getattr(pl, k), golden[f"{name}_{k}"]), k


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_solve_bit_exact_vs_reference_cpu_build(case, golden):
    """ORC_MODE_CPUREF restates the reference's HAS_NO_CUDA arithmetic: identical bits expected."""
    name, prob, prec, tol, maxit, tA, tB = case
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    A_int, B_int = _operands(prob, prec, tA, tB)
    o = O.solve(pl, prob.lm, prob.ln, A_int, B_int, golden[f"{name}_v3"], tol, maxit, mode=O.MODE_CPUREF)
    status, iters, res, flops, bufsize = golden[f"{name}_scalars"]
    assert o["status"] == int(status)
    assert o["iterations"] == int(iters)
    assert o["residuum"] == res
    assert o["flops"] == flops
    assert pl.ref_buffer_size(prob.lm, prob.ln, prec == "z") == int(bufsize)
    assert np.array_equal(o["X"], golden[f"{name}_X"])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_arithmetic_mode_close_to_reference(case, golden):
    """The GPU-path flavour (float products in dotp, one accumulator in the block product) must agree
    with the CPU-path flavour to rounding: same iterations +-1, X within 10*tol*max|X| per column."""
    name, prob, prec, tol, maxit, tA, tB = case
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    A_int, B_int = _operands(prob, prec, tA, tB)
    o = O.solve(pl, prob.lm, prob.ln, A_int, B_int, golden[f"{name}_v3"], tol, maxit, mode=O.MODE_GPU)
    assert o["status"] == 0
    assert abs(o["iterations"] - int(golden[f"{name}_scalars"][1])) <= 1
    Xr = golden[f"{name}_X"]
    assert np.abs(o["X"] - Xr).max() <= 10*tol*np.abs(Xr).max()


def test_julia_known_answer():
    """example/tfqmrgpu_Julia_example.jl:117-120: X_k = k/8 * B exactly."""
    prob = P.julia_kat()
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    A_int, B_int = _operands(prob, "z", "n", "n")
    v3 = O.v3_glibc(pl.nnzbX*2*prob.lm*prob.ln)
    o = O.solve(pl, prob.lm, prob.ln, A_int, B_int, v3, 1.2e-8, 210)
    Xc = o["X"][:, 0] + 1j*o["X"][:, 1]
    assert o["status"] == 0 and o["iterations"] == 7 and o["flops"] == 285440.0
    assert np.abs(Xc - prob.X_exact).max() < 1e-13


def test_fd_problem_golden_numbers(golden):
    """FD_problem.xml z -> 42 iterations, 596 397 312 flop, 2 974 208 B workspace."""
    prob = P.read_xml(os.path.join(os.path.dirname(__file__), "golden", "FD_problem.xml"))
    assert (prob.mb, prob.lm, prob.ln, prob.A.nnzb, prob.X.nnzb, prob.B.nnzb) == (171, 8, 8, 1557, 171, 1)
    pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    assert pl.nPairs == 1557 and pl.nCols == 1
    assert list(pl.pairs[:16]) == [0, 0, 1, 1, 2, 7, 3, 37, 4, 6, 5, 30, 6, 134, 7, 2]   # SURVEY 8c fingerprint
    assert list(pl.starts[:6]) == [0, 13, 26, 38, 45, 52] and pl.starts[-1] == 1557
    assert pl.ref_buffer_size(8, 8, True) == 2974208
    A_int, B_int = _operands(prob, "z", "t", "t")
    v3 = golden["fd_z_v3"]  # the glibc rand() stream the reference drew for this fixture
    o = O.solve(pl, 8, 8, A_int, B_int, v3, prob.tolerance, 2000, mode=O.MODE_CPUREF)
    assert (o["status"], o["iterations"], o["probes"], o["flops"]) == (0, 42, 2, 596397312.0)
    assert o["residuum"] < 1e-9
    # true residual of the returned X
    Xc = o["X"][:, 0] + 1j*o["X"][:, 1]
    Ac = np.transpose(prob.A.val, (0, 2, 1))      # file blocks are column-major, uploaded with 't'
    R = np.zeros_like(Xc)
    rows = np.repeat(np.arange(prob.mb), np.diff(prob.A.rowptr))
    for a in range(prob.A.nnzb):
        R[rows[a]] += Ac[a] @ Xc[prob.A.colind[a]]
    Bc = np.transpose(prob.B.val, (0, 2, 1))
    brow = np.repeat(np.arange(prob.mb), np.diff(prob.B.rowptr))
    R[brow[0]] -= Bc[0]
    assert np.sqrt((np.abs(R)**2).sum(axis=(0, 1)).max()) < 1e-9


def test_plan_file_is_a_createplan_dump(plan_unordered):
    """Reconstruct BSR patterns from the reference's plan file and rebuild it with createPlan."""
    starts, pairs = plan_unordered["starts"], plan_unordered["pairs"]
    nY, nA, nX = plan_unordered["nnz"]
    mb, rpA, ciA, rpX, ciX = P.bsr_from_multiplication_plan(starts, pairs, int(nA))
    assert mb == 1063 and ciX.max() + 1 == 16
    pl = O.OraclePlan(mb, rpA, ciA, rpX, ciX, rpX, ciX)
    assert pl.status == 0
    assert np.array_equal(pl.starts, starts)
    assert np.array_equal(pl.pairs.reshape(-1, 2), pairs)


def test_config1_spmm_fingerprint(plan_unordered):
    starts, pairs = plan_unordered["starts"], plan_unordered["pairs"]
    nY, nA, nX = [int(v) for v in plan_unordered["nnz"]]
    A = O.fill_cos_sin(nA, 16, 16, np.float32).astype(np.float64)
    X = O.fill_cos_sin(nX, 16, 16, np.float32).astype(np.float64)
    Y = O.multiply(A, X, starts, pairs, 16, 16, nthreads=8)
    f = CFG1_FINGERPRINT
    assert abs(Y[:, 0].sum() - f["sum_re"]) < 1e-5 and abs(Y[:, 1].sum() - f["sum_im"]) < 1e-5
    assert abs((Y**2).sum() - f["sum_abs2"]) < 1.
    assert abs(np.sqrt(Y[:, 0]**2 + Y[:, 1]**2).max() - f["max_abs"]) < 1e-5
    assert abs(Y[0, 0, 0, 0] - f["y000"][0]) < 1e-8 and abs(Y[0, 1, 0, 0] - f["y000"][1]) < 1e-8
    assert abs(Y[0, 0, 3, 5] - f["y035"][0]) < 1e-8 and abs(Y[-1, 1, 15, 15] - f["ylast"][1]) < 1e-8
    # fp32 accumulation stays within the reference's pass bar of 1e-4 (bench_tfqmrgpu.cu:414)
    Yf = O.multiply(A.astype(np.float32), X.astype(np.float32), starts, pairs, 16, 16, nthreads=8)
    assert np.abs(Yf - Y).max() < 1e-4


def test_reordered_plan_same_result(plan_unordered, plan_reordered):
    """plan_reordered lists the same Y blocks in another order: identical blocks expected."""
    nY, nA, nX = [int(v) for v in plan_unordered["nnz"]]
    A = O.fill_cos_sin(nA, 16, 16, np.float32)
    X = O.fill_cos_sin(nX, 16, 16, np.float32)
    Yu = O.multiply(A, X, plan_unordered["starts"], plan_unordered["pairs"], 16, 16, nthreads=8)
    Yr = O.multiply(A, X, plan_reordered["starts"], plan_reordered["pairs"], 16, 16, nthreads=8)
    ou, orr = plan_unordered["yorder"], plan_reordered["yorder"]
    assert sorted(ou.tolist()) == sorted(orr.tolist())
    lut = {int(y): i for i, y in enumerate(ou)}
    idx = np.array([lut[int(y)] for y in orr])
    assert np.array_equal(Yr, Yu[idx])


@pytest.mark.parametrize("layout", [O.LAYOUT_RIRIRIRI, O.LAYOUT_RRIIRRII, O.LAYOUT_RRRRIIII])
@pytest.mark.parametrize("trans", ["n", "t", "c", "*"])
def test_layout_roundtrip_and_meaning(layout, trans):
    rng = np.random.default_rng(3)
    rows, cols, nb = 4, 5, 3
    Z = rng.normal(size=(nb, rows, cols)) + 1j*rng.normal(size=(nb, rows, cols))
    # build the host array for this layout / transposition by definition
    S = {"n": Z, "t": np.transpose(Z, (0, 2, 1)), "c": np.conj(np.transpose(Z, (0, 2, 1))), "*": np.conj(Z)}[trans]
    if layout == O.LAYOUT_RIRIRIRI:
        host = np.stack([S.real, S.imag], axis=-1)
    elif layout == O.LAYOUT_RRIIRRII:
        host = np.stack([S.real, S.imag], axis=2)
    else:
        host = np.stack([S.real, S.imag], axis=1)
    host = np.ascontiguousarray(host).reshape(-1)
    internal = O.import_blocks(host, nb, rows, cols, layout, trans, "X")
    assert np.allclose(internal[:, 0] + 1j*internal[:, 1], Z)
    back = O.export_blocks(internal, rows, cols, layout, trans)
    assert np.array_equal(back, host)
    # 'A' is stored transposed internally (tfqmrgpu.cu:509-520)
    Zs = Z[:, :, :4]
    Ss = {"n": Zs, "t": np.transpose(Zs, (0, 2, 1)), "c": np.conj(np.transpose(Zs, (0, 2, 1))), "*": np.conj(Zs)}[trans]
    hostA = np.ascontiguousarray(np.stack([Ss.real, Ss.imag], axis=-1)).reshape(-1)
    intA = O.import_blocks(hostA, nb, 4, 4, O.LAYOUT_RIRIRIRI, trans, "A")
    assert np.allclose(intA[:, 0] + 1j*intA[:, 1], np.transpose(Zs, (0, 2, 1)))


def test_createplan_error_codes():
    """tfqmrgpu.cu:166-172,245,335: payload conventions."""
    k = P.julia_kat()
    a = (k.A.rowptr, k.A.colind, k.X.rowptr, k.X.colind)
    # B block in a column X does not have
    pl = O.OraclePlan(k.mb, *a, k.B.rowptr, np.array([3], np.int32))
    assert pl.status == 13 + 1000*6
    # X has two columns but B only covers one -> 1 zero column
    rpX = np.arange(0, 2*k.mb + 1, 2, dtype=np.int32); ciX = np.tile(np.array([0, 1], np.int32), k.mb)
    pl = O.OraclePlan(k.mb, k.A.rowptr, k.A.colind, rpX, ciX, k.B.rowptr, k.B.colind)
    assert pl.status == 11 + 1000*1
    # Fortran offsets give the same lists
    p0 = O.OraclePlan(k.mb, *a, k.B.rowptr, k.B.colind)
    p1 = O.OraclePlan(k.mb, k.A.rowptr + 1, k.A.colind + 1, k.X.rowptr + 1, k.X.colind + 1, k.B.rowptr + 1, k.B.colind + 1, 1)
    assert p0.status == 0 and p1.status == 0
    for key in ("starts", "pairs", "subset", "colindx"):
        assert np.array_equal(getattr(p0, key), getattr(p1, key))
    assert O.OraclePlan(0, *a, k.B.rowptr, k.B.colind).status == 14


@pytest.mark.skipif(O.ref_cpu() is None, reason="oracle/_ref not built (reference not present)")
def test_oracle_vs_live_reference_random_systems():
    """Where the reference itself is available: fresh seeded systems, bit-exact again."""
    ref = O.ref_cpu()
    for seed, (lm, ln), prec, tol in [(21, (4, 4), "z", 1e-9), (22, (8, 10), "z", 1e-9), (23, (16, 16), "c", 1e-4)]:
        prob = P.random_system(8, lm, ln, seed=seed, unsorted=True)
        dt = np.float64 if prec == "z" else np.float32
        vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
        r = ref.solve(prob.mb, lm, ln, prob.A.rowptr, prob.A.colind, vA, prob.X.rowptr, prob.X.colind,
                      prob.B.rowptr, prob.B.colind, vB, tol, 200, prec)
        pl = O.OraclePlan(prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
        A_int, B_int = _operands(prob, prec, "n", "n")
        o = O.solve(pl, lm, ln, A_int, B_int, r["v3"], tol, 200, mode=O.MODE_CPUREF)
        assert (o["status"], o["iterations"], o["flops"]) == (r["status"], r["iterations"], r["flops"])
        assert np.array_equal(o["X"], r["X"])
