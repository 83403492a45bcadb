"""Drop-in evidence on the GPU box: the reference's OWN callers, compiled unmodified from /root/reference in the
build container (oracle/build_ref.sh callers|gpu; the binaries travel in oracle/_ref/, nothing here reads
/root/reference) and linked against OUR libtfQMRgpu.so, plus solve parity against the reference's own CUDA
kernels (oracle/_ref/libtfqmr_ref_gpu.so, nvcc -arch=sm_100 of the unmodified sources) on the same GPU with the
same cuRAND shadow vector."""
import os
import re
import subprocess

import numpy as np
import pytest

import orclib as O
from cases import golden_cases
from tfqmrgpu_b200 import api, problems as P, _lib as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _have(name):
    return os.path.exists(os.path.join(REFDIR, name))


@pytest.mark.skipif(not _have("c_example_ours"), reason="oracle/_ref/c_example_ours not built")
def test_reference_c_example_runs_against_our_library():
    """example/tfqmrgpu_C_example.c (args of SURVEY 8c) through tfqmrgpu_bsrsv_z of OUR library."""
    out = subprocess.run([os.path.join(REFDIR, "c_example_ours"), "6", "4", "4", ".125", ".5", ".125", "100", "1e-6"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"tfQMRgpu converged to ([0-9.e+-]+) in (\d+) iterations", out.stdout)
    assert m, out.stdout
    assert float(m.group(1)) <= 1e-6 and 3 <= int(m.group(2)) <= 40


@pytest.mark.skipif(not (_have("bench_tfqmrgpu_ours") and _have("bench_tfqmrgpu_ref")), reason="reference bench not built")
def test_reference_bench_harness_same_output_with_both_libraries():
    """`bench_tfqmrgpu tfQMR FD_problem.xml z` (README :61-63): the reference's harness linked against our library and
    against the reference's own CUDA build must report the same solution statistics (the file's reference X is all
    zeros, so maxdev/avgdev are max|X| and mean|X|)."""
    xml = os.path.join(ROOT, "tests", "golden", "FD_problem.xml")
    vals = {}
    for which in ("ours", "ref"):
        out = subprocess.run([os.path.join(REFDIR, "bench_tfqmrgpu_" + which), "tfQMR", xml, "z"],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        m = re.search(r"# GPU maxdev ([0-9.e+-]+) avgdev ([0-9.e+-]+)", out.stdout)
        assert m, out.stdout[-2000:]
        vals[which] = (float(m.group(1)), float(m.group(2)))
    assert abs(vals["ours"][0] - vals["ref"][0]) <= 1e-8*vals["ref"][0]
    assert abs(vals["ours"][1] - vals["ref"][1]) <= 1e-8*vals["ref"][1]


@pytest.mark.skipif(not _have("bench_tfqmrgpu_ours"), reason="reference bench not built")
def test_reference_bench_reads_files_we_wrote(tmp_path):
    """SURVEY 8f item 2 end to end on the GPU: the reference's harness (its XML reader and its legacy text reader) + our library
    solve the original FD_problem.xml, our re-written XML (no indirection, scale folded in) and our legacy dump alike."""
    from tfqmrgpu_b200 import formats as F
    gold = os.path.join(ROOT, "tests", "golden", "FD_problem.xml")
    prob = P.read_xml(gold)
    ours_xml, ours_txt = str(tmp_path / "ours.xml"), str(tmp_path / "ours_problem.txt")
    F.write_xml(ours_xml, F.xml_from_problem(prob), lossless=True)
    F.write_legacy(ours_txt, prob)
    vals = []
    for f in (gold, ours_xml, ours_txt):
        out = subprocess.run([os.path.join(REFDIR, "bench_tfqmrgpu_ours"), "tfQMR", f, "z"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        m = re.search(r"# GPU maxdev ([0-9.e+-]+) avgdev ([0-9.e+-]+)", out.stdout)
        assert m, out.stdout[-2000:]
        vals.append((float(m.group(1)), float(m.group(2))))
    assert vals[1] == vals[0] and vals[2] == vals[0]


def _cli(*args):
    import sys
    out = subprocess.run([sys.executable, "-m", "tfqmrgpu_b200.bench_cli", *args], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_bench_cli_multiply_on_the_reference_plan(tmp_path):
    """`bench_tfqmrgpu multiply <plan> f|d reps samples lm ln` conventions and result lines, on this library's
    product kernels, for the reference's plan_unordered.14-287-16 (written back to its text format from the golden arrays)."""
    from tfqmrgpu_b200 import formats as F
    g = np.load(os.path.join(ROOT, "tests", "golden", "plan_unordered.npz"))
    plan = str(tmp_path / "plan_unordered.14-287-16")
    F.write_multiplication_plan(plan, g["starts"], g["pairs"], int(g["nnz"][1]), int(g["nnz"][2]))
    for fF, bar in (("f", 1e-4), ("d", 1e-10)):
        txt = _cli("multiply", plan, fF, "20", "3", "16", "16")
        assert "# found 4490 result elements" in txt and "# found 50526 operations" in txt and "Warning" not in txt
        m = re.search(r"# GPU maxdev ([0-9.e+-]+) avgdev ([0-9.e+-]+)", txt)
        assert m and float(m.group(1)) <= bar, txt[-1500:]
        m = re.search(r"# GPU performance \(lm,ln,tune\)=\( 16, 16,0\) is  ([0-9.]+) G[fF]lop/sec", txt)
        assert m and float(m.group(1)) > 1000., txt[-1500:]


@pytest.mark.parametrize("lmln", [(32, 32), (32, 64), (64, 64)], ids=lambda v: f"{v[0]}x{v[1]}")
def test_bench_cli_multiply_other_block_sizes_print_no_warning(lmln, tmp_path):
    """`bench_tfqmrgpu multiply <plan> f 1 1 LM LN`: the harness prints "Warning! GPU result has large deviations" and withholds the
    performance line when maxdev > 1e-4 (bench_tfqmrgpu.cu:414).  On this library's default tensor-core product it must not."""
    from tfqmrgpu_b200 import formats as F
    g = np.load(os.path.join(ROOT, "tests", "golden", "plan_unordered.npz"))
    plan = str(tmp_path / "plan_unordered.14-287-16")
    F.write_multiplication_plan(plan, g["starts"], g["pairs"], int(g["nnz"][1]), int(g["nnz"][2]))
    txt = _cli("multiply", plan, "f", "1", "1", str(lmln[0]), str(lmln[1]))
    assert "Warning" not in txt, txt[-1500:]
    m = re.search(r"# GPU maxdev ([0-9.e+-]+) avgdev ([0-9.e+-]+)", txt)
    assert m and float(m.group(1)) <= 1e-4, txt[-1500:]
    assert "# GPU performance" in txt


@pytest.mark.skipif(not _have("bench_tfqmrgpu_ours"), reason="reference bench not built")
def test_bench_cli_tfqmr_matches_the_reference_harness(tmp_path):
    """`tfQMR <file> z`: same `# GPU maxdev ... avgdev ...` line as the reference's harness linked against this library, for the
    XML file and for the legacy dump of the same problem."""
    from tfqmrgpu_b200 import formats as F
    gold = os.path.join(ROOT, "tests", "golden", "FD_problem.xml")
    legacy = str(tmp_path / "fd_problem.txt")
    F.write_legacy(legacy, P.read_xml(gold))
    out = subprocess.run([os.path.join(REFDIR, "bench_tfqmrgpu_ours"), "tfQMR", gold, "z"], capture_output=True, text=True, timeout=300)
    ref = re.search(r"# GPU maxdev ([0-9.e+-]+) avgdev ([0-9.e+-]+) maxrel ([0-9.e+-]+)", out.stdout)
    assert ref, out.stdout[-1500:]
    for f in (gold, legacy):
        txt = _cli("tfQMR", f, "z", "1", "2000")
        m = re.search(r"# GPU maxdev ([0-9.e+-]+) avgdev ([0-9.e+-]+) maxrel ([0-9.e+-]+)", txt)
        assert m and m.groups() == ref.groups(), (m and m.groups(), ref.groups())
        s = re.search(r"# solve: status (\d+), (\d+) iterations, residual ([0-9.e+-]+)", txt)
        assert s and int(s.group(1)) == 0 and 40 <= int(s.group(2)) <= 43 and float(s.group(3)) < 1e-9
        assert "# found tolerance= 1e-09" in txt and "# requested precision= 'z' for LM= 8, LN= 8" in txt


CASES = golden_cases()


@pytest.mark.skipif(not _have("libtfqmr_ref_gpu.so"), reason="oracle/_ref/libtfqmr_ref_gpu.so not built")
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_solve_vs_reference_cuda_build_same_gpu(case):
    """Same inputs, same cuRAND v3 (both draw XORWOW seed 1234 in the caller's block order): status and iteration
    count identical, residual and X within the stated tolerance, flop count identical when iterations agree."""
    name, prob, prec, tol, maxit, tA, tB = case
    dt = np.float64 if prec == "z" else np.float32
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
    ref = O.ref_gpu()
    with O.quiet_stdout():
        r = ref.solve(prob.mb, prob.lm, prob.ln, prob.A.rowptr, prob.A.colind, vA, prob.X.rowptr, prob.X.colind,
                      prob.B.rowptr, prob.B.colind, vB, tol, maxit, prec, transA=tA, trans_b=tB)
    h = api.Handle()
    pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
    pl.buffer_size_for(prob.lm, prob.ln, prec); pl.set_buffer()
    assert np.array_equal(pl.get_v3().reshape(-1), r["v3"])          # identical shadow vector
    pl.set_matrix("A", vA, tA); pl.set_matrix("B", vB, tB)
    st = pl.solve(tol, maxit)
    info = pl.info()
    X = pl.get_matrix("X", "n", L.LAYOUT_RRRRIIII).reshape(pl.nnzbX, 2, prob.lm, prob.ln)
    lists = pl.plan_lists()
    pl.close(); h.close()
    for k in ("starts", "pairs", "subset", "colindx"):
        assert np.array_equal(lists[k], r["lists"][k]), k
    assert st == r["status"]
    assert abs(info["iterations"] - r["iterations"]) <= 1
    if info["iterations"] == r["iterations"]:
        assert info["flops"] == r["flops"]
    assert info["residuum"] <= tol and r["residuum"] <= tol
    scale = np.abs(r["X"]).max()
    assert np.abs(X - r["X"]).max() <= (10 if prec == "z" else 50)*tol*scale


@pytest.mark.skipif(not _have("libtfqmr_ref_gpu.so"), reason="oracle/_ref/libtfqmr_ref_gpu.so not built")
def test_config3_full_size_vs_reference_cuda_build():
    """BASELINE config 3 at full size (27-point stencil, 32^3 block rows of 32 x 32 complex fp32 blocks, 64 right-hand sides):
    the reference's own CUDA kernels (unmodified sources, sm_100 build) and this library on the same GPU with the same operands
    and the same cuRAND shadow vector: plan lists bit-identical, same status, iterations within one, X within 50*tol*max|X|."""
    import torch
    from tfqmrgpu_b200 import synthetic
    n, lm, ln, ncols, tol, maxit = 32, 32, 32, 2, 1e-3, 100
    sp = synthetic.Stencil27(n, lm, ln, ncols, sigma=8.0, dtype=np.float32, device="cuda")
    vA = sp.valA_host.numpy().reshape(-1); vB = sp.valB.reshape(-1)
    with O.quiet_stdout():
        r = O.ref_gpu().solve(sp.mb, lm, ln, sp.rpA, sp.ciA, vA, sp.rpX, sp.ciX, sp.rpB, sp.ciB, vB, tol, maxit, "c")
    h = api.Handle()
    pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
    pl.buffer_size_for(lm, ln, "c"); pl.set_buffer()
    assert pl.plan_info()["use_tc"] == 1
    assert np.array_equal(pl.get_v3().reshape(-1), r["v3"])
    pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr()); pl.set_matrix("B", sp.valB)
    st = pl.solve(tol, maxit)
    info = pl.info()
    X = pl.get_matrix("X", "n", L.LAYOUT_RRRRIIII).reshape(pl.nnzbX, 2, lm, ln)
    lists = pl.plan_lists()
    pl.close(); h.close()
    for k in ("starts", "pairs", "subset", "colindx"):
        assert np.array_equal(lists[k], r["lists"][k]), k
    assert st == r["status"] == 0
    assert abs(info["iterations"] - r["iterations"]) <= 1
    assert info["residuum"] <= tol and r["residuum"] <= tol
    assert np.abs(X - r["X"]).max() <= 50*tol*np.abs(r["X"]).max()
    torch.cuda.empty_cache()



def test_c_example_mixed_precision_one_character_change():
    """examples/mixed_precision.c: a plain C caller (tfqmrgpu.h only) solves the same system with precision 'z' and with 'm' - the
    precision the reference documents (tfqmrgpu.h:72) and refuses (tfqmrgpu.cu:42-44) - and compares the solutions."""
    exe = os.path.join(ROOT, "examples", "_build", "mixed_precision")
    if not os.path.exists(exe):
        pytest.skip("examples/_build/mixed_precision not built (make -C examples)")
    for args in (["64", "16", "2", "1e-10"], ["40", "32", "3", "1e-9"], ["30", "8", "1", "1e-10"]):
        out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "mixed_precision: OK" in out.stdout, out.stdout + out.stderr
        m = re.search(r"precision m: status 0, (\d+) iterations, residual ([0-9.e+-]+)", out.stdout)
        assert m and int(m.group(1)) > 0 and float(m.group(2)) <= float(args[3]), out.stdout
    env = dict(os.environ, TFQMRGPU_MIXED="0")           # the reference's behaviour: 'm' is refused with PRECISION_MISSMATCH (16)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode != 0 and "status 1090" in out.stdout.replace(",", ""), out.stdout + out.stderr
