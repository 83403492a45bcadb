"""Problem-file formats either side of the solve: LinearProblem XML and the legacy Fortran text
dump, read AND written.  Parity is pinned two ways: the writer reproduces the reference generator's FD_problem.xml byte for
byte, and the reference's own readers (its unmodified bench harness on the reference CPU path, oracle/_ref) run files we
wrote to the same iteration log as the original."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from tfqmrgpu_b200 import formats as F, problems as P

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden", "FD_problem.xml")
REF = os.path.join(ROOT, "oracle", "_ref")


def test_xml_rewrite_is_byte_identical(tmp_path):
    """read -> write gives the generator's file back: md5 8bbd6b1f... (SURVEY.md section 8c fixture table)."""
    xp = F.read_xml_raw(GOLDEN)
    out = tmp_path / "FD_problem.xml"
    F.write_xml(str(out), xp)
    data = out.read_bytes()
    assert data == open(GOLDEN, "rb").read()
    assert hashlib.md5(data).hexdigest() == "8bbd6b1fda2c267fa7e5d49aaf5c67a6"
    A = xp.op("A")
    assert A.indirection is not None and A.data.shape == (13, 8, 8) and not A.is_complex
    assert abs(A.scale - 1/5040.) < 1e-18 and xp.op("X").data.shape[0] == 0 and xp.tolerance == 1e-9


def _same(p, q):
    for a, b in ((p.A, q.A), (p.B, q.B), (p.X, q.X)):
        assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.colind, b.colind)
    assert np.array_equal(p.A.val, q.A.val) and np.array_equal(p.B.val, q.B.val)
    assert (p.lm, p.ln) == (q.lm, q.ln)


@pytest.mark.parametrize("lm,ln", [(4, 4), (4, 5), (8, 16)])
def test_xml_roundtrip_lossless(tmp_path, lm, ln):
    """complex blocks, rectangular B/X blocks, no indirection: every double survives (17 digits)."""
    p = P.random_system(9, lm, ln, seed=3)
    f = str(tmp_path / "p.xml")
    F.write_xml(f, F.xml_from_problem(p, tolerance=1e-7, comment=" made by test "), lossless=True)
    q = P.read_xml(f)
    _same(p, q)
    assert q.tolerance == 1e-7 and F.read_xml_raw(f).comment == " made by test "
    assert q.X.val.shape == (p.X.nnzb, lm, ln) and not np.any(q.X.val)      # pattern only, like the generator's X


def test_xml_reference_format_keeps_6_digits_of_imaginary_parts(tmp_path):
    """the generator prints Im with %g: a faithful writer loses digits there, and says so by having the lossless switch"""
    p = P.random_system(5, 4, 4, seed=1)
    f = str(tmp_path / "p.xml")
    F.write_xml(f, F.xml_from_problem(p))
    q = P.read_xml(f)
    assert np.allclose(q.A.val.real, p.A.val.real, rtol=1e-14, atol=0)
    assert np.allclose(q.A.val.imag, p.A.val.imag, rtol=1e-5, atol=0) and not np.array_equal(q.A.val.imag, p.A.val.imag)


def test_xml_rowstart_and_indirection(tmp_path):
    xp = F.read_xml_raw(GOLDEN)
    for o in xp.operators:
        o.rowstart = True
    f = str(tmp_path / "rs.xml")
    F.write_xml(f, xp)
    assert "<RowStart" in open(f).read() and "<NonzerosPerRow" not in open(f).read()
    _same(P.read_xml(f), P.read_xml(GOLDEN))
    bad = F.read_xml_raw(GOLDEN)
    bad.op("A").indirection[5] = 13                     # only 13 blocks are stored
    with pytest.raises(ValueError):
        bad.to_problem()


def test_xml_malformed(tmp_path):
    text = open(GOLDEN).read()
    f = tmp_path / "bad.xml"
    f.write_text(text.replace('<ColumnIndex nonzeros="1557">\n0 1 7', '<ColumnIndex nonzeros="1557">\n1 7', 1))
    with pytest.raises(ValueError):
        F.read_xml_raw(str(f))
    f.write_text(text.replace('dimensions="13 8 8"', 'dimensions="14 8 8"', 1))
    with pytest.raises(ValueError):
        F.read_xml_raw(str(f))
    f.write_text(text.replace("LinearProblem", "SomethingElse"))
    with pytest.raises(ValueError):
        F.read_xml_raw(str(f))


@pytest.mark.parametrize("lm,ln", [(4, 4), (8, 8)])
def test_legacy_roundtrip(tmp_path, lm, ln):
    """the Fortran dump needs square blocks of one size for A, B and X (reader asserts, example_reader.hxx:113-114)"""
    p = P.random_system(7, lm, ln, seed=5)
    f = str(tmp_path / "problem.txt")
    F.write_legacy(f, p, tolerance=2.5e-8)
    q = F.read_legacy(f)
    _same(p, q)
    assert q.tolerance == 2.5e-8
    head = open(f).read().split("\n")
    assert head[0] == "nRHSs %d" % lm and head[3].startswith("bsr_A%nCols") and head[4] == "sizebsr_A%%RowStart %d" % 8
    assert head[5].split()[0] == "1"                                          # 1-based


def test_legacy_fewer_b_rows_and_errors(tmp_path):
    p = P.read_xml(GOLDEN)
    f = str(tmp_path / "fd.txt")
    F.write_legacy(f, p)
    text = open(f).read()
    # drop the trailing empty rows of B, as a Fortran code that only knows its source rows would
    rpB = np.asarray(p.B.rowptr) + 1
    last = int(np.flatnonzero(np.diff(p.B.rowptr))[-1]) + 1
    lines = text.split("\n")
    i = lines.index("sizebsr_B%%RowStart %d" % rpB.size)
    j = lines.index("sizebsr_B%%ColIndex %d" % p.B.nnzb)
    lines[i:j] = ["sizebsr_B%%RowStart %d" % (last + 1), " ".join(str(v) for v in rpB[:last + 1])]
    open(f, "w").write("\n".join(lines))
    q = F.read_legacy(f)
    _same(p, q)
    open(f, "w").write(text.replace("shapemat_X", "shapemat_Q"))
    with pytest.raises(ValueError):
        F.read_legacy(f)
    open(f, "w").write(text[:len(text)//2])
    with pytest.raises(ValueError):
        F.read_legacy(f)


def _ref_bench_log(path):
    exe, pre = os.path.join(REF, "bench_tfqmrgpu_cpu"), os.path.join(REF, "libalign256.so")
    env = dict(os.environ, LD_PRELOAD=pre, OMP_NUM_THREADS="1")
    r = subprocess.run([exe, "tfQMR", path, "z", "1", "2000"], capture_output=True, text=True, timeout=300, env=env,
                       cwd=os.path.dirname(path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    keep = [ln for ln in r.stdout.split("\n") if re.match(r"# (in iteration|ran \d+ iterations|GPU maxdev|norms of B|nnzbB)", ln)]
    assert any(ln.startswith("# ran ") for ln in keep), r.stdout[-2000:]
    return keep


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "bench_tfqmrgpu_cpu")), reason="oracle/_ref/bench_tfqmrgpu_cpu not built")
def test_reference_readers_accept_our_files(tmp_path):
    """The reference's bench harness (its XML reader and its legacy reader, unmodified, on the reference CPU solver) runs the
    original FD_problem.xml, our XML without indirection/scale, and our legacy dump to the same iteration log."""
    p = P.read_xml(GOLDEN)
    ours_xml, ours_txt = str(tmp_path / "ours.xml"), str(tmp_path / "ours_problem.txt")
    F.write_xml(ours_xml, F.xml_from_problem(p), lossless=True)
    F.write_legacy(ours_txt, p)
    gold = _ref_bench_log(GOLDEN)
    assert "# ran 42 iterations" in gold
    assert _ref_bench_log(ours_xml) == gold
    assert _ref_bench_log(ours_txt) == gold


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "bench_tfqmrgpu_cpu")), reason="oracle/_ref/bench_tfqmrgpu_cpu not built")
def test_reference_readers_agree_on_a_complex_problem(tmp_path):
    """A complex-valued system (the FD file is real): our XML (type complex64, 17 digits) and our legacy dump of the same problem
    run to the same iteration log through the reference's two readers."""
    p = P.random_system(9, 4, 4, ncols=5, seed=11)
    xml, txt = str(tmp_path / "cplx.xml"), str(tmp_path / "cplx_problem.txt")
    F.write_xml(xml, F.xml_from_problem(p, tolerance=1e-8), lossless=True)
    F.write_legacy(txt, p, tolerance=1e-8)
    assert F.read_xml_raw(xml).op("A").is_complex
    a, b = _ref_bench_log(xml), _ref_bench_log(txt)
    assert a == b and any(ln.startswith("# ran ") for ln in a)


def test_multiplication_plan_file_roundtrip(tmp_path, plan_unordered=None):
    """write_multiplication_plan produces the format of test/multiplication/plan_unordered.14-287-16: reading it back gives the
    lists, and the golden plan (committed as arrays) survives the trip."""
    g = np.load(os.path.join(HERE, "golden", "plan_unordered.npz"))
    starts, pairs = g["starts"], g["pairs"].reshape(-1, 2)
    nnzA = int(g["nnz"][1])
    f = str(tmp_path / "plan.txt")
    F.write_multiplication_plan(f, starts, pairs, nnzA, starts.size - 1)
    s2, p2, nY, nA, nX, _ = P.read_multiplication_plan(f)
    assert np.array_equal(s2, starts) and np.array_equal(p2, pairs) and (nY, nA, nX) == (starts.size - 1, nnzA, starts.size - 1)
    assert open(f).readline() == "#nnzb_for_Y_A_X= 4490 13109 4490 \n"



def test_bench_cli_fill_and_checker_match_the_oracle():
    """bench_cli's cos/sin fill and its numpy check of Y = A*X (the CLI's stand-in for the harness's CPU check loop,
    bench_tfqmrgpu.cu:277-285 and :366-392) against the oracle's restatement of both."""
    import orclib as O
    from tfqmrgpu_b200 import bench_cli as B
    for lm, ln, dt in ((4, 5, np.float32), (8, 8, np.float64)):
        assert np.array_equal(B.fill_cos_sin(7, lm, ln, dt), O.fill_cos_sin(7, lm, ln, dt))
    g = np.load(os.path.join(HERE, "golden", "plan_unordered.npz"))
    starts, pairs = g["starts"][:201], g["pairs"].reshape(-1, 2)
    pairs = pairs[:starts[-1]]
    nA, nX = int(pairs[:, 0].max()) + 1, int(pairs[:, 1].max()) + 1
    A = B.fill_cos_sin(nA, 16, 16, np.float64); X = B.fill_cos_sin(nX, 16, 16, np.float64)
    Y = B._pairs_product(A, X, starts, pairs)
    Yo = O.multiply(A, X, starts, pairs, 16, 16)[:200]
    assert np.abs(np.stack([Y.real, Y.imag], axis=1) - Yo).max() <= 1e-12
