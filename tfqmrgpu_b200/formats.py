"""Problem files either side of the solve: the reference's two input formats.

* ``LinearProblem`` XML -- read by ``tfqmrgpu_example_xml_reader.hxx:105-295``, written by the FD generator
  (``example/tfqmrgpu_generate_FD_example.cxx:156-234`` and ``:859-874``).  ``write_xml`` emits the generator's exact
  text (number formats, 16 integers per line, blank line after each block), so a file read with ``read_xml_raw``
  and written back is byte-identical (tests: md5 of ``FD_problem.xml``).
* the legacy Fortran text dump -- read by ``tfqmrgpu_example_reader.hxx:41-216`` (keyword lines ``nRHSs``, ``nCols``,
  ``tolerance``, ``bsr_?%nCols``, ``sizebsr_?%RowStart``, ``sizebsr_?%ColIndex``, ``shapemat_?``; 1-based indices).

Blocks are kept exactly as stored in the files, ``val[nnzb][slow][fast]``; the reference bench uploads such data with
trans ``'t'`` (``bench_tfqmrgpu.cu:153-157``).  Host-side code only; nothing here touches the GPU.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from .problems import Bsr, Problem


@dataclass
class XmlOperator:
    """One ``<BlockSparseMatrix>``: CSR pattern, optional indirection into the stored blocks, the stored blocks."""
    id: str
    rp: np.ndarray                      # int32 [rows + 1]
    ci: np.ndarray                      # int32 [nnzb]
    data: np.ndarray                    # float64 or complex128 [stored blocks][slow][fast], unscaled
    dims: tuple[int, int]               # (slow, fast) block dimensions (also when no block is stored)
    indirection: np.ndarray | None = None   # int64 [nnzb] -> stored block, None: block inzb is stored block inzb
    scale: float = 1.0
    scale_text: str | None = None       # the attribute as found in a file (kept for byte-exact rewriting)
    rowstart: bool = False              # pattern given as <RowStart> instead of <NonzerosPerRow>

    @property
    def is_complex(self) -> bool:
        return np.iscomplexobj(self.data)

    def values(self) -> np.ndarray:
        """complex128 [nnzb][slow][fast]: indirection and scale applied (xml_reader.hxx:268-283)."""
        if self.data.shape[0] < 1:
            return np.zeros((self.ci.size,) + tuple(self.dims), np.complex128)
        ind = self.indirection if self.indirection is not None else np.arange(self.ci.size)
        if ind.size and (ind.min() < 0 or ind.max() >= self.data.shape[0]):
            raise ValueError(f"operator {self.id}: indirection outside of the {self.data.shape[0]} stored blocks")
        return self.data[ind].astype(np.complex128)*self.scale


@dataclass
class XmlProblem:
    tolerance: float
    operators: list[XmlOperator] = field(default_factory=list)     # in file order (the generator writes A, B, X)
    comment: str | None = None          # text of the first comment, verbatim
    tolerance_text: str | None = None

    def op(self, name: str) -> XmlOperator:
        for o in self.operators:
            if o.id[:1] == name:
                return o
        raise KeyError(name)

    def to_problem(self, name: str = "xml") -> Problem:
        A, B, X = self.op("A"), self.op("B"), self.op("X")
        return Problem(Bsr(A.rp, A.ci, A.values()), Bsr(X.rp, X.ci, X.values()), Bsr(B.rp, B.ci, B.values()),
                       A.dims[1], B.dims[1], self.tolerance, name, X_exact=None)


def _ints(text: str | None, dtype) -> np.ndarray:
    return np.array((text or "").split(), dtype=dtype)


def read_xml_raw(path: str) -> XmlProblem:
    """Parse a LinearProblem file keeping what is needed to write it back unchanged."""
    with open(path, "r") as f:
        text = f.read()
    root = ET.fromstring(text)
    if root.tag != "LinearProblem":
        raise ValueError(f"{path}: root element is <{root.tag}>, expected <LinearProblem>")
    m = re.search(r"<!--(.*?)-->", text, flags=re.S)
    xp = XmlProblem(float(root.attrib.get("tolerance", "0")), comment=(m.group(1) if m else None),
                    tolerance_text=root.attrib.get("tolerance"))
    for bsm in root:
        oid = bsm.attrib.get("id", "?")
        sm = bsm.find("SparseMatrix")
        if sm is None:
            raise ValueError(f"{path}: operator {oid} has no <SparseMatrix>")
        csr = sm.find("CompressedSparseRow")
        if csr is None:
            raise ValueError(f"{path}: operator {oid} has no <CompressedSparseRow>")
        nzpr, rowstart = csr.find("NonzerosPerRow"), csr.find("RowStart")
        if nzpr is not None:
            rp = np.concatenate([[0], np.cumsum(_ints(nzpr.text, np.int64))]).astype(np.int32)
        elif rowstart is not None:
            rp = _ints(rowstart.text, np.int32)
        else:
            raise ValueError(f"{path}: operator {oid} has neither <NonzerosPerRow> nor <RowStart>")
        col = csr.find("ColumnIndex")
        if col is None:
            raise ValueError(f"{path}: operator {oid} has no <ColumnIndex>")
        ci = _ints(col.text, np.int32)
        if ci.size != rp[-1]:
            raise ValueError(f"{path}: operator {oid}: {ci.size} column indices for {rp[-1]} nonzero blocks")
        ind = sm.find("Indirection")
        indirection = _ints(ind.text, np.int64) if ind is not None else None
        if indirection is not None and indirection.size != ci.size:
            raise ValueError(f"{path}: operator {oid}: indirection has {indirection.size} entries for {ci.size} blocks")
        dt = bsm.find("DataTensor")
        if dt is None:
            raise ValueError(f"{path}: operator {oid} has no <DataTensor>")
        d = [int(v) for v in dt.attrib.get("dimensions", "0 0 0").split()]
        is_complex = dt.attrib.get("type", "complex")[:1].lower() == "c"
        raw = np.array((dt.text or "").split(), dtype=np.float64)
        want = d[0]*d[1]*d[2]*(2 if is_complex else 1)
        if raw.size != want:
            raise ValueError(f"{path}: operator {oid}: {raw.size} numbers in the DataTensor, dimensions ask for {want}")
        if is_complex:
            raw = raw.reshape(d[0], d[1], d[2], 2)
            data = raw[..., 0] + 1j*raw[..., 1]
        else:
            data = raw.reshape(d[0], d[1], d[2])
        xp.operators.append(XmlOperator(oid, rp, ci, data, (d[1], d[2]), indirection, float(dt.attrib.get("scale", "1")),
                                        dt.attrib.get("scale"), rowstart=(nzpr is None)))
    return xp


def _int_lines(values) -> str:
    """'\\n' before every 16th number, ' ' before the others (generate_FD_example.cxx:169-172)."""
    return "".join(("\n" if 0 == (i & 15) else " ") + "%d" % v for i, v in enumerate(values))


def write_xml(path: str, xp: XmlProblem, lossless: bool = False) -> None:
    """Write ``xp`` in the generator's format.  ``lossless``: 17 significant digits for both parts of every number
    (the generator prints the imaginary parts with ``%g``, i.e. 6 digits)."""
    re_fmt, im_fmt = ("%.17g ", " %.17g  ") if lossless else ("%.15g ", " %g  ")
    out = ['<?xml version="1.0"?>\n',
           '<LinearProblem problem_kind="A*X==B"\n'
           '               generator_version="0.1" tolerance="%s">\n'
           % (xp.tolerance_text if xp.tolerance_text is not None else "%.3e" % xp.tolerance)]
    if xp.comment is not None:
        out.append("  <!--%s-->\n" % xp.comment)
    for o in xp.operators:
        sp = "    "
        out.append('  <BlockSparseMatrix id="%s">\n' % o.id)
        out.append('%s<SparseMatrix type="CSR">\n%s  <CompressedSparseRow>\n' % (sp, sp))
        if o.rowstart:
            out.append('%s    <RowStart rows="%d">%s\n%s    </RowStart>\n' % (sp, o.rp.size - 1, _int_lines(o.rp), sp))
        else:
            out.append('%s    <NonzerosPerRow rows="%d">%s\n%s    </NonzerosPerRow>\n'
                       % (sp, o.rp.size - 1, _int_lines(np.diff(o.rp)), sp))
        out.append('%s    <ColumnIndex nonzeros="%d">%s\n%s    </ColumnIndex>\n' % (sp, o.ci.size, _int_lines(o.ci), sp))
        out.append('%s  </CompressedSparseRow>\n' % sp)
        if o.indirection is not None:
            out.append('%s  <Indirection nonzeros="%d">%s\n%s  </Indirection>\n' % (sp, o.ci.size, _int_lines(o.indirection), sp))
        out.append('%s</SparseMatrix>\n' % sp)
        nb, (d1, d2) = o.data.shape[0], o.dims
        out.append('    <DataTensor type="%s" rank="3" dimensions="%d %d %d"' % ("complex64" if o.is_complex else "real", nb, d1, d2))
        if o.scale_text is not None:
            out.append(' scale="%s"' % o.scale_text)
        elif 1 != o.scale:
            out.append(' scale="%.16e"' % o.scale)
        out.append(">\n")
        for b in range(nb):
            for i in range(d1):
                if o.is_complex:
                    out.append("".join(re_fmt % v.real + im_fmt % v.imag for v in o.data[b, i]) + "\n")
                else:
                    out.append("".join(re_fmt % v for v in o.data[b, i]) + "\n")
            if d1 > 1:
                out.append("\n")
        out.append("    </DataTensor>\n  </BlockSparseMatrix>\n")
    out.append("</LinearProblem>\n")
    with open(path, "w") as f:
        f.write("".join(out))


def xml_from_problem(p: Problem, tolerance: float | None = None, comment: str | None = None,
                     store_x: bool = False) -> XmlProblem:
    """A, B (and optionally X) of a Problem as an XmlProblem: no indirection, scale 1, type real if every imaginary part is 0.
    X is written as a pattern with an empty DataTensor, as the generator does (the solver ignores the initial X)."""
    def one(name: str, bsr: Bsr, dims, with_values: bool) -> XmlOperator:
        val = np.asarray(bsr.val) if with_values and bsr.val is not None else np.zeros((0,) + tuple(dims), np.float64)
        if np.iscomplexobj(val) and not np.any(val.imag):
            val = val.real.astype(np.float64)
        return XmlOperator(name, np.asarray(bsr.rowptr, np.int32), np.asarray(bsr.colind, np.int32), val, tuple(dims))
    dA = tuple(p.A.val.shape[1:]); dB = tuple(p.B.val.shape[1:])
    tol = p.tolerance if tolerance is None else tolerance
    return XmlProblem(tol, [one("A", p.A, dA, True), one("B", p.B, dB, True), one("X", p.X, dB, store_x)], comment)


# ------------------------------------------------------------------------------------------------
def write_legacy(path: str, p: Problem, tolerance: float | None = None) -> None:
    """The Fortran text dump the reference bench reads when the file name does not contain 'xml'
    (tfqmrgpu_example_reader.hxx:41-216): 1-based RowStart/ColIndex, every operator with values, blocks in file order
    ``[block][slow][fast][Re,Im]``."""
    lm = p.A.val.shape[2]
    ncols = int(max(np.max(p.X.colind, initial=-1), np.max(p.B.colind, initial=-1))) + 1
    out = ["nRHSs %d\n" % lm, "nCols %d\n" % ncols, "tolerance %.17g\n" % (p.tolerance if tolerance is None else tolerance)]
    for name, bsr in (("A", p.A), ("B", p.B), ("X", p.X)):
        nrows = bsr.rowptr.size - 1
        nc = nrows if "A" == name else ncols
        val = bsr.val if bsr.val is not None else None
        if val is None or val.shape[0] != bsr.colind.size:
            shape = p.B.val.shape[1:]
            val = np.zeros((bsr.colind.size,) + tuple(shape), np.complex128)
        out.append("bsr_%s%%nCols %d\n" % (name, nc))
        out.append("sizebsr_%s%%RowStart %d\n" % (name, nrows + 1))
        out.append(_int_lines(np.asarray(bsr.rowptr, np.int64) + 1).lstrip("\n") + "\n")
        out.append("sizebsr_%s%%ColIndex %d\n" % (name, bsr.colind.size))
        out.append(_int_lines(np.asarray(bsr.colind, np.int64) + 1).lstrip("\n") + "\n")
        out.append("shapemat_%s %d %d %d\n" % (name, val.shape[2], val.shape[1], val.shape[0]))
        flat = np.empty(val.size*2, np.float64)
        flat[0::2] = val.real.ravel(); flat[1::2] = val.imag.ravel()
        for i in range(0, flat.size, 8):
            out.append(" ".join("%.17g" % v for v in flat[i:i + 8]) + "\n")
    with open(path, "w") as f:
        f.write("".join(out))


def read_legacy(path: str) -> Problem:
    """Reader for the same dump.  Like the reference it appends empty rows to B when B has fewer rows than X (:196-203)."""
    with open(path, "r") as f:
        tok = f.read().split()
    pos, n = 0, len(tok)
    block_size = ncols = 0
    tol = 0.0
    ops: dict[str, dict] = {k: {} for k in "ABX"}

    def take(count: int, dtype):
        nonlocal pos
        if pos + count > n:
            raise ValueError(f"{path}: file ends inside a list of {count} numbers")
        a = np.array(tok[pos:pos + count], dtype=dtype); pos += count
        return a

    while pos < n:
        key = tok[pos]; pos += 1
        if "nRHSs" == key:
            block_size = int(take(1, np.int64)[0])
        elif "nCols" == key:
            ncols = int(take(1, np.int64)[0])
        elif "tolerance" == key:
            tol = float(take(1, np.float64)[0])
        elif re.fullmatch(r"bsr_[ABX]%nCols", key):
            ops[key[4]]["ncols"] = int(take(1, np.int64)[0])
        elif re.fullmatch(r"sizebsr_[ABX]%RowStart", key):
            cnt = int(take(1, np.int64)[0])
            ops[key[8]]["rp"] = (take(cnt, np.int64) - 1).astype(np.int32)
        elif re.fullmatch(r"sizebsr_[ABX]%ColIndex", key):
            cnt = int(take(1, np.int64)[0])
            ops[key[8]]["ci"] = (take(cnt, np.int64) - 1).astype(np.int32)
        elif re.fullmatch(r"shapemat_[ABX]", key):
            n1, n2, n3 = (int(v) for v in take(3, np.int64))
            raw = take(n3*n2*n1*2, np.float64).reshape(n3, n2, n1, 2)
            ops[key[9]]["val"] = raw[..., 0] + 1j*raw[..., 1]
        else:
            raise ValueError(f"{path}: keyword {key} unknown")
    for k, o in ops.items():
        for need in ("rp", "ci", "val"):
            if need not in o:
                raise ValueError(f"{path}: operator {k} lacks {need}")
        if o["ci"].size != o["rp"][-1] or o["val"].shape[0] != o["ci"].size:
            raise ValueError(f"{path}: operator {k}: list lengths disagree")
    A, B, X = ops["A"], ops["B"], ops["X"]
    if A["val"].shape[1] != A["val"].shape[2] or (block_size and A["val"].shape[2] != block_size):
        raise ValueError(f"{path}: A blocks must be square with edge nRHSs")
    if B["rp"].size < X["rp"].size:
        B["rp"] = np.concatenate([B["rp"], np.full(X["rp"].size - B["rp"].size, B["rp"][-1], np.int32)])
    if X["rp"].size != A["rp"].size or B["rp"].size != A["rp"].size:
        raise ValueError(f"{path}: A, B and X must have the same number of block rows")
    del ncols
    return Problem(Bsr(A["rp"], A["ci"], A["val"]), Bsr(X["rp"], X["ci"], X["val"]), Bsr(B["rp"], B["ci"], B["val"]),
                   A["val"].shape[2], B["val"].shape[2], tol, path, X_exact=None)


# ------------------------------------------------------------------------------------------------
def write_multiplication_plan(path: str, starts: np.ndarray, pairs: np.ndarray, nnzA: int, nnzX: int) -> None:
    """The plan dump ``bench_tfqmrgpu multiply`` reads (bench_tfqmrgpu.cu:456-498; test/multiplication/plan_*): a header
    ``#nnzb_for_Y_A_X= nY nA nX`` and one line ``iY iA iX beta`` per pair, beta = 0 on the first pair of a Y block."""
    starts = np.asarray(starts, np.int64); pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    nY = starts.size - 1
    out = ["#nnzb_for_Y_A_X= %d %d %d \n" % (nY, nnzA, nnzX)]
    for y in range(nY):
        for p in range(starts[y], starts[y + 1]):
            out.append("%d %d %d %d \n" % (y, pairs[p, 0], pairs[p, 1], 0 if p == starts[y] else 1))
    with open(path, "w") as f:
        f.write("".join(out))

