"""Problem generators and readers for the tfQMR hot path (host side, numpy only).

Everything here produces plain BSR index arrays + complex block values in the HOST layout that the
reference's C-ABI takes (``val[nnzb][rows][cols][2]``, TFQMRGPU_LAYOUT_RIRIRIRI, trans 'n').

Sources mirrored (file:line relative to the reference checkout):
  * Julia known-answer test        example/tfqmrgpu_Julia_example.jl:10-66
  * Fortran example patterns       example/tfqmrgpu_Fortran_example.F90:22-44
  * C example random system        example/tfqmrgpu_C_example.c:47-137 (numpy RNG instead of glibc rand)
  * XML LinearProblem reader       tfQMRgpu/include/tfqmrgpu_example_xml_reader.hxx:105-295
  * multiplication plan file       tfQMRgpu/source/bench_tfqmrgpu.cu:456-498
  * synthetic 27-point stencil     SURVEY.md section 8(d) config 3 (counter-based hash values)
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class Bsr:
    """One block-sparse-row operator: pattern + complex blocks ``val[nnzb, rows, cols]``."""
    rowptr: np.ndarray          # int32[mb+1]
    colind: np.ndarray          # int32[nnzb]
    val: np.ndarray | None      # complex128[nnzb, rows, cols] or None (pattern only)

    @property
    def nnzb(self) -> int:
        return int(self.colind.shape[0])

    @property
    def mb(self) -> int:
        return int(self.rowptr.shape[0] - 1)


@dataclasses.dataclass
class Problem:
    """A*X == B in BSR form. ``lm`` = block rows of A/X/B, ``ln`` = block columns of X/B."""
    A: Bsr
    X: Bsr
    B: Bsr
    lm: int
    ln: int
    tolerance: float = 1e-9
    name: str = ""
    X_exact: np.ndarray | None = None   # complex[nnzbX, lm, ln] when an analytic solution is known

    @property
    def mb(self) -> int:
        return self.A.mb


def interleave(val: np.ndarray, dtype) -> np.ndarray:
    """complex[nnzb, r, c] -> real[nnzb, r, c, 2] (RIRIRIRI host layout)."""
    out = np.empty(val.shape + (2,), dtype=dtype)
    out[..., 0] = val.real
    out[..., 1] = val.imag
    return np.ascontiguousarray(out)


def deinterleave(arr: np.ndarray) -> np.ndarray:
    return arr[..., 0].astype(np.float64) + 1j*arr[..., 1].astype(np.float64)


# ------------------------------------------------------------------------------------------------
def julia_kat(ldA: int = 4, ldB: int = 5, mb: int = 7) -> Problem:
    """1-D finite-difference tridiag(-1,2,-1) (x) I_ldA with a single B block in the LAST block row.

    Mirrors example/tfqmrgpu_Julia_example.jl:40-66. Julia arrays are column-major
    ``Bmat[ldB, ldA, nnzb]`` which in C order is ``B[nnzb][ldA][ldB]``: B[0][j][i] = i^p for
    i in 0..ldB-1, j = i % ldA, p = i // ldA.  rowPtrB = [0,...,0,1] puts that block in the last block
    row (column index 0).  Exact solution: X_k = (k+1)/(mb+1) * B  for block row k = 0..mb-1.
    """
    rpA = [0]
    ciA, valA = [], []
    for ib in range(mb):
        for jb in range(max(0, ib - 1), min(mb - 1, ib + 1) + 1):
            ciA.append(jb)
            valA.append((2.0 if ib == jb else -1.0)*np.eye(ldA, dtype=np.complex128))
        rpA.append(len(ciA))
    Bblk = np.zeros((ldA, ldB), dtype=np.complex128)
    for i in range(ldB):
        Bblk[i % ldA, i] = (1j)**(i // ldA)
    A = Bsr(np.array(rpA, np.int32), np.array(ciA, np.int32), np.array(valA))
    X = Bsr(np.arange(mb + 1, dtype=np.int32), np.zeros(mb, np.int32), None)
    B = Bsr(np.array([0]*mb + [1], np.int32), np.zeros(1, np.int32), Bblk[None])
    exact = np.array([(k + 1)/(mb + 1.)*Bblk for k in range(mb)])
    return Problem(A, X, B, ldA, ldB, 1.2e-8, "julia_kat", exact)


def fortran_pattern(which: int, seed: int = 7) -> Problem:
    """The three sparsity patterns of example/tfqmrgpu_Fortran_example.F90:22-44 with seeded random A
    (diagonally boosted so that tfQMR converges) and a dense random B with X's pattern."""
    rng = np.random.default_rng(seed + which)
    if which == 0:      # one 32x32 block
        mb, lm, ln, pat = 1, 32, 32, [[0]]
    elif which == 1:    # dense 4x4 of 16x16 blocks
        mb, lm, ln, pat = 4, 16, 16, [list(range(4)) for _ in range(4)]
    else:               # tridiagonal 4x4 of 4x4 blocks
        mb, lm, ln, pat = 4, 4, 4, [[j for j in range(4) if abs(i - j) <= 1] for i in range(4)]
    rp, ci, vals = [0], [], []
    for i, cols in enumerate(pat):
        for j in cols:
            blk = rng.uniform(-1, 1, (lm, lm)) + 1j*rng.uniform(-1, 1, (lm, lm))
            if i == j:
                blk += (3. + 0.5*lm)*np.eye(lm)
            ci.append(j)
            vals.append(blk)
        rp.append(len(ci))
    A = Bsr(np.array(rp, np.int32), np.array(ci, np.int32), np.array(vals))
    # X and B dense in mb block columns
    rpx = np.arange(0, mb*mb + 1, mb, dtype=np.int32)
    cix = np.tile(np.arange(mb, dtype=np.int32), mb)
    valB = rng.uniform(-.5, .5, (mb*mb, lm, ln)) + 1j*rng.uniform(-.5, .5, (mb*mb, lm, ln))
    X = Bsr(rpx, cix, None)
    B = Bsr(rpx.copy(), cix.copy(), valB)
    return Problem(A, X, B, lm, ln, 1e-9, f"fortran_pattern{which}")


def random_system(mb: int, lm: int, ln: int, ncols: int | None = None, pA: float = .125, pX: float = .5,
                  pB: float = .125, seed: int = 1, unsorted: bool = False) -> Problem:
    """Random block patterns in the spirit of example/tfqmrgpu_C_example.c:47-137: A random with the
    diagonal always present and boosted by +3 (scaled with lm), X random with its diagonal, B a random
    subset of X plus the diagonal.  ``unsorted`` shuffles the column order inside every row of A, X
    and B (createPlan must keep A's given order and find X/B blocks by value)."""
    rng = np.random.default_rng(seed)
    ncols = ncols or (mb//2 + 1)
    nzA = (rng.random((mb, mb)) < pA) | np.eye(mb, dtype=bool)
    nzX = (rng.random((mb, ncols)) < pX)
    nzX[np.arange(min(mb, ncols)), np.arange(min(mb, ncols))] = True
    nzB = nzX & (rng.random((mb, ncols)) < pB)
    nzB[np.arange(min(mb, ncols)), np.arange(min(mb, ncols))] = True

    def pattern(nz):
        rp, ci = [0], []
        for i in range(nz.shape[0]):
            cols = np.flatnonzero(nz[i])
            if unsorted:
                cols = rng.permutation(cols)
            ci.extend(cols.tolist())
            rp.append(len(ci))
        return np.array(rp, np.int32), np.array(ci, np.int32)

    rpA, ciA = pattern(nzA)
    rpX, ciX = pattern(nzX)
    rpB, ciB = pattern(nzB)
    valA = rng.uniform(-1, 1, (len(ciA), lm, lm)) + 1j*rng.uniform(-1, 1, (len(ciA), lm, lm))
    rows = np.repeat(np.arange(mb), np.diff(rpA))
    # strong diagonal so that the system is comfortably solvable in both precisions
    boost = 3. + 1.5*lm*nzA.sum(axis=1).max()*.25
    valA[rows == ciA] += boost*np.eye(lm)
    valB = rng.uniform(-.5, .5, (len(ciB), lm, ln)) + 1j*rng.uniform(-.5, .5, (len(ciB), lm, ln))
    return Problem(Bsr(rpA, ciA, valA), Bsr(rpX, ciX, None), Bsr(rpB, ciB, valB), lm, ln, 1e-6,
                   f"random_mb{mb}_{lm}x{ln}")


# ------------------------------------------------------------------------------------------------
def _hash_uniform(idx: np.ndarray, seed: int) -> np.ndarray:
    """Counter-based hash -> uniform(-1,1) doubles; identical on every host (no RNG state)."""
    x = idx.astype(np.uint64) + np.uint64((int(seed)*0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(30); x = (x*np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(27); x = (x*np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64)*(2.0/9007199254740992.0) - 1.0


def stencil27_pattern(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Periodic n^3 grid, 27-point stencil, natural ordering: exactly 27 blocks per row (for n >= 3),
    column indices ascending."""
    idx = np.arange(n**3)
    z, y, x = idx//(n*n), (idx//n) % n, idx % n
    cols = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                cols.append(((z + dz) % n)*n*n + ((y + dy) % n)*n + ((x + dx) % n))
    cols = np.sort(np.stack(cols, axis=1), axis=1)
    if n < 3:  # duplicates collapse on tiny grids
        rows = [np.unique(r) for r in cols]
        rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
        return rp, np.concatenate(rows).astype(np.int32)
    rp = (27*np.arange(n**3 + 1)).astype(np.int32)
    return rp, cols.reshape(-1).astype(np.int32)


def stencil27_values_rows(rp: np.ndarray, ci: np.ndarray, lm: int, sigma: float, blocks: np.ndarray,
                          seed: int = 1234, dtype=np.float32) -> np.ndarray:
    """Host-layout values real[len(blocks), lm, lm, 2] of the A blocks with indices ``blocks``: diagonal block
    (27+sigma)*I + 0.05*U, off-diagonal -I + 0.05*U, U complex uniform(-1,1) from a hash of
    (block, i, k, re/im, seed).  Pure function of the block index, so any subset can be regenerated."""
    blocks = np.asarray(blocks, dtype=np.int64)
    rows = np.searchsorted(rp, blocks, side="right") - 1
    per = lm*lm*2
    idx = (blocks.astype(np.uint64)[:, None]*np.uint64(per) + np.arange(per, dtype=np.uint64)[None, :])
    u = _hash_uniform(idx, seed).reshape(blocks.size, lm, lm, 2)*0.05
    diag = (rows == ci[blocks])
    u[..., 0] += np.where(diag, 27. + sigma, -1.)[:, None, None]*np.eye(lm)[None]
    return u.astype(dtype)


def stencil27_values(rp: np.ndarray, ci: np.ndarray, lm: int, sigma: float, seed: int = 1234,
                     dtype=np.float32, chunk: int = 1 << 14) -> np.ndarray:
    """All A blocks, real[nnzb, lm, lm, 2] (see stencil27_values_rows)."""
    nnzb = int(ci.shape[0])
    out = np.empty((nnzb, lm, lm, 2), dtype=dtype)
    for b0 in range(0, nnzb, chunk):
        b1 = min(nnzb, b0 + chunk)
        out[b0:b1] = stencil27_values_rows(rp, ci, lm, sigma, np.arange(b0, b1), seed, dtype)
    return out


def stencil27(n: int, lm: int, ln: int, nrhs_blockcols: int, sigma: float = 1.0, seed: int = 1234,
              dtype=np.float32, unit_rhs: bool = True) -> dict:
    """Config-3 style synthetic problem: X dense in ``nrhs_blockcols`` block columns,
    B = one unit block per column (in block row = column index) or dense hashed values.
    Returns a dict of raw arrays already in the host layout/dtype for the C-ABI."""
    rpA, ciA = stencil27_pattern(n)
    mb = n**3
    valA = stencil27_values(rpA, ciA, lm, sigma, seed, dtype)
    rpX = (nrhs_blockcols*np.arange(mb + 1)).astype(np.int32)
    ciX = np.tile(np.arange(nrhs_blockcols, dtype=np.int32), mb)
    if unit_rhs:
        # block column c gets its unit block in block row c*(mb//ncols) (spread over the grid)
        brow = (np.arange(nrhs_blockcols)*(mb//nrhs_blockcols)).astype(np.int64)
        counts = np.zeros(mb, np.int64)
        counts[brow] += 1
        rpB = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        order = np.argsort(brow, kind="stable")
        ciB = np.arange(nrhs_blockcols, dtype=np.int32)[order]
        valB = np.zeros((nrhs_blockcols, lm, ln, 2), dtype=dtype)
        for j in range(ln):
            valB[:, j % lm, j, 0] = 1
    else:
        rpB, ciB = rpX.copy(), ciX.copy()
        idx = np.arange(mb*nrhs_blockcols*lm*ln*2, dtype=np.uint64)
        valB = (_hash_uniform(idx, seed + 17)*.5).reshape(mb*nrhs_blockcols, lm, ln, 2).astype(dtype)
    return dict(mb=mb, lm=lm, ln=ln, rpA=rpA, ciA=ciA, valA=valA, rpX=rpX, ciX=ciX, rpB=rpB, ciB=ciB,
                valB=valB, nnzbA=int(ciA.size), nnzbX=int(ciX.size), nnzbB=int(ciB.size))


# ------------------------------------------------------------------------------------------------
def read_xml(path: str) -> Problem:
    """LinearProblem XML reader (tfqmrgpu_example_xml_reader.hxx:105-295): NonzerosPerRow | RowStart,
    ColumnIndex, optional Indirection, DataTensor real|complex with ``scale``.  Blocks in the file are
    Fortran/column-major per block, which is why the reference bench uploads with trans 't'
    (bench_tfqmrgpu.cu:153-157); here blocks are returned as stored, ``val[nnzb, dim1, dim2]``.
    The parser (and the matching writers) live in ``formats.py``."""
    from . import formats
    return formats.read_xml_raw(path).to_problem(path)


# ------------------------------------------------------------------------------------------------
def read_multiplication_plan(path: str):
    """``iY iA iX beta`` lines -> (starts u32[nY+1], pairs u32[nPairs,2], nnzY, nnzA, nnzX);
    Y blocks are numbered by order of appearance (bench_tfqmrgpu.cu:456-498)."""
    with open(path) as f:
        head = f.readline().split()
        nnzY, nnzA, nnzX = int(head[1]), int(head[2]), int(head[3])
        data = np.loadtxt(f, dtype=np.int64)
    iY, iA, iX, beta = data.T
    first = np.flatnonzero(np.concatenate([[True], iY[1:] != iY[:-1]]))
    assert np.all(beta[first] == 0) and first.size == nnzY
    starts = np.concatenate([first, [iY.size]]).astype(np.uint32)
    pairs = np.stack([iA, iX], axis=1).astype(np.uint32)
    return starts, pairs, nnzY, nnzA, nnzX, iY[first]


def bsr_from_multiplication_plan(starts, pairs, nnzA_total: int):
    """Reconstruct BSR index arrays (A, X) whose createPlan reproduces ``starts``/``pairs``
    (SURVEY.md 8c 'plan-parity test'): block rows are the connected components of Y blocks sharing an
    A block, numbered by first appearance; block columns are the components of the iY~iX relation.
    Requires the file to be a natural-order createPlan dump (plan_unordered). Returns
    (mb, rpA, ciA, rpX, ciX, a_perm) where a_perm[new_inza] = file iA (identity for natural dumps)."""
    nY = starts.size - 1
    yA = [pairs[starts[y]:starts[y + 1], 0] for y in range(nY)]
    yX = [pairs[starts[y]:starts[y + 1], 1] for y in range(nY)]
    # rows: union-find over Y blocks via shared A blocks
    parent = np.arange(nY)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    ownerA = {}
    for y in range(nY):
        for a in yA[y]:
            a = int(a)
            if a in ownerA:
                ra, rb = find(ownerA[a]), find(y)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
            else:
                ownerA[a] = y
    rootrow = np.array([find(y) for y in range(nY)])
    _, first_idx = np.unique(rootrow, return_index=True)
    order = np.sort(first_idx)
    rowid = {int(rootrow[i]): r for r, i in enumerate(order)}
    rowY = np.array([rowid[int(r)] for r in rootrow])       # X block index == Y block index
    mb = len(order)
    assert np.all(np.diff(rowY) >= 0), "Y blocks must be listed row by row"
    # columns: Y block y and all its X partners share the block column
    parent = np.arange(nY)
    for y in range(nY):
        for x in yX[y]:
            ra, rb = find(int(x)), find(y)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    rootcol = np.array([find(y) for y in range(nY)])
    uniq = {int(c): i for i, c in enumerate(np.unique(rootcol))}
    colY = np.array([uniq[int(c)] for c in rootcol], dtype=np.int32)
    rpX = np.concatenate([[0], np.cumsum(np.bincount(rowY, minlength=mb))]).astype(np.int32)
    ciX = colY
    # A: block iA lives in row(rowY of a Y that uses it), column = row of the X partner
    rowA = np.full(nnzA_total, -1, np.int64)
    colA = np.full(nnzA_total, -1, np.int64)
    for y in range(nY):
        rowA[yA[y]] = rowY[y]
        colA[yA[y]] = rowY[yX[y]]
    # unreferenced A blocks: keep the natural order (monotone rows), give them a column whose X row
    # shares no block column with their own row so that they never create a pair
    cols_of_row = [set(ciX[rpX[r]:rpX[r + 1]].tolist()) for r in range(mb)]
    last_row = 0
    for a in range(nnzA_total):
        if rowA[a] >= 0:
            last_row = rowA[a]
            continue
        # row: same as the previous referenced block (natural order dump)
        nxt = rowA[a + 1:][rowA[a + 1:] >= 0]
        r = int(last_row)
        rowA[a] = r
        used = set(colA[rowA == r].tolist())
        for k in range(mb):
            if k not in used and not (cols_of_row[k] & cols_of_row[r]):
                colA[a] = k
                break
        else:
            raise RuntimeError("no harmless column found for unreferenced A block")
        del nxt
    assert np.all(np.diff(rowA) >= 0), "A blocks must be stored row by row"
    rpA = np.concatenate([[0], np.cumsum(np.bincount(rowA, minlength=mb))]).astype(np.int32)
    ciA = colA.astype(np.int32)
    return mb, rpA, ciA, rpX, ciX
