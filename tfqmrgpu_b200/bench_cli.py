"""Command line with the argument conventions and result lines of the reference's ``bench_tfqmrgpu``
(``tfQMRgpu/source/bench_tfqmrgpu.cu:442-590``), running THIS library:

    python -m tfqmrgpu_b200.bench_cli tfQMR    <problem file> [z|c] [#repetitions] [#max iterations]
    python -m tfqmrgpu_b200.bench_cli multiply <plan file> [f|d]   [#repetitions] [#samples] [lm] [ln]

``tfQMR``   reads a LinearProblem XML (file name contains "xml") or a legacy text dump, solves A*X == B through the
            C-ABI with the reference's transposition flags ('t' for A, B and X: the files hold Fortran-ordered blocks,
            bench_tfqmrgpu.cu:153-173) and prints the ``# GPU maxdev``, ``# GPU converged to`` and ``# GPU performed`` lines.
``multiply`` reads a multiplication plan (``iY iA iX beta`` lines), rebuilds the BSR patterns that produce it, fills A and X
            with the harness's cos/sin values (:277-285), times Y = A*X (the product kernel this library selects) and
            checks Y against a numpy evaluation of the same pair list with the harness's pass bar (maxdev <= 1e-4, :414).

The unmodified reference harness itself also runs against this library for the ``tfQMR`` task (oracle/build_ref.sh callers);
its ``multiply`` task calls kernels compiled into the harness, which is why this command exists.
"""
from __future__ import annotations

import sys
import time

import numpy as np

from . import _lib as L, api, formats, problems as P


def fill_cos_sin(nblocks: int, lm: int, ln: int, dtype) -> np.ndarray:
    """[nblocks][2][lm][ln]: cos | sin of ((m*LM + i)*LN + j), evaluated in double (bench_tfqmrgpu.cu:277-285)."""
    m = np.arange(nblocks, dtype=np.float64)[:, None, None]
    i = np.arange(lm, dtype=np.float64)[None, :, None]
    j = np.arange(ln, dtype=np.float64)[None, None, :]
    arg = (m*lm + i)*ln + j
    return np.stack([np.cos(arg), np.sin(arg)], axis=1).astype(dtype)


def _pairs_product(A, X, starts, pairs):
    """Y[y] = sum over the pairs of y of A[iA] (x) X[iX] with A stored [k][i] (the harness's convention), in double."""
    Ac = A[:, 0].astype(np.float64) + 1j*A[:, 1]
    Xc = X[:, 0].astype(np.float64) + 1j*X[:, 1]
    prod = np.einsum("pki,pkj->pij", Ac[pairs[:, 0]], Xc[pairs[:, 1]])
    Y = np.add.reduceat(prod, starts[:-1].astype(np.int64), axis=0)
    Y[np.diff(starts.astype(np.int64)) == 0] = 0
    return Y


def run_multiply(argv) -> int:
    import torch
    fnm = argv[2] if len(argv) > 2 else "plan"
    fF = (argv[3] if len(argv) > 3 else "f")[0]
    nrep = int(argv[4]) if len(argv) > 4 else 1
    nsamp = int(argv[5]) if len(argv) > 5 else 1
    lm = int(argv[6]) if len(argv) > 6 else 16
    ln = int(argv[7]) if len(argv) > 7 else lm
    prec, dt, flop_char = ("z", np.float64, "F") if fF.lower() in "dz" else ("c", np.float32, "f")
    try:
        starts, pairs, nnzY, nnzA, nnzX, _ = P.read_multiplication_plan(fnm)
    except OSError:
        print(f"{argv[0]}: error: did not find file")
        return 1
    print(f"# nnz Y {nnzY}\n# nnz A {nnzA}\n# nnz X {nnzX}")
    print(f"# found {starts.size - 1} result elements")
    print(f"# found {pairs.shape[0]} operations")
    mb, rpA, ciA, rpX, ciX = P.bsr_from_multiplication_plan(starts, pairs, nnzA)
    h = api.Handle()
    pl = api.BsrsvPlan(h, mb, rpA, ciA, rpX, ciX, rpX, ciX)
    lists = pl.plan_lists()
    if not (np.array_equal(lists["starts"], starts) and np.array_equal(lists["pairs"].reshape(-1, 2), pairs)):
        print("# Warning! the plan rebuilt from the file differs from the file")
    pl.buffer_size_for(lm, ln, prec); pl.set_buffer()
    info = pl.plan_info()
    print(f"\n# bench_multi<{lm},{ln}> on GPU !!!!")
    print(f"# Execute {nrep} repetitions, sample {nsamp} times.")
    A = fill_cos_sin(nnzA, lm, lm, dt); X = fill_cos_sin(nnzX, lm, ln, dt)
    # 'A' uploaded with 't' in RRRRIIII lands untransposed in the internal [k][i] storage, as the harness fills it
    pl.set_matrix("A", A, "t", L.LAYOUT_RRRRIIII); pl.set_matrix("X", X, "n", L.LAYOUT_RRRRIIII)
    kernel = "dmma" if info["use_dmma"] else ("tcgen05" if info["use_tc"] else ("simt-small" if info.get("use_small") else "simt"))
    print(f"# product kernel: {kernel}, {info['nUnits']} units of up to {info['gmax']} block columns")
    pl.multiply(1); torch.cuda.synchronize()
    times = []
    for _ in range(nsamp):
        t0 = time.perf_counter()
        pl.multiply(nrep)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = np.array(times); tsum, tavg = float(t.sum()), float(t.mean())
    trms = float(np.sqrt(max(0.0, float((t*t).mean()) - tavg*tavg)))
    print("# GPU needed %.3f seconds, %.6f +/- %.6f sec per sample, %.1f%% dev" % (tsum, tavg, trms, trms*100./max(tavg, 1e-30)))
    Y = pl.get_vector("Y", "n", L.LAYOUT_RRRRIIII).reshape(nnzX, 2, lm, ln)
    pl.close(); h.close()
    t0 = time.perf_counter()
    Yr = _pairs_product(A, X, starts, pairs)
    dev = np.abs(np.stack([Yr.real, Yr.imag], axis=1) - Y.astype(np.float64))
    maxdev, avgdev = float(dev.max()), float(dev.mean())
    print("# GPU maxdev %g avgdev %g" % (maxdev, avgdev))
    if maxdev > 1e-4:
        print("# Warning! GPU result has large deviations (%g) for blockDim=%d x %d" % (maxdev, lm, ln))
        return 0
    print("# CPU result checking with %d threads took %.3f sec" % (1, time.perf_counter() - t0))
    nflop = float(pairs.shape[0])*(8.*lm)*(lm*ln)*nrep*nsamp
    print("# GPU performed %.3f T%clop in %.3f seconds" % (nflop*1e-12, flop_char, tsum))
    print("# GPU performance (lm,ln,tune)=(%3d,%3d,%d) is  %.1f G%clop/sec" % (lm, ln, 0, nflop*1e-9/max(tsum, 1e-30), flop_char))
    return 0


def run_tfqmr(argv) -> int:
    fnm = argv[2] if len(argv) > 2 else "problem"
    prec = (argv[3] if len(argv) > 3 else "z")[0].lower()
    nrep = int(argv[4]) if len(argv) > 4 else 1
    maxit = int(argv[5]) if len(argv) > 5 else 2000
    print(f"\n# read file '{fnm}' as input.")
    prob = P.read_xml(fnm) if "xml" in fnm else formats.read_legacy(fnm)
    print("# found tolerance= %g" % prob.tolerance)
    print(f"# Execute {nrep} repetitions with max. {maxit} iterations")
    print(f"# requested precision= '{prec}' for LM= {prob.lm}, LN= {prob.ln}")
    if prec not in "zc":     # 'm' (mixed) is declared by the reference but not implemented (tfqmrgpu.cu:42-44)
        print(api.TfqmrError(L.PRECISION_MISSMATCH + L.CODE_CHAR*ord(prec), "precision"))
        return 1
    dt = np.float64 if "z" == prec else np.float32
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
    ref = prob.X.val if prob.X.val is not None else np.zeros((prob.X.nnzb, prob.ln, prob.lm), np.complex128)
    for _ in range(max(nrep, 1)):
        print("\n# nnzb for A=%d, X=%d, B=%d" % (prob.A.nnzb, prob.X.nnzb, prob.B.nnzb))
        h = api.Handle()
        pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
        print(f"# compute the GPU memory requirements for precision='{prec}' LM={prob.lm} LN={prob.ln}")
        nbytes = pl.buffer_size_for(prob.lm, prob.ln, prec); pl.set_buffer()
        print("# use %.6f GByte GPU memory" % (nbytes*1e-9))
        pl.set_matrix("A", vA, "t"); pl.set_matrix("B", vB, "t")
        t0 = time.perf_counter()
        st = pl.solve(prob.tolerance, maxit)
        solver_time = time.perf_counter() - t0
        if st not in (0, 6, 9):
            print(api.TfqmrError(st, "solve"))
        res = pl.get_matrix("X", "t").astype(np.float64).reshape(-1)
        refflat = np.empty(res.size, np.float64)
        refflat[0::2] = np.asarray(ref).real.reshape(-1); refflat[1::2] = np.asarray(ref).imag.reshape(-1)
        dev = np.abs(res - refflat)
        nz = refflat != 0
        maxrel = float((dev[nz]/refflat[nz]).max()) if nz.any() else 0.0
        print("# GPU maxdev %g avgdev %g maxrel %g" % (float(dev.max()), float(dev.mean()), maxrel))
        info = pl.info()
        c = "F" if "z" == prec else "f"
        tflop = 1e-12*info["flops"]
        if dev.max() < 1e-5:          # the harness reports these two lines only when X matches the file's X (:192-205)
            print("# GPU converged to %.1e in %d iterations" % (info["residuum"], info["iterations"]))
            print("# GPU performed %.3f T%clop in %.3f seconds = %.3f T%clop/s" % (tflop, c, solver_time, tflop/max(solver_time, 1e-6), c))
        # always (the generator's files carry no reference X, so the harness stays silent about the solve there)
        print("# solve: status %d, %d iterations, residual %.1e, %.6f GFlop in %.3f ms"
              % (st, info["iterations"], info["residuum"], 1e-9*info["flops"], 1e3*solver_time))
        pl.close(); h.close()
    return 0


def main(argv=None) -> int:
    argv = list(sys.argv if argv is None else argv)
    if len(argv) < 2:
        print("Usage:  %s  [tfQMR/multiply]  [file]  [float/double]  [#repetitions]  [#iterations]  [#blocksize]" % argv[0])
        return 1
    if "m" == argv[1][0]:
        return run_multiply(argv)
    return run_tfqmr(argv)


if __name__ == "__main__":
    sys.exit(main())
