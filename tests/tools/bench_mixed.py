#!/usr/bin/env python
"""Mixed precision ('m': fp64 refinement around the fp32 tensor-core solver, csrc/mixed.cu) against the plain complex-fp64 solve
('z') on the same operands, through the C-ABI.  Default workload: one GPU's share of BASELINE config 4 on 8 GPUs - the 27-point
block stencil on 32^3 block rows, 32x32 complex fp64 blocks, sigma = 1, 128 right-hand sides (4 block columns), tol 1e-9.
Prints ONE JSON line: per precision the solve time (CUDA events and host clock, K solves after a warm-up), iterations, residual
reached, workspace bytes, setMatrix('A') time; the largest deviation between the two solutions; the true residual of the mixed
solution on sampled block rows, evaluated with numpy in fp64 from the generator's values."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32); ap.add_argument("--lm", type=int, default=32); ap.add_argument("--ln", type=int, default=32)
    ap.add_argument("--ncols", type=int, default=4); ap.add_argument("--sigma", type=float, default=1.0)
    ap.add_argument("--tol", type=float, default=1e-9); ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--only", default="zm")
    ap.add_argument("--variants", nargs="*", default=[], help="mixed plan only: ITER:TOL:FREEZE settings of TFQMRGPU_MIXED_INNER_ITER / "
                    "_INNER_TOL / _FREEZE (read at every solve), one timed solve each on the warm plan; 'x' keeps a default")
    args = ap.parse_args()
    import torch
    from tfqmrgpu_b200 import api, synthetic, problems as P, _lib as L
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n, lm, ln, ncol, tol = args.n, args.lm, args.ln, args.ncols, args.tol
    sp = synthetic.Stencil27(n, lm, ln, ncol, sigma=args.sigma, dtype=np.float64, device=dev)
    out = {"workload": f"stencil27 n={n}^3 block rows, {lm}x{ln} complex fp64 operands, {ncol*ln} right-hand sides, sigma={args.sigma:g}, tol={tol:g}"}
    X = {}
    for prec in args.only:
        h = api.Handle(torch.cuda.current_stream(dev).cuda_stream)
        pl = api.BsrsvPlan(h, sp.mb, sp.rpA, sp.ciA, sp.rpX, sp.ciX, sp.rpB, sp.ciB)
        nbytes = pl.buffer_size_for(lm, ln, prec); pl.set_buffer()
        torch.cuda.synchronize(dev); t0 = time.perf_counter()
        pl.set_matrix("A", None, "n", raw_ptr=sp.valA_host.data_ptr())
        torch.cuda.synchronize(dev); t_a = time.perf_counter() - t0
        pl.set_matrix("B", sp.valB)
        st = pl.solve(tol, 200)                                  # warm-up (graph capture, first touch)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev); t0 = time.perf_counter(); e0.record()
        for _ in range(args.steps):
            st = pl.solve(tol, 200)
        e1.record(); torch.cuda.synchronize(dev); t_host = (time.perf_counter() - t0)/args.steps
        info = pl.info()
        r = {"status": int(st), "iterations": info["iterations"], "residual": info["residuum"], "gflop": info["flops"]*1e-9,
             "ms_per_solve": e0.elapsed_time(e1)/args.steps, "ms_per_solve_host_clock": 1e3*t_host, "workspace_bytes": int(nbytes),
             "setMatrixA_ms": 1e3*t_a, "plan": {k: pl.plan_info()[k] for k in ("use_tc", "use_dmma")}}
        if prec == "m":
            r["mixed"] = pl.mixed_info()
            r["variants"] = []
            for var in args.variants:
                names = ("TFQMRGPU_MIXED_INNER_ITER", "TFQMRGPU_MIXED_INNER_TOL", "TFQMRGPU_MIXED_FREEZE")
                for k, v in zip(names, var.split(":")):
                    os.environ.pop(k, None)
                    if v != "x":
                        os.environ[k] = v
                torch.cuda.synchronize(dev); t0 = time.perf_counter()
                stv = pl.solve(tol, 200)
                tv = time.perf_counter() - t0
                iv, mv = pl.info(), pl.mixed_info()
                r["variants"].append({"setting": var, "ms": 1e3*tv, "status": int(stv), "iterations": iv["iterations"], "passes": mv["passes"],
                                      "residual": iv["residuum"]})
                for k in names:
                    os.environ.pop(k, None)
        X[prec] = pl.get_matrix("X", "n", L.LAYOUT_RRRRIIII).reshape(sp.nnzbX, 2, lm, ln)
        out[prec] = r
        pl.close(); h.close()
        torch.cuda.empty_cache()
    if "z" in X and "m" in X:
        out["max_abs_diff_m_vs_z"] = float(np.abs(X["m"] - X["z"]).max()); out["max_abs_x"] = float(np.abs(X["z"]).max())
        out["speedup_m_over_z"] = out["z"]["ms_per_solve"]/out["m"]["ms_per_solve"]
    # true residual of every solution on sampled block rows (numpy, fp64, from the generator's hashed values)
    rng = np.random.default_rng(1)
    rows = rng.choice(sp.mb, 8, replace=False)
    Bc = np.zeros((sp.mb, ncol, lm, ln), np.complex128)
    brow = np.repeat(np.arange(sp.mb), np.diff(sp.rpB))
    vB = sp.valB.reshape(-1, lm, ln, 2)
    for ib, (r_, c) in enumerate(zip(brow, sp.ciB)):
        Bc[r_, c] = vB[ib, ..., 0] + 1j*vB[ib, ..., 1]
    for prec, Xs in X.items():
        Xc = (Xs[:, 0] + 1j*Xs[:, 1]).reshape(sp.mb, ncol, lm, ln)
        worst = 0.
        for r_ in rows:
            a0, a1 = sp.rpA[r_], sp.rpA[r_ + 1]
            Ab = P.stencil27_values_rows(sp.rpA, sp.ciA, lm, args.sigma, np.arange(a0, a1), dtype=np.float64)
            Ac = Ab[..., 0] + 1j*Ab[..., 1]
            worst = max(worst, float(np.abs(np.einsum("aik,ackj->cij", Ac, Xc[sp.ciA[a0:a1]]) - Bc[r_]).max()))
        out[prec]["max_abs_residual_on_8_sampled_block_rows"] = worst
    print(json.dumps(out))


if __name__ == "__main__":
    main()
