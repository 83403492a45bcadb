#!/usr/bin/env python
"""dev: the right-preconditioner slot with P = identity / a diagonal scaling, applied by cudaMemcpyAsync, by torch on the legacy
stream, by torch with a device synchronisation, and on a torch side stream (status, iterations, residual, error against the plain solve)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from tfqmrgpu_b200 import api, problems as P, _lib as L

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]


class DevArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


for (lm, ln, prec) in [(16, 16, "z"), (32, 32, "z"), (32, 32, "c")]:
    tol = 1e-9 if prec == "z" else 1e-4
    prob = P.random_system(12, lm, ln, seed=lm*10 + ln, unsorted=True)
    dt, ts, tdt = (np.float64, "<f8", torch.float64) if prec == "z" else (np.float32, "<f4", torch.float32)
    vA = P.interleave(prob.A.val, dt); vB = P.interleave(prob.B.val, dt)
    Xbase = None
    for mode in ("none", "memcpy", "torch_identity", "torch_identity_sync", "torch_identity_sidestream", "per_row", "per_row_sync", "per_row_sidestream"):
        side = torch.cuda.Stream() if mode.endswith("sidestream") else None
        h = api.Handle(side.cuda_stream if side else 0)
        pl = api.BsrsvPlan(h, prob.mb, prob.A.rowptr, prob.A.colind, prob.X.rowptr, prob.X.colind, prob.B.rowptr, prob.B.colind)
        pl.buffer_size_for(lm, ln, prec); pl.set_buffer()
        shape = (pl.nnzbX, 2, lm, ln)
        nbytes = int(np.prod(shape))*(8 if prec == "z" else 4)
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        d = (0.5 + torch.rand((pl.nnzbX, 1, lm, 1), generator=g, device="cuda", dtype=tdt)) if mode.startswith("per_row") else torch.ones((1, 1, 1, 1), device="cuda", dtype=tdt)
        torch.cuda.synchronize()

        def pc(z_ptr, x_ptr, state_ptr, expect, stream, mode=mode, d=d):
            if mode == "memcpy":
                return rt.cudaMemcpyAsync(z_ptr, x_ptr, nbytes, 3, stream)
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                x = torch.as_tensor(DevArray(x_ptr, shape, ts), device="cuda")
                z = torch.as_tensor(DevArray(z_ptr, shape, ts), device="cuda")
                torch.mul(x, d, out=z)
            if mode.endswith("_sync"):
                torch.cuda.synchronize()
            return 0
        if mode != "none":
            pl.set_preconditioner(pc)
        pl.set_matrix("A", vA, "n"); pl.set_matrix("B", vB, "n")
        st = pl.solve(tol, 200)
        info = pl.info(); stats = pl.solve_stats()
        X = pl.get_matrix("X", "n", L.LAYOUT_RRRRIIII)
        if Xbase is None:
            Xbase = X
        print(f"{lm}x{ln} {prec} {mode:26s} status {st} iterations {info['iterations']:3d} residual {info['residuum']:.3e} probes {stats['probes']:.0f} "
              f"err {np.abs(X - Xbase).max()/np.abs(Xbase).max():.2e}", flush=True)
        pl.close(); h.close()
