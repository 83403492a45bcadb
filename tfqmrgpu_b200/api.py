"""Host-side mirror of the reference's stepwise C-ABI workflow (tfqmrgpu.h:44, bench_tfqmrgpu.cu:64-217):

    createHandle -> setStream -> createPlan -> bufferSize -> (alloc) -> setBuffer -> setMatrix A,B ->
    solve -> getInfo / getMatrix X -> destroyPlan -> destroyHandle

Every method is a thin call into ``libtfQMRgpu.so``; nothing is computed in Python.  Device memory
for the workspace comes either from ``tfqmrgpuCreateWorkspace`` or from a caller-supplied device
pointer (e.g. a torch uint8 tensor, which lets a multi-GPU driver gather X with NCCL).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class TfqmrError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        self.code, self.line, self.key = L.decode_status(status)
        msg = L.load().tfqmrgpuGetErrorString(status).decode()
        super().__init__(f"{where}: status {status} ({msg})")


def _check(status: int, where: str):
    if status != 0:
        raise TfqmrError(status, where)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


class Handle:
    def __init__(self, stream: int | None = 0):
        self.lib = L.load()
        self.h = C.c_void_p()
        _check(self.lib.tfqmrgpuCreateHandle(C.byref(self.h)), "tfqmrgpuCreateHandle")
        self.set_stream(stream or 0)

    def set_stream(self, stream: int):
        _check(self.lib.tfqmrgpuSetStream(self.h, C.c_void_p(stream)), "tfqmrgpuSetStream")

    def get_stream(self) -> int:
        s = C.c_void_p()
        _check(self.lib.tfqmrgpuGetStream(self.h, C.byref(s)), "tfqmrgpuGetStream")
        return s.value or 0

    def close(self):
        if self.h:
            self.lib.tfqmrgpuDestroyHandle(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BsrsvPlan:
    """One ``A*X == B`` solve.  Mirrors tfqmrgpu_bsrsv_* (tfqmrgpu.h:47-117)."""

    def __init__(self, handle: Handle, mb, rpA, ciA, rpX, ciX, rpB, ciB, index_offset=0, echo=0, check=True):
        self.handle, self.lib = handle, handle.lib
        self.plan = C.c_void_p()
        self._own_buffer = None
        self._keep = None
        a = [_i32(v) for v in (rpA, ciA, rpX, ciX, rpB, ciB)]
        self.nnzbA, self.nnzbX, self.nnzbB, self.mb = len(a[1]), len(a[3]), len(a[5]), mb
        self.status = self.lib.tfqmrgpu_bsrsv_createPlan(
            handle.h, C.byref(self.plan), mb, L.ptr(a[0]), self.nnzbA, L.ptr(a[1]), L.ptr(a[2]), self.nnzbX, L.ptr(a[3]),
            L.ptr(a[4]), self.nnzbB, L.ptr(a[5]), index_offset, echo)
        if check:
            _check(self.status, "tfqmrgpu_bsrsv_createPlan")
        self.lm = self.ln = 0
        self.precision = "z"
        self.buffer = None
        self.buffer_size = 0

    # ---- sizes / buffer -------------------------------------------------------------------------
    def buffer_size_for(self, lm, ln, precision="z", check=True) -> int:
        size = C.c_size_t()
        st = self.lib.tfqmrgpu_bsrsv_bufferSize(self.handle.h, self.plan, lm, lm, ln, ln, precision.encode(), C.byref(size))
        if check:
            _check(st, "tfqmrgpu_bsrsv_bufferSize")
        elif st:
            return -st
        self.lm, self.ln = lm, ln
        self.precision = {"f": "c", "c": "c", "d": "z", "z": "z"}.get(precision.lower(), "z")   # ('m' exchanges doubles)
        self.mixed = precision.lower() == "m"
        self.buffer_size = size.value
        return size.value

    def set_buffer(self, device_ptr: int | None = None, keep_alive=None):
        """Register a device workspace; allocates one with tfqmrgpuCreateWorkspace if none is given."""
        if device_ptr is None:
            buf = C.c_void_p()
            _check(self.lib.tfqmrgpuCreateWorkspace(C.byref(buf), self.buffer_size, b"d"), "tfqmrgpuCreateWorkspace")
            self._own_buffer = buf
            device_ptr = buf.value
        self._keep = keep_alive
        self.buffer = device_ptr
        _check(self.lib.tfqmrgpu_bsrsv_setBuffer(self.handle.h, self.plan, C.c_void_p(device_ptr)), "tfqmrgpu_bsrsv_setBuffer")

    @property
    def dtype(self):
        return np.float64 if self.precision == "z" else np.float32

    # ---- operands ------------------------------------------------------------------------------
    def set_matrix(self, var: str, values, trans="n", layout=L.LAYOUT_RIRIRIRI, precision=None, check=True, raw_ptr=None):
        prec = (precision or self.precision)
        if raw_ptr is None:
            values = np.ascontiguousarray(values, dtype=np.float64 if prec.lower() in "zd" else np.float32)
            raw_ptr = values.ctypes.data
        st = self.lib.tfqmrgpu_bsrsv_setMatrix(self.handle.h, self.plan, var.encode(), C.c_void_p(raw_ptr), prec.encode(),
                                               self.ln, self.lm, trans.encode(), layout)
        if check:
            _check(st, f"tfqmrgpu_bsrsv_setMatrix('{var}')")
        return st

    def get_matrix(self, var="X", trans="n", layout=L.LAYOUT_RIRIRIRI, precision=None, check=True, out=None):
        prec = (precision or self.precision)
        if out is None:
            out = np.zeros(self.nnzbX*self.lm*self.ln*2, dtype=np.float64 if prec.lower() in "zd" else np.float32)
        st = self.lib.tfqmrgpu_bsrsv_getMatrix(self.handle.h, self.plan, var.encode(), L.ptr(out), prec.encode(),
                                               self.ln, self.lm, trans.encode(), layout)
        if check:
            _check(st, f"tfqmrgpu_bsrsv_getMatrix('{var}')")
            return out
        return st, out

    # ---- solve -----------------------------------------------------------------------------------
    def solve(self, threshold=1e-9, max_iterations=200) -> int:
        """Returns the status (0 converged, 9 max iterations, 6 breakdown); other codes raise."""
        st = self.lib.tfqmrgpu_bsrsv_solve(self.handle.h, self.plan, float(threshold), int(max_iterations))
        if st not in (0, L.STATUS_MAX_ITERATIONS, L.STATUS_BREAKDOWN):
            _check(st, "tfqmrgpu_bsrsv_solve")
        return st

    def info(self) -> dict:
        res, it, fl, fla = C.c_double(), C.c_int32(), C.c_double(), C.c_double()
        _check(self.lib.tfqmrgpu_bsrsv_getInfo(self.handle.h, self.plan, C.byref(res), C.byref(it), C.byref(fl), C.byref(fla)),
               "tfqmrgpu_bsrsv_getInfo")
        return dict(residuum=res.value, iterations=it.value, flops=fl.value, flops_all=fla.value)

    # ---- extensions (include/tfqmrgpu_b200_ext.h) ---------------------------------------------------
    def plan_array(self, kind: int) -> np.ndarray:
        n = C.c_size_t()
        _check(self.lib.tfqmrgpux_bsrsv_getPlanArray(self.plan, kind, None, C.byref(n)), "getPlanArray")
        out = np.zeros(max(n.value, 1), np.uint16 if kind == 3 else np.uint32)
        _check(self.lib.tfqmrgpux_bsrsv_getPlanArray(self.plan, kind, L.ptr(out), C.byref(n)), "getPlanArray")
        return out[:n.value]

    def plan_lists(self) -> dict:
        return dict(starts=self.plan_array(0), pairs=self.plan_array(1), subset=self.plan_array(2),
                    colindx=self.plan_array(3), perm=self.plan_array(4), colstart=self.plan_array(5))

    def set_operator(self, fn):
        """User-defined operator (tfqmrgpux_bsrsv_setOperator): ``fn(y_ptr, x_ptr, state_ptr, expect, stream) -> int`` is called
        for every product Y = A*X of solve() with device pointers to X-shaped vectors in storage order
        ``[nnzbX][2][LM][LN]`` (``plan_lists()['perm']`` maps caller block index -> storage index); ``None`` restores the
        built-in block-sparse product."""
        proto = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p)
        if fn is None:
            self._op_keepalive = None
            _check(self.lib.tfqmrgpux_bsrsv_setOperator(self.plan, None, None), "setOperator")
            return

        def trampoline(_ctx, y, x, state, expect, stream):
            try:
                return int(fn(int(y or 0), int(x or 0), int(state or 0), int(expect), int(stream or 0)) or 0)
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return L.UNDOCUMENTED_ERROR
        cb = proto(trampoline)
        self._op_keepalive = cb
        _check(self.lib.tfqmrgpux_bsrsv_setOperator(self.plan, C.cast(cb, C.c_void_p), None), "setOperator")

    def set_preconditioner(self, fn):
        """Right preconditioner (tfqmrgpux_bsrsv_setPreconditioner): ``fn(z_ptr, x_ptr, state_ptr, expect, stream) -> int`` computes
        z = P*x on X-shaped device vectors in storage order, like the callback of set_operator; ``None`` removes it."""
        proto = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p)
        if fn is None:
            self._pc_keepalive = None
            _check(self.lib.tfqmrgpux_bsrsv_setPreconditioner(self.plan, None, None), "setPreconditioner")
            return

        def trampoline(_ctx, z, x, state, expect, stream):
            try:
                return int(fn(int(z or 0), int(x or 0), int(state or 0), int(expect), int(stream or 0)) or 0)
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return L.UNDOCUMENTED_ERROR
        cb = proto(trampoline)
        self._pc_keepalive = cb
        _check(self.lib.tfqmrgpux_bsrsv_setPreconditioner(self.plan, C.cast(cb, C.c_void_p), None), "setPreconditioner")

    def plan_info(self) -> dict:
        info = (C.c_int64*16)()
        _check(self.lib.tfqmrgpux_bsrsv_getPlanInfo(self.plan, info), "getPlanInfo")
        keys = ["nnzbX", "nnzbB", "nnzbA", "nCols", "nPairs", "LM", "LN", "precision", "nTiles", "nUnits", "gmax",
                "nEntries", "mb", "use_tc", "use_dmma", "use_small"]
        return {k: int(info[i]) for i, k in enumerate(keys)}

    def set_v3(self, v3, on_device=False):
        if on_device:
            p = C.c_void_p(int(v3))
        else:
            v3 = np.ascontiguousarray(v3, np.float32)
            assert v3.size == self.nnzbX*2*self.lm*self.ln
            p = L.ptr(v3)
        _check(self.lib.tfqmrgpux_bsrsv_setV3(self.handle.h, self.plan, p, int(on_device)), "setV3")

    def get_v3(self) -> np.ndarray:
        out = np.zeros((self.nnzbX, 2, self.lm, self.ln), np.float32)
        _check(self.lib.tfqmrgpux_bsrsv_getV3(self.handle.h, self.plan, L.ptr(out)), "getV3")
        return out

    def multiply(self, nrep=1):
        _check(self.lib.tfqmrgpux_bsrsv_multiply(self.handle.h, self.plan, nrep), "multiply")

    def get_vector(self, var="Y", trans="n", layout=L.LAYOUT_RIRIRIRI):
        out = np.zeros(self.nnzbX*self.lm*self.ln*2, dtype=self.dtype)
        _check(self.lib.tfqmrgpux_bsrsv_getVector(self.handle.h, self.plan, var.encode(), L.ptr(out), self.precision.encode(),
                                                  trans.encode(), layout), "getVector")
        return out

    def window(self, var: str) -> tuple[int, int]:
        off, ln = C.c_size_t(), C.c_size_t()
        _check(self.lib.tfqmrgpux_bsrsv_getWindow(self.plan, var.encode(), C.byref(off), C.byref(ln)), "getWindow")
        return off.value, ln.value

    def rhs_status(self) -> np.ndarray:
        info = self.plan_info()
        out = np.zeros(info["nCols"]*self.ln, np.int8)
        _check(self.lib.tfqmrgpux_bsrsv_getRhsStatus(self.handle.h, self.plan, L.ptr(out)), "getRhsStatus")
        return out

    def solve_stats(self) -> dict:
        s = (C.c_double*8)()
        _check(self.lib.tfqmrgpux_bsrsv_getSolveStats(self.plan, s), "getSolveStats")
        return dict(probes=int(s[0]), launches=int(s[1]), bodies=int(s[2]), host_ms=s[3], max_bound2=s[4], target_bound2=s[5])

    def set_profiling(self, on=True):
        _check(self.lib.tfqmrgpux_bsrsv_setProfiling(self.plan, int(on)), "setProfiling")

    def solve_profile(self) -> dict:
        s = (C.c_double*8)()
        _check(self.lib.tfqmrgpux_bsrsv_getSolveProfile(self.plan, s), "getSolveProfile")
        return dict(solve_ms=s[0], spmm_ms=s[1], spmm_launches=int(s[2]), iterations=int(s[3]), launches=int(s[4]), probes=int(s[5]))

    # ---- several GPUs (include/tfqmrgpu_b200_ext.h) -------------------------------------------------
    def set_devices(self, n_devices: int, devices=None):
        """One process, several devices: shard the right-hand-side block columns over `n_devices` GPUs (call before
        buffer_size_for).  From then on the ordinary calls drive all devices."""
        d = None if devices is None else np.ascontiguousarray(devices, np.int32)
        _check(self.lib.tfqmrgpux_bsrsv_setDevices(self.handle.h, self.plan, int(n_devices),
                                                   None if d is None else d.ctypes.data_as(C.POINTER(C.c_int32))), "setDevices")

    def get_devices(self) -> list[int]:
        n = C.c_int(0)
        d = np.zeros(64, np.int32)
        _check(self.lib.tfqmrgpux_bsrsv_getDevices(self.plan, C.byref(n), d.ctypes.data_as(C.POINTER(C.c_int32)), 64), "getDevices")
        return [int(v) for v in d[:n.value]] if n.value > 1 else []

    def set_shard_exchange(self, shard: int, n_shards: int, n_rhs_global: int, slots_ptr: int, hook):
        """One process per GPU: register the per-iteration exchange of the convergence monitors that keeps the reference's
        GLOBAL iteration / probe rule.  ``hook(slots_ptr, count, stream) -> int`` must all-gather, in place and on `stream`,
        the ``n_shards`` blocks of 4 doubles starting at ``slots_ptr`` (block `shard` is this rank's).  ``hook=None`` removes it."""
        proto = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
        if hook is None:
            self._exch_keepalive = None
            _check(self.lib.tfqmrgpux_bsrsv_setShardExchange(self.plan, 0, 1, 0, None, None, None), "setShardExchange")
            return

        def trampoline(_ctx, slots, count, stream):
            try:
                return int(hook(int(slots or 0), int(count), int(stream or 0)) or 0)
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return L.UNDOCUMENTED_ERROR
        cb = proto(trampoline)
        self._exch_keepalive = cb
        _check(self.lib.tfqmrgpux_bsrsv_setShardExchange(self.plan, int(shard), int(n_shards), int(n_rhs_global),
                                                         C.c_void_p(int(slots_ptr)), C.cast(cb, C.c_void_p), None), "setShardExchange")

    def matrix_part_info(self, part: int, n_parts: int) -> dict:
        """Row range `part` of `n_parts` of A: byte windows of its converted blocks and row scales inside the workspace."""
        info = (C.c_int64*6)()
        _check(self.lib.tfqmrgpux_bsrsv_getMatrixPartInfo(self.plan, int(part), int(n_parts), info), "getMatrixPartInfo")
        return dict(off=int(info[0]), length=int(info[1]), scale_off=int(info[2]), scale_length=int(info[3]), block0=int(info[4]), nblocks=int(info[5]))

    def set_matrix_part(self, raw_ptr: int, part: int, n_parts: int, trans="n", layout=L.LAYOUT_RIRIRIRI) -> dict:
        """Upload and convert row range `part` of A; `raw_ptr` points to the FIRST BLOCK OF THAT RANGE in host memory."""
        info = (C.c_int64*6)()
        _check(self.lib.tfqmrgpux_bsrsv_setMatrixPart(self.handle.h, self.plan, C.c_void_p(int(raw_ptr)), self.precision.encode(),
                                                     trans.encode(), layout, int(part), int(n_parts), info), "setMatrixPart")
        return dict(off=int(info[0]), length=int(info[1]), scale_off=int(info[2]), scale_length=int(info[3]), block0=int(info[4]), nblocks=int(info[5]))

    def set_early_freeze(self, on=True):
        """Opt-in (not in the reference): a right-hand side whose true residual passes a probe keeps its X (status 2)."""
        _check(self.lib.tfqmrgpux_bsrsv_setEarlyFreeze(self.plan, int(bool(on))), "setEarlyFreeze")

    def set_initial_guess(self, on=True, check=True):
        """Mixed-precision plans ('m') only: start the solve from the uploaded X / the previous solution instead of zero."""
        st = self.lib.tfqmrgpux_bsrsv_setInitialGuess(self.plan, int(bool(on)))
        if check:
            _check(st, "setInitialGuess")
        return st

    def mixed_info(self) -> dict:
        v = (C.c_double*8)()
        _check(self.lib.tfqmrgpux_bsrsv_getMixedInfo(self.plan, v), "getMixedInfo")
        return dict(mixed=bool(v[0]), passes=int(v[1]), inner_iterations=int(v[2]), inner_bytes=int(v[3]),
                    inner_product={0: "simt", 1: "tcgen05 direct", 2: "tcgen05 planar"}[int(v[4])], dmma=bool(v[5]))

    def set_rhs_trivial(self):
        """B := unit blocks (the reference's rhs_trivial right-hand sides) instead of set_matrix('B', ...)."""
        _check(self.lib.tfqmrgpux_bsrsv_setRhsTrivial(self.handle.h, self.plan), "setRhsTrivial")

    def set_shard_hints(self, tile_blocks: int, max_cols_per_row: int = 0):
        _check(self.lib.tfqmrgpux_bsrsv_setShardHints(self.plan, int(tile_blocks), int(max_cols_per_row)), "setShardHints")

    def tile_blocks(self) -> int:
        v = C.c_int64(0)
        _check(self.lib.tfqmrgpux_bsrsv_getTileBlocks(self.plan, C.byref(v)), "getTileBlocks")
        return int(v.value)

    def close(self):
        if self.plan:
            self.lib.tfqmrgpu_bsrsv_destroyPlan(self.handle.h, self.plan)
            self.plan = C.c_void_p()
        if self._own_buffer:
            self.lib.tfqmrgpuDestroyWorkspace(self._own_buffer)
            self._own_buffer = None
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tile_blocks_for(nnzbX: int, block_bytes: int) -> int:
    """X blocks per vector tile that bufferSize chooses for a plan with nnzbX blocks of block_bytes on the current device."""
    v = C.c_int64(0)
    _check(L.load().tfqmrgpux_tileBlocksFor(int(nnzbX), int(block_bytes), C.byref(v)), "tileBlocksFor")
    return int(v.value)


def bsrsv(precision, mb, lm, ln, rpA, ciA, valA, transA, rpX, ciX, transX, rpB, ciB, valB, transB,
          max_iterations=200, threshold=1e-9, index_offset=0, echo=0):
    """The quick starter ``tfqmrgpu_bsrsv_z`` / ``_c`` (tfqmrgpu.h:138-156) on host arrays.
    Returns (status, X host array [nnzbX, lm, ln, 2], iterations, residual)."""
    lib = L.load()
    dt = np.float64 if precision == "z" else np.float32
    a = [_i32(v) for v in (rpA, ciA, rpX, ciX, rpB, ciB)]
    valA = np.ascontiguousarray(valA, dt); valB = np.ascontiguousarray(valB, dt)
    X = np.zeros((len(a[3]), lm, ln, 2), dt)
    it = C.c_int32(max_iterations)
    res = C.c_float(threshold)
    fn = lib.tfqmrgpu_bsrsv_z if precision == "z" else lib.tfqmrgpu_bsrsv_c
    st = fn(mb, lm, ln, L.ptr(a[0]), len(a[1]), L.ptr(a[1]), L.ptr(valA), transA.encode(),
            L.ptr(a[2]), len(a[3]), L.ptr(a[3]), L.ptr(X), transX.encode(),
            L.ptr(a[4]), len(a[5]), L.ptr(a[5]), L.ptr(valB), transB.encode(),
            C.byref(it), C.byref(res), index_offset, echo)
    return st, X, it.value, res.value
