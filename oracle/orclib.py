"""ctypes bindings for the TEST-ONLY oracle (oracle/liboracle.so) and, when it was built in this
container, the unmodified reference (oracle/_ref/libtfqmr_ref_cpu.so / _gpu.so + harness).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_CPU_SO = os.path.join(ROOT, "oracle", "_ref", "libtfqmr_ref_cpu.so")
REF_GPU_SO = os.path.join(ROOT, "oracle", "_ref", "libtfqmr_ref_gpu.so")

MODE_GPU, MODE_CPUREF = 0, 1
LAYOUT_RRRRIIII, LAYOUT_RRIIRRII, LAYOUT_RIRIRIRI = 0x0f, 0x33, 0x55

i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")


class OrcPlan(C.Structure):
    _fields_ = [("mb", C.c_int32), ("nnzbA", C.c_int32), ("nnzbX", C.c_int32), ("nnzbB", C.c_int32),
                ("nCols", C.c_uint32), ("nPairs", C.c_uint64),
                ("starts", C.POINTER(C.c_uint32)), ("pairs", C.POINTER(C.c_uint32)),
                ("subset", C.POINTER(C.c_uint32)), ("colindx", C.POINTER(C.c_uint16))]


class OrcInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations_needed", C.c_int32), ("iterations_run", C.c_int32),
                ("probes", C.c_int32), ("residuum_reached", C.c_double), ("flops_performed", C.c_double),
                ("last_max_bound2", C.c_double), ("last_target_bound2", C.c_double)]


class quiet_stdout:
    """Silence the C-level stdout chatter of the reference library (its debug printf is always on)."""
    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1); self.null = os.open(os.devnull, os.O_WRONLY); os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1); os.close(self.saved); os.close(self.null)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(ORACLE_SO)
        lib.orc_create_plan.restype = C.c_int
        lib.orc_create_plan.argtypes = [C.c_int, i32p, C.c_int, i32p, i32p, C.c_int, i32p, i32p, C.c_int, i32p,
                                        C.c_int, C.POINTER(C.POINTER(OrcPlan))]
        lib.orc_destroy_plan.argtypes = [C.POINTER(OrcPlan)]
        lib.orc_plan_array.restype = C.c_uint64
        lib.orc_plan_array.argtypes = [C.POINTER(OrcPlan), C.c_int, C.c_void_p]
        lib.orc_ref_buffer_size.restype = C.c_uint64
        lib.orc_ref_buffer_size.argtypes = [C.POINTER(OrcPlan), C.c_int, C.c_int, C.c_int]
        lib.orc_v3_glibc.argtypes = [C.c_void_p, C.c_size_t]
        for sfx, ct in (("d", C.c_double), ("f", C.c_float)):
            getattr(lib, f"orc_import_blocks_{sfx}").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                                                 C.c_int, C.c_char, C.c_char]
            getattr(lib, f"orc_export_blocks_{sfx}").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                                                 C.c_int, C.c_char]
            getattr(lib, f"orc_fill_cos_sin_{sfx}").argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
            getattr(lib, f"orc_multiply_{sfx}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u32p, u32p, C.c_uint32,
                                                            C.c_int, C.c_int, C.c_int, C.c_int]
        for name in ("orc_solve_z", "orc_solve_c"):
            f = getattr(lib, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(OrcPlan), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_double, C.c_int, C.c_int, C.POINTER(OrcInfo), C.c_void_p]
        _oracle = lib
    return _oracle


class OraclePlan:
    """createPlan restatement; holds the reference-format lists as numpy arrays."""

    def __init__(self, mb, rpA, ciA, rpX, ciX, rpB, ciB, index_offset=0):
        lib = oracle()
        self._p = C.POINTER(OrcPlan)()
        a = [np.ascontiguousarray(v, np.int32) for v in (rpA, ciA, rpX, ciX, rpB, ciB)]
        # ctypes ndpointer rejects zero-length views of None; keep 1 element minimum
        a = [v if v.size else np.zeros(1, np.int32) for v in a]
        self.status = lib.orc_create_plan(mb, a[0], len(ciA), a[1], a[2], len(ciX), a[3], a[4], len(ciB), a[5],
                                          index_offset, C.byref(self._p))
        if self.status == 0:
            p = self._p.contents
            self.nnzbX, self.nnzbB, self.nnzbA = p.nnzbX, p.nnzbB, p.nnzbA
            self.nCols, self.nPairs = p.nCols, p.nPairs
            self.starts = self._arr(0, np.uint32)
            self.pairs = self._arr(1, np.uint32)
            self.subset = self._arr(2, np.uint32)
            self.colindx = self._arr(3, np.uint16)

    def _arr(self, kind, dt):
        lib = oracle()
        n = lib.orc_plan_array(self._p, kind, None)
        out = np.zeros(max(int(n), 1), dt)
        lib.orc_plan_array(self._p, kind, out.ctypes.data)
        return out[:int(n)]

    def ref_buffer_size(self, lm, ln, is_double):
        return int(oracle().orc_ref_buffer_size(self._p, lm, ln, int(is_double)))

    def __del__(self):
        try:
            if self._p:
                oracle().orc_destroy_plan(self._p)
        except Exception:
            pass


def _sfx(dtype):
    return "d" if np.dtype(dtype) == np.float64 else "f"


def import_blocks(host, nnzb, rows, cols, layout=LAYOUT_RIRIRIRI, trans="n", var="X"):
    host = np.ascontiguousarray(host)
    out = np.zeros((nnzb, 2, rows, cols), host.dtype)
    if nnzb:
        st = getattr(oracle(), f"orc_import_blocks_{_sfx(host.dtype)}")(
            out.ctypes.data, host.ctypes.data, nnzb, rows, cols, layout, trans.encode(), var.encode())
        assert st == 0, st
    return out


def export_blocks(internal, rows, cols, layout=LAYOUT_RIRIRIRI, trans="n"):
    internal = np.ascontiguousarray(internal)
    nnzb = internal.shape[0]
    out = np.zeros(nnzb*2*rows*cols, internal.dtype)
    st = getattr(oracle(), f"orc_export_blocks_{_sfx(internal.dtype)}")(
        out.ctypes.data, internal.ctypes.data, nnzb, rows, cols, layout, trans.encode())
    assert st == 0, st
    return out


def fill_cos_sin(nmat, lm, ln, dtype):
    out = np.zeros((nmat, 2, lm, ln), dtype)
    getattr(oracle(), f"orc_fill_cos_sin_{_sfx(dtype)}")(out.ctypes.data, nmat, lm, ln)
    return out


def multiply(A, X, starts, pairs, lm, ln, mode=MODE_GPU, nthreads=1):
    A = np.ascontiguousarray(A); X = np.ascontiguousarray(X)
    nY = starts.size - 1
    Y = np.zeros((nY, 2, lm, ln), A.dtype)
    getattr(oracle(), f"orc_multiply_{_sfx(A.dtype)}")(
        Y.ctypes.data, A.ctypes.data, X.ctypes.data, np.ascontiguousarray(starts, np.uint32),
        np.ascontiguousarray(pairs, np.uint32).reshape(-1), nY, lm, ln, mode, nthreads)
    return Y


def v3_glibc(n):
    out = np.zeros(n, np.float32)
    oracle().orc_v3_glibc(out.ctypes.data, n)
    return out


def solve(plan: OraclePlan, lm, ln, A_int, B_int, v3, tol, maxit, mode=MODE_GPU):
    """A_int [nnzbA,2,lm,lm] (stored [k][i]), B_int [nnzbB,2,lm,ln], v3 float32 [nnzbX,2,lm,ln]."""
    A_int = np.ascontiguousarray(A_int); B_int = np.ascontiguousarray(B_int, A_int.dtype)
    v3 = np.ascontiguousarray(v3, np.float32)
    X = np.zeros((plan.nnzbX, 2, lm, ln), A_int.dtype)
    info = OrcInfo()
    status = np.zeros(plan.nCols*ln, np.int8)
    fn = oracle().orc_solve_z if A_int.dtype == np.float64 else oracle().orc_solve_c
    st = fn(plan._p, lm, ln, A_int.ctypes.data, B_int.ctypes.data, v3.ctypes.data, X.ctypes.data,
            float(tol), int(maxit), mode, C.byref(info), status.ctypes.data)
    return dict(status=st, X=X, iterations=info.iterations_needed, iterations_run=info.iterations_run,
                probes=info.probes, residuum=info.residuum_reached, flops=info.flops_performed,
                rhs_status=status)


# ------------------------------------------------------------------------------------------------
# the unmodified reference (when built): same C-ABI as the product + the harness peeks
class RefLib:
    def __init__(self, path):
        self.lib = lib = C.CDLL(path)
        self.is_cpu = bool(lib.refh_is_cpu_build())
        lib.refh_aligned_alloc.restype = C.c_void_p
        lib.refh_aligned_alloc.argtypes = [C.c_size_t]
        lib.refh_aligned_free.argtypes = [C.c_void_p]
        lib.refh_plan_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.refh_plan_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.refh_v3_count.restype = C.c_uint64
        lib.refh_v3_count.argtypes = [C.c_void_p]
        lib.refh_v3_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.refh_x_copy.argtypes = [C.c_void_p, C.c_void_p]
        lib.tfqmrgpuCreateHandle.argtypes = [C.POINTER(C.c_void_p)]
        lib.tfqmrgpuDestroyHandle.argtypes = [C.c_void_p]
        lib.tfqmrgpuSetStream.argtypes = [C.c_void_p, C.c_void_p if not self.is_cpu else C.c_int]
        lib.tfqmrgpu_bsrsv_createPlan.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, i32p, C.c_int, i32p,
                                                  i32p, C.c_int, i32p, i32p, C.c_int, i32p, C.c_int, C.c_int]
        lib.tfqmrgpu_bsrsv_destroyPlan.argtypes = [C.c_void_p, C.c_void_p]
        lib.tfqmrgpu_bsrsv_bufferSize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char,
                                                  C.POINTER(C.c_size_t)]
        lib.tfqmrgpu_bsrsv_setBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.tfqmrgpu_bsrsv_setMatrix.argtypes = [C.c_void_p, C.c_void_p, C.c_char, C.c_void_p, C.c_char, C.c_int, C.c_int,
                                                 C.c_char, C.c_int]
        lib.tfqmrgpu_bsrsv_getMatrix.argtypes = lib.tfqmrgpu_bsrsv_setMatrix.argtypes
        lib.tfqmrgpu_bsrsv_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int]
        lib.tfqmrgpu_bsrsv_getInfo.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                               C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.tfqmrgpuCreateWorkspace.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_char]
        lib.tfqmrgpuDestroyWorkspace.argtypes = [C.c_void_p]

    def plan_only(self, mb, rpA, ciA, rpX, ciX, rpB, ciB, index_offset=0):
        """createPlan through the reference, return (status, dict of lists)."""
        lib = self.lib
        h = C.c_void_p(); lib.tfqmrgpuCreateHandle(C.byref(h))
        plan = C.c_void_p()
        a = [np.ascontiguousarray(v, np.int32) for v in (rpA, ciA, rpX, ciX, rpB, ciB)]
        a = [v if v.size else np.zeros(1, np.int32) for v in a]
        st = lib.tfqmrgpu_bsrsv_createPlan(h, C.byref(plan), mb, a[0], len(ciA), a[1], a[2], len(ciX), a[3],
                                           a[4], len(ciB), a[5], index_offset, 0)
        out = None
        if st == 0:
            out = self._lists(plan)
            lib.tfqmrgpu_bsrsv_destroyPlan(h, plan)
        lib.tfqmrgpuDestroyHandle(h)
        return st, out

    def _lists(self, plan):
        cnt = (C.c_uint64*8)()
        self.lib.refh_plan_counts(plan, cnt)
        nX, nB, nC, nP = int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3])
        starts = np.zeros(nX + 1, np.uint32); pairs = np.zeros(max(2*nP, 1), np.uint32)
        subset = np.zeros(max(nB, 1), np.uint32); colindx = np.zeros(nX, np.uint16)
        self.lib.refh_plan_copy(plan, 0, starts.ctypes.data)
        self.lib.refh_plan_copy(plan, 1, pairs.ctypes.data)
        self.lib.refh_plan_copy(plan, 2, subset.ctypes.data)
        self.lib.refh_plan_copy(plan, 3, colindx.ctypes.data)
        return dict(starts=starts, pairs=pairs[:2*nP], subset=subset[:nB], colindx=colindx, nCols=nC, nPairs=nP)

    def solve(self, mb, lm, ln, rpA, ciA, valA, rpX, ciX, rpB, ciB, valB, tol, maxit, precision="z",
              v3=None, transA="n", index_offset=0, trans_b="n"):
        """Stepwise API run through the unmodified reference. valA/valB host layout RIRIRIRI.
        For the CPU build transA is flipped by the caller-independent rule of SURVEY 8c-3 here."""
        lib = self.lib
        dt = np.float64 if precision == "z" else np.float32
        h = C.c_void_p(); assert lib.tfqmrgpuCreateHandle(C.byref(h)) == 0
        plan = C.c_void_p()
        a = [np.ascontiguousarray(v, np.int32) for v in (rpA, ciA, rpX, ciX, rpB, ciB)]
        t_plan0 = time.perf_counter()
        st = lib.tfqmrgpu_bsrsv_createPlan(h, C.byref(plan), mb, a[0], len(ciA), a[1], a[2], len(ciX), a[3],
                                           a[4], len(ciB), a[5], index_offset, 0)
        t_plan = time.perf_counter() - t_plan0
        assert st == 0, st
        size = C.c_size_t()
        st = lib.tfqmrgpu_bsrsv_bufferSize(h, plan, lm, lm, ln, ln, precision.encode(), C.byref(size))
        assert st == 0, st
        if self.is_cpu:
            buf = C.c_void_p(lib.refh_aligned_alloc(size.value))
        else:
            buf = C.c_void_p(); assert lib.tfqmrgpuCreateWorkspace(C.byref(buf), size.value, b"d") == 0
        assert lib.tfqmrgpu_bsrsv_setBuffer(h, plan, buf) == 0
        nX = len(ciX)
        if v3 is not None:
            v3 = np.ascontiguousarray(v3, np.float32).reshape(-1)
            assert v3.size == lib.refh_v3_count(plan)
            assert lib.refh_v3_copy(plan, v3.ctypes.data, 0) == 0
        v3_used = np.zeros(int(lib.refh_v3_count(plan)), np.float32)
        lib.refh_v3_copy(plan, v3_used.ctypes.data, 1)
        tA = transA
        if self.is_cpu:  # the CPU path multiplies with block-transposed A (blocksparse.hxx:167)
            tA = {"n": "t", "t": "n"}[transA]
        vA = np.ascontiguousarray(valA, dt); vB = np.ascontiguousarray(valB, dt)
        assert lib.tfqmrgpu_bsrsv_setMatrix(h, plan, b"A", vA.ctypes.data, precision.encode(), lm, lm, tA.encode(), 0x55) == 0
        assert lib.tfqmrgpu_bsrsv_setMatrix(h, plan, b"B", vB.ctypes.data, precision.encode(), ln, lm, trans_b.encode(), 0x55) == 0
        lists = self._lists(plan)
        t0 = time.perf_counter()
        status = lib.tfqmrgpu_bsrsv_solve(h, plan, float(tol), int(maxit))
        t_solve = time.perf_counter() - t0
        res = C.c_double(); it = C.c_int32(); fl = C.c_double(); fla = C.c_double()
        lib.tfqmrgpu_bsrsv_getInfo(h, plan, C.byref(res), C.byref(it), C.byref(fl), C.byref(fla))
        Xint = np.zeros((nX, 2, lm, ln), dt)
        lib.refh_x_copy(plan, Xint.ctypes.data)
        lib.tfqmrgpu_bsrsv_destroyPlan(h, plan)
        if self.is_cpu:
            lib.refh_aligned_free(buf)
        else:
            lib.tfqmrgpuDestroyWorkspace(buf)
        lib.tfqmrgpuDestroyHandle(h)
        return dict(status=status, X=Xint, iterations=it.value, residuum=res.value, flops=fl.value,
                    buffer_size=size.value, v3=v3_used, lists=lists, t_solve=t_solve, t_plan=t_plan)


_ref_cpu = None


def ref_cpu():
    global _ref_cpu
    if _ref_cpu is None and os.path.exists(REF_CPU_SO):
        _ref_cpu = RefLib(REF_CPU_SO)
    return _ref_cpu


def ref_gpu():
    if os.path.exists(REF_GPU_SO):
        return RefLib(REF_GPU_SO)
    return None
