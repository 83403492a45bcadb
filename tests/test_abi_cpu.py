"""CPU tests of the drop-in boundary: the library loads, exports every symbol that include/*.h declares,
and the GPU-free entry points behave like the reference (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np

from tfqmrgpu_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tfqmrgpux?_?\w*)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(L.LIB_PATH)
    abi = _declared("tfqmrgpu.h")
    assert sorted(abi) == sorted(L.ABI_SYMBOLS) and len(abi) == 21   # the reference exports exactly 21
    ext = _declared("tfqmrgpu_b200_ext.h")
    assert sorted(ext) == sorted(L.EXT_SYMBOLS)
    for name in abi + ext + L.FORTRAN_SYMBOLS:
        assert hasattr(lib, name), name


def test_constants_match_reference_header():
    txt = open(os.path.join(ROOT, "include", "tfqmrgpu.h")).read()
    want = dict(TFQMRGPU_STATUS_SUCCESS=0, TFQMRGPU_STATUS_MAX_ITERATIONS=9, TFQMRGPU_STATUS_BREAKDOWN=6,
                TFQMRGPU_STATUS_NO_INFO_PASSED=3, TFQMRGPU_POINTER_INVALID=7, TFQMRGPU_STATUS_ALLOCATION_FAILED=4,
                TFQMRGPU_STATUS_RANDOM_GEN_FAILED=5, TFQMRGPU_STATUS_LAUNCH_FAILED=2, TFQMRGPU_NO_IMPLEMENTATION=19,
                TFQMRGPU_UNDOCUMENTED_ERROR=14, TFQMRGPU_DATALAYOUT_UNKNOWN=15, TFQMRGPU_B_IS_NOT_SUBSET_OF_X=13,
                TFQMRGPU_B_HAS_A_ZERO_COLUMN=11, TFQMRGPU_BLOCKSIZE_MISSING=12, TFQMRGPU_TANSPOSITION_UNKNOWN=17,
                TFQMRGPU_VARIABLENAME_UNKNOWN=18, TFQMRGPU_PRECISION_MISSMATCH=16, TFQMRGPU_CODE_LINE=1000,
                TFQMRGPU_MEMORY_ALIGNMENT=8, TFQMRGPU_NUMBER_OF_INSTANCES_OF_X=7)
    for k, v in want.items():
        m = re.search(rf"\b{k}\s*=\s*([0-9x]+)\s*;", txt)
        assert m and int(m.group(1), 0) == v, k
    assert re.search(r"TFQMRGPU_CODE_CHAR\s*=\s*10000\*1000", txt)
    for k, v in dict(RRRRIIII=0x0f, RRIIRRII=0x33, RIRIRIRI=0x55).items():
        assert int(re.search(rf"TFQMRGPU_LAYOUT_{k}\s*=\s*(0x[0-9a-f]+)", txt).group(1), 16) == v


def test_error_strings_decode_like_reference():
    lib = L.load()
    s = lambda code: lib.tfqmrgpuGetErrorString(code).decode()
    assert s(0) == ""
    assert s(9) == "tfQMRgpu: Max number of iterations exceeded!"
    assert s(6) == "tfQMRgpu: All components have broken down!"
    assert s(13 + 1000*6) == "tfQMRgpu: B is not a subset of X in row 6!"
    assert s(11 + 1000*3) == "tfQMRgpu: B has 3 zero columns, will break!"
    assert s(12 + 10_000_000*7 + 1000*9) == "tfQMRgpu: Missing blocksize 7 x 9!"
    assert s(17 + 10_000_000*ord("q") + 1000*498) == "tfQMRgpu: Unknown transposition 'q' at line 498!"
    assert s(15 + 1000*0x77) == "tfQMRgpu: Unknown data layout '0x77'!"
    assert s(16 + 10_000_000*ord("c") + 1000*12) == "tfQMRgpu: Missmatch in precision 'c' at line 12!"
    assert "Unknown status" in s(1)
    assert L.decode_status(12 + 10_000_000*7 + 1000*9) == (12, 9, 7)


def test_allowed_block_sizes_semantics():
    """tfqmrgpu.cu:75-106 incl. the `2*n < arrayLength` rule."""
    lib = L.load()
    n = C.c_int32(0)
    arr = (C.c_int32*200)()
    assert lib.tfqmrgpu_bsrsv_allowedBlockSizes(C.byref(n), arr, 200) == 0
    pairs = [(arr[2*i], arr[2*i + 1]) for i in range(n.value)]
    assert n.value == 15
    assert pairs == [(4, 4), (4, 5), (4, 8), (4, 32), (8, 8), (8, 9), (8, 10), (8, 32), (8, 64), (16, 16), (16, 32),
                     (16, 64), (32, 32), (32, 64), (64, 64)]
    # arrayLength 30 is too short (needs > 2*15): error, count still reported
    n2 = C.c_int32(0)
    st = lib.tfqmrgpu_bsrsv_allowedBlockSizes(C.byref(n2), arr, 30)
    assert L.decode_status(st)[0] == L.UNDOCUMENTED_ERROR and n2.value == 15
    assert lib.tfqmrgpu_bsrsv_allowedBlockSizes(C.byref(n2), arr, 31) == 0
    assert L.decode_status(lib.tfqmrgpu_bsrsv_allowedBlockSizes(None, arr, 31))[0] == L.UNDOCUMENTED_ERROR
    for lm, ln in pairs:
        assert lib.tfqmrgpu_bsrsv_blockSizeMissing(lm, ln) == 0
    assert L.decode_status(lib.tfqmrgpu_bsrsv_blockSizeMissing(7, 9)) == (L.BLOCKSIZE_MISSING, 9, 7)
    assert L.decode_status(lib.tfqmrgpu_bsrsv_blockSizeMissing(8, 4))[0] == L.BLOCKSIZE_MISSING


def test_handle_and_stream_without_gpu():
    lib = L.load()
    h = C.c_void_p()
    assert lib.tfqmrgpuCreateHandle(C.byref(h)) == 0 and h.value
    # *handle must be NULL on entry (tfqmrgpu.cu:112)
    assert L.decode_status(lib.tfqmrgpuCreateHandle(C.byref(h)))[0] == L.UNDOCUMENTED_ERROR
    assert L.decode_status(lib.tfqmrgpuCreateHandle(None))[0] == L.UNDOCUMENTED_ERROR
    assert lib.tfqmrgpuSetStream(h, C.c_void_p(0x1234)) == 0
    s = C.c_void_p()
    assert lib.tfqmrgpuGetStream(h, C.byref(s)) == 0 and s.value == 0x1234
    assert lib.tfqmrgpuDestroyHandle(h) == 0
    assert L.decode_status(lib.tfqmrgpuDestroyHandle(None))[0] == L.UNDOCUMENTED_ERROR


def test_createplan_argument_checks_before_any_gpu_work():
    """tfqmrgpu.cu:161-172: these return before the device is touched."""
    lib = L.load()
    h = C.c_void_p(); lib.tfqmrgpuCreateHandle(C.byref(h))
    rp = np.array([0, 1], np.int32); ci = np.array([0], np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def create(mb, nA, nX, nB, plan=None, rpA=rp):
        plan = plan if plan is not None else C.c_void_p()
        return lib.tfqmrgpu_bsrsv_createPlan(h, C.byref(plan), mb, p(rpA), nA, p(ci), p(rp), nX, p(ci), p(rp), nB, p(ci), 0, 0)
    assert L.decode_status(create(1, 1, 1, 1, plan=C.c_void_p(0xdead)))[0] == L.POINTER_INVALID
    assert L.decode_status(create(0, 1, 1, 1))[0] == L.UNDOCUMENTED_ERROR     # mb < 1
    assert L.decode_status(create(1, 1, 0, 0))[0] == L.UNDOCUMENTED_ERROR     # nnzbX < 1
    assert L.decode_status(create(1, 1, 1, 2))[0] == L.UNDOCUMENTED_ERROR     # nnzbB > nnzbX
    assert L.decode_status(create(1, 2, 1, 1))[0] == L.UNDOCUMENTED_ERROR     # nnzbA > mb*mb
    assert L.decode_status(create(1, 1, 1, 1, rpA=np.array([0, 0], np.int32)))[0] == L.UNDOCUMENTED_ERROR  # nnz != rowPtr span
    assert L.decode_status(lib.tfqmrgpu_bsrsv_destroyPlan(h, None))[0] == L.POINTER_INVALID
    # getInfo with no out pointer -> NO_INFO_PASSED is only reachable with a plan; solve/getInfo on NULL plan are refused
    assert L.decode_status(lib.tfqmrgpu_bsrsv_solve(h, None, 1e-9, 10))[0] == L.POINTER_INVALID
    lib.tfqmrgpuDestroyHandle(h)


def test_fortran_shims_without_gpu():
    lib = C.CDLL(L.LIB_PATH)
    h = C.c_void_p(0xbeef); stat = C.c_int32(-1)
    lib.tfqmrgpucreatehandle_(C.byref(h), C.byref(stat))        # shim NULLs the handle first
    assert stat.value == 0 and h.value
    stream = C.c_int64(77); got = C.c_int64(0)
    lib.tfqmrgpusetstream_(C.byref(h), C.byref(stream), C.byref(stat)); assert stat.value == 0
    lib.tfqmrgpugetstream_(C.byref(h), C.byref(got), C.byref(stat)); assert stat.value == 0 and got.value == 77
    lib.tfqmrgpudestroyhandle_(C.byref(h), C.byref(stat))
    assert stat.value == 0 and not h.value
    code = C.c_int32(0)
    lib.tfqmrgpuprinterror_(C.byref(code), C.byref(stat)); assert stat.value == 0


def test_fortran_module_binds_the_exported_shims_with_their_argument_counts():
    """include/tfqmrgpu_Fortran_module.F90 cannot be compiled in this image (no Fortran compiler): check statically that every
    bind(C) interface names one of the 18 exported shims and passes as many arguments as the C shim takes
    (tfqmrgpu_Fortran_wrappers.c:58-187), and that all 18 are bound."""
    f90 = open(os.path.join(ROOT, "include", "tfqmrgpu_Fortran_module.F90")).read()
    f90 = re.sub(r"&\s*\n\s*&?", " ", f90)                                   # join continuation lines
    f90 = "\n".join(line.split("!")[0] for line in f90.splitlines())           # strip comments
    binds = re.findall(r"subroutine\s+\w+\s*\(([^)]*)\)\s*bind\s*\(\s*C\s*,\s*name\s*=\s*\"(\w+)\"\s*\)", f90, flags=re.I)
    assert binds
    csrc = open(os.path.join(ROOT, "tfqmrgpu_b200", "csrc", "fortran_wrappers.c")).read()
    csrc = re.sub(r"/\*.*?\*/", "", csrc, flags=re.S)
    cargs = {name: len([a for a in args.split(",") if a.strip()])
             for name, args in re.findall(r"^void\s+(\w+_)\s*\(([^)]*)\)", csrc, flags=re.M | re.S)}
    assert sorted(cargs) == sorted(L.FORTRAN_SYMBOLS) and len(cargs) == 18
    bound = {}
    for args, name in binds:
        assert name in cargs, name
        bound[name] = len([a for a in args.split(",") if a.strip()])
    assert sorted(bound) == sorted(cargs)
    for name, n in bound.items():
        assert n == cargs[name], (name, n, cargs[name])
    # the constants of the fixed-form header are those of tfqmrgpu.h
    hdr = open(os.path.join(ROOT, "include", "tfqmrgpu_Fortran.h")).read()
    for name, value in (("TFQMRGPU_STATUS_SUCCESS", 0), ("TFQMRGPU_LAYOUT_RRRRIIII", 15), ("TFQMRGPU_LAYOUT_RIRIRIRI", 85)):
        m = re.search(name + r"\s*=?\s*(\d+)", hdr)
        assert m and int(m.group(1)) == value, name


def test_python_package_refuses_to_run_without_library(tmp_path, monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "missing.so"))
    try:
        L.load()
        raise AssertionError("expected a loud failure")
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
