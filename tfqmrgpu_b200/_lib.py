"""ctypes loader for the product library ``libtfQMRgpu.so`` (built in-tree by ``make -C tfqmrgpu_b200/csrc``).

Fails loudly when the CUDA library is missing: there is no CPU fallback in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFQMRGPU_LIB", os.path.join(HERE, "lib", "libtfQMRgpu.so"))   # override: A/B builds

# constants of include/tfqmrgpu.h
STATUS_SUCCESS, STATUS_MAX_ITERATIONS, STATUS_BREAKDOWN = 0, 9, 6
NO_INFO_PASSED, POINTER_INVALID, ALLOCATION_FAILED, RANDOM_GEN_FAILED, LAUNCH_FAILED = 3, 7, 4, 5, 2
NO_IMPLEMENTATION, UNDOCUMENTED_ERROR, DATALAYOUT_UNKNOWN = 19, 14, 15
B_IS_NOT_SUBSET_OF_X, B_HAS_A_ZERO_COLUMN, BLOCKSIZE_MISSING = 13, 11, 12
TANSPOSITION_UNKNOWN, VARIABLENAME_UNKNOWN, PRECISION_MISSMATCH = 17, 18, 16
CODE_LINE, CODE_CHAR = 1000, 10_000_000
LAYOUT_RRRRIIII, LAYOUT_RRIIRRII, LAYOUT_RIRIRIRI = 0x0f, 0x33, 0x55

ABI_SYMBOLS = [
    "tfqmrgpuPrintError", "tfqmrgpuGetErrorString", "tfqmrgpuCreateHandle", "tfqmrgpuDestroyHandle",
    "tfqmrgpuSetStream", "tfqmrgpuGetStream", "tfqmrgpuCreateWorkspace", "tfqmrgpuDestroyWorkspace",
    "tfqmrgpu_bsrsv_allowedBlockSizes", "tfqmrgpu_bsrsv_blockSizeMissing", "tfqmrgpu_bsrsv_createPlan",
    "tfqmrgpu_bsrsv_destroyPlan", "tfqmrgpu_bsrsv_bufferSize", "tfqmrgpu_bsrsv_setBuffer",
    "tfqmrgpu_bsrsv_getBuffer", "tfqmrgpu_bsrsv_setMatrix", "tfqmrgpu_bsrsv_getMatrix", "tfqmrgpu_bsrsv_solve",
    "tfqmrgpu_bsrsv_getInfo", "tfqmrgpu_bsrsv_z", "tfqmrgpu_bsrsv_c",
]
EXT_SYMBOLS = [
    "tfqmrgpux_getVersion", "tfqmrgpux_setVerbosity", "tfqmrgpux_bsrsv_getPlanArray", "tfqmrgpux_bsrsv_getPlanInfo",
    "tfqmrgpux_bsrsv_setV3", "tfqmrgpux_bsrsv_getV3", "tfqmrgpux_bsrsv_multiply", "tfqmrgpux_bsrsv_getVector",
    "tfqmrgpux_bsrsv_getWindow", "tfqmrgpux_bsrsv_getRhsStatus", "tfqmrgpux_bsrsv_getSolveStats", "tfqmrgpux_randomShadow",
    "tfqmrgpux_bsrsv_setProfiling", "tfqmrgpux_bsrsv_getSolveProfile", "tfqmrgpux_bsrsv_setOperator",
    "tfqmrgpux_bsrsv_setDevices", "tfqmrgpux_bsrsv_getDevices", "tfqmrgpux_bsrsv_setShardExchange",
    "tfqmrgpux_bsrsv_setShardHints", "tfqmrgpux_bsrsv_getMatrixPartInfo", "tfqmrgpux_bsrsv_setMatrixPart", "tfqmrgpux_bsrsv_getTileBlocks", "tfqmrgpux_tileBlocksFor", "tfqmrgpux_bsrsv_setRhsTrivial", "tfqmrgpux_bsrsv_setEarlyFreeze", "tfqmrgpux_bsrsv_setInitialGuess", "tfqmrgpux_bsrsv_setPreconditioner", "tfqmrgpux_bsrsv_getMixedInfo",
]
FORTRAN_SYMBOLS = [
    "tfqmrgpuprinterror_", "tfqmrgpucreatehandle_", "tfqmrgpudestroyhandle_", "tfqmrgpusetstream_",
    "tfqmrgpugetstream_", "tfqmrgpu_bsrsv_createplan_", "tfqmrgpu_bsrsv_destroyplan_", "tfqmrgpu_bsrsv_buffersize_",
    "tfqmrgpucreateworkspace_", "tfqmrgpudestroyworkspace_", "tfqmrgpu_bsrsv_setbuffer_", "tfqmrgpu_bsrsv_getbuffer_",
    "tfqmrgpu_bsrsv_setmatrix_c_", "tfqmrgpu_bsrsv_setmatrix_z_", "tfqmrgpu_bsrsv_getmatrix_c_",
    "tfqmrgpu_bsrsv_getmatrix_z_", "tfqmrgpu_bsrsv_solve_", "tfqmrgpu_bsrsv_getinfo_",
]

_lib = None


def decode_status(status: int) -> tuple[int, int, int]:
    """status -> (code, line/payload, char payload) (tfqmrgpu_error_tool.cxx:38-42)."""
    key = status // CODE_CHAR
    rest = status - key*CODE_CHAR
    line = rest // CODE_LINE
    return rest - line*CODE_LINE, line, key


def load():
    """Load libtfQMRgpu.so and declare the prototypes of include/tfqmrgpu.h (+ extensions)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing - build it with `make -C tfqmrgpu_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    st = C.c_int32
    lib.tfqmrgpuPrintError.restype = st; lib.tfqmrgpuPrintError.argtypes = [st]
    lib.tfqmrgpuGetErrorString.restype = C.c_char_p; lib.tfqmrgpuGetErrorString.argtypes = [st]
    lib.tfqmrgpuCreateHandle.restype = st; lib.tfqmrgpuCreateHandle.argtypes = [C.POINTER(vp)]
    lib.tfqmrgpuDestroyHandle.restype = st; lib.tfqmrgpuDestroyHandle.argtypes = [vp]
    lib.tfqmrgpuSetStream.restype = st; lib.tfqmrgpuSetStream.argtypes = [vp, vp]
    lib.tfqmrgpuGetStream.restype = st; lib.tfqmrgpuGetStream.argtypes = [vp, C.POINTER(vp)]
    lib.tfqmrgpuCreateWorkspace.restype = st; lib.tfqmrgpuCreateWorkspace.argtypes = [C.POINTER(vp), C.c_size_t, C.c_char]
    lib.tfqmrgpuDestroyWorkspace.restype = st; lib.tfqmrgpuDestroyWorkspace.argtypes = [vp]
    lib.tfqmrgpu_bsrsv_allowedBlockSizes.restype = st
    lib.tfqmrgpu_bsrsv_allowedBlockSizes.argtypes = [i32p, i32p, C.c_int]
    lib.tfqmrgpu_bsrsv_blockSizeMissing.restype = st; lib.tfqmrgpu_bsrsv_blockSizeMissing.argtypes = [C.c_int, C.c_int]
    lib.tfqmrgpu_bsrsv_createPlan.restype = st
    lib.tfqmrgpu_bsrsv_createPlan.argtypes = [vp, C.POINTER(vp), C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp,
                                              C.c_int, C.c_int]
    lib.tfqmrgpu_bsrsv_destroyPlan.restype = st; lib.tfqmrgpu_bsrsv_destroyPlan.argtypes = [vp, vp]
    lib.tfqmrgpu_bsrsv_bufferSize.restype = st
    lib.tfqmrgpu_bsrsv_bufferSize.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char, C.POINTER(C.c_size_t)]
    lib.tfqmrgpu_bsrsv_setBuffer.restype = st; lib.tfqmrgpu_bsrsv_setBuffer.argtypes = [vp, vp, vp]
    lib.tfqmrgpu_bsrsv_getBuffer.restype = st; lib.tfqmrgpu_bsrsv_getBuffer.argtypes = [vp, vp, C.POINTER(vp)]
    for name in ("tfqmrgpu_bsrsv_setMatrix", "tfqmrgpu_bsrsv_getMatrix"):
        f = getattr(lib, name)
        f.restype = st
        f.argtypes = [vp, vp, C.c_char, vp, C.c_char, C.c_int, C.c_int, C.c_char, C.c_int]
    lib.tfqmrgpu_bsrsv_solve.restype = st; lib.tfqmrgpu_bsrsv_solve.argtypes = [vp, vp, C.c_double, C.c_int]
    lib.tfqmrgpu_bsrsv_getInfo.restype = st
    lib.tfqmrgpu_bsrsv_getInfo.argtypes = [vp, vp, C.POINTER(C.c_double), i32p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in ("tfqmrgpu_bsrsv_z", "tfqmrgpu_bsrsv_c"):
        f = getattr(lib, name)
        f.restype = st
        f.argtypes = [C.c_int, C.c_int, C.c_int,
                      vp, C.c_int, vp, vp, C.c_char,
                      vp, C.c_int, vp, vp, C.c_char,
                      vp, C.c_int, vp, vp, C.c_char,
                      i32p, C.POINTER(C.c_float), C.c_int, C.c_int]
    # extensions
    lib.tfqmrgpux_getVersion.restype = st; lib.tfqmrgpux_getVersion.argtypes = [C.POINTER(C.c_int)]*3
    lib.tfqmrgpux_setVerbosity.restype = st; lib.tfqmrgpux_setVerbosity.argtypes = [C.c_int]
    lib.tfqmrgpux_bsrsv_getPlanArray.restype = st
    lib.tfqmrgpux_bsrsv_getPlanArray.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_size_t)]
    lib.tfqmrgpux_bsrsv_getPlanInfo.restype = st; lib.tfqmrgpux_bsrsv_getPlanInfo.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.tfqmrgpux_bsrsv_setV3.restype = st; lib.tfqmrgpux_bsrsv_setV3.argtypes = [vp, vp, vp, C.c_int]
    lib.tfqmrgpux_bsrsv_getV3.restype = st; lib.tfqmrgpux_bsrsv_getV3.argtypes = [vp, vp, vp]
    lib.tfqmrgpux_bsrsv_multiply.restype = st; lib.tfqmrgpux_bsrsv_multiply.argtypes = [vp, vp, C.c_int]
    lib.tfqmrgpux_bsrsv_getVector.restype = st
    lib.tfqmrgpux_bsrsv_getVector.argtypes = [vp, vp, C.c_char, vp, C.c_char, C.c_char, C.c_int]
    lib.tfqmrgpux_bsrsv_getWindow.restype = st
    lib.tfqmrgpux_bsrsv_getWindow.argtypes = [vp, C.c_char, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.tfqmrgpux_bsrsv_getRhsStatus.restype = st; lib.tfqmrgpux_bsrsv_getRhsStatus.argtypes = [vp, vp, vp]
    lib.tfqmrgpux_bsrsv_getSolveStats.restype = st; lib.tfqmrgpux_bsrsv_getSolveStats.argtypes = [vp, C.POINTER(C.c_double)]
    lib.tfqmrgpux_randomShadow.restype = st; lib.tfqmrgpux_randomShadow.argtypes = [vp, vp, C.c_size_t]
    lib.tfqmrgpux_bsrsv_setProfiling.restype = st; lib.tfqmrgpux_bsrsv_setProfiling.argtypes = [vp, C.c_int]
    lib.tfqmrgpux_bsrsv_getSolveProfile.restype = st; lib.tfqmrgpux_bsrsv_getSolveProfile.argtypes = [vp, C.POINTER(C.c_double)]
    lib.tfqmrgpux_bsrsv_setOperator.restype = st; lib.tfqmrgpux_bsrsv_setOperator.argtypes = [vp, vp, vp]
    lib.tfqmrgpux_bsrsv_setPreconditioner.restype = st; lib.tfqmrgpux_bsrsv_setPreconditioner.argtypes = [vp, vp, vp]
    lib.tfqmrgpux_bsrsv_setDevices.restype = st; lib.tfqmrgpux_bsrsv_setDevices.argtypes = [vp, vp, C.c_int, i32p]
    lib.tfqmrgpux_bsrsv_getDevices.restype = st; lib.tfqmrgpux_bsrsv_getDevices.argtypes = [vp, C.POINTER(C.c_int), i32p, C.c_int]
    lib.tfqmrgpux_bsrsv_setShardExchange.restype = st
    lib.tfqmrgpux_bsrsv_setShardExchange.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp]
    lib.tfqmrgpux_bsrsv_setShardHints.restype = st; lib.tfqmrgpux_bsrsv_setShardHints.argtypes = [vp, C.c_int64, C.c_int32]
    lib.tfqmrgpux_bsrsv_getTileBlocks.restype = st; lib.tfqmrgpux_bsrsv_getTileBlocks.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.tfqmrgpux_bsrsv_getMatrixPartInfo.restype = st; lib.tfqmrgpux_bsrsv_getMatrixPartInfo.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.tfqmrgpux_bsrsv_setMatrixPart.restype = st
    lib.tfqmrgpux_bsrsv_setMatrixPart.argtypes = [vp, vp, vp, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.tfqmrgpux_bsrsv_setRhsTrivial.restype = st; lib.tfqmrgpux_bsrsv_setRhsTrivial.argtypes = [vp, vp]
    lib.tfqmrgpux_bsrsv_setEarlyFreeze.restype = st; lib.tfqmrgpux_bsrsv_setEarlyFreeze.argtypes = [vp, C.c_int]
    lib.tfqmrgpux_bsrsv_setInitialGuess.restype = st; lib.tfqmrgpux_bsrsv_setInitialGuess.argtypes = [vp, C.c_int]
    lib.tfqmrgpux_bsrsv_getMixedInfo.restype = st; lib.tfqmrgpux_bsrsv_getMixedInfo.argtypes = [vp, C.POINTER(C.c_double)]
    lib.tfqmrgpux_tileBlocksFor.restype = st; lib.tfqmrgpux_tileBlocksFor.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
