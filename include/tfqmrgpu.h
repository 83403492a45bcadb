/*
 * tfqmrgpu.h - C-ABI of the B200-native tfQMR solver (libtfQMRgpu.so).
 *
 * Drop-in boundary: the 21 entry points, types and constants below are binary- and source-compatible
 * with real-space/tfQMRgpu's  tfQMRgpu/include/tfqmrgpu.h:5-191.  Each declaration cites the
 * reference declaration it replaces and the reference implementation whose behaviour it keeps.
 * Existing C, Fortran (via tfqmrgpu_Fortran_wrappers.c), Julia (@ccall) and Python (ctypes) callers
 * link against this library unchanged.
 *
 * The includer must provide `cudaStream_t` (either through <cuda_runtime.h> or, for callers without
 * CUDA headers, `typedef size_t cudaStream_t;` as in the reference's C example).
 *
 * Workflow:  CreateHandle -> SetStream -> bsrsv_createPlan -> bsrsv_bufferSize -> (allocate device
 * memory) -> bsrsv_setBuffer -> bsrsv_setMatrix('A'), ('B') -> bsrsv_solve -> bsrsv_getInfo,
 * bsrsv_getMatrix('X') -> bsrsv_destroyPlan -> DestroyHandle.   All functions return 0 on success,
 * otherwise  code + 1000*payload + 10^7*char_payload  (decode with tfqmrgpuGetErrorString).
 */
#ifndef TFQMRGPU_H
#define TFQMRGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types (ref tfqmrgpu.h:5-9) ---------------------------------------------------------------- */
typedef int32_t tfqmrgpuStatus_t;      /* status / error code                                        */
typedef void*   tfqmrgpuHandle_t;      /* opaque library handle (holds the stream)                   */
typedef int*    tfqmrgpuBsrsvPlan_t;   /* opaque plan of one block-sparse solve  A * X == B          */
typedef int     tfqmrgpuDataLayout_t;  /* arrangement of real/imaginary parts inside a block         */

/* ---- error reporting (ref tfqmrgpu.h:16-17, tfqmrgpu_error_tool.cxx:33-76) --------------------- */
tfqmrgpuStatus_t tfqmrgpuPrintError(tfqmrgpuStatus_t const status);
char const*      tfqmrgpuGetErrorString(tfqmrgpuStatus_t const status); /* static buffer, not thread-safe */

/* ---- handle and stream (ref tfqmrgpu.h:20-28, tfqmrgpu.cu:110-134) ----------------------------- */
tfqmrgpuStatus_t tfqmrgpuCreateHandle(tfqmrgpuHandle_t *handle);   /* *handle must be NULL on entry */
tfqmrgpuStatus_t tfqmrgpuDestroyHandle(tfqmrgpuHandle_t handle);
tfqmrgpuStatus_t tfqmrgpuSetStream(tfqmrgpuHandle_t handle, cudaStream_t const streamId);
tfqmrgpuStatus_t tfqmrgpuGetStream(tfqmrgpuHandle_t handle, cudaStream_t *streamId);

/* ---- workspace helpers (ref tfqmrgpu.h:30-31, tfqmrgpu.cu:682-698) ------------------------------
 * memType 'm'/'M' = managed memory, anything else = device memory. */
tfqmrgpuStatus_t tfqmrgpuCreateWorkspace(void* *pBuffer, size_t const pBufferSizeInBytes, char const memType);
tfqmrgpuStatus_t tfqmrgpuDestroyWorkspace(void* pBuffer);

/* ---- supported block sizes (ref tfqmrgpu.h:33-38, tfqmrgpu.cu:75-106, allowed_block_sizes.h) ----
 * *number receives the count of (ldA, ldB) pairs; pairs are written while 2*count < arrayLength. */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_allowedBlockSizes(int32_t *number, int32_t *blockSizes, int const arrayLength);
tfqmrgpuStatus_t tfqmrgpu_bsrsv_blockSizeMissing(int const ldA, int const ldB); /* 0 if supported */

/* ---- plan analysis (ref tfqmrgpu.h:47-60, tfqmrgpu.cu:136-351) ----------------------------------
 * A is mb x mb blocks, X and B are mb block rows; Y = A*X has the pattern of X; B must be a subset
 * of X and every block column of X must hold at least one B block.  *plan must be NULL on entry.
 * indexOffset 0 (C) or 1 (Fortran/Julia) applies to the RowPtr arrays and to bsrColIndA. */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_createPlan(tfqmrgpuHandle_t handle,
    tfqmrgpuBsrsvPlan_t *plan,
    int     const mb,
    int32_t const *bsrRowPtrA, int const nnzbA, int32_t const *bsrColIndA,
    int32_t const *bsrRowPtrX, int const nnzbX, int32_t const *bsrColIndX,
    int32_t const *bsrRowPtrB, int const nnzbB, int32_t const *bsrColIndB,
    int     const indexOffset,
    int     const echo);

/* ref tfqmrgpu.h:62-63, tfqmrgpu.cu:353-361 */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_destroyPlan(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan);

/* ---- workspace size (ref tfqmrgpu.h:66-73, tfqmrgpu.cu:364-412) ---------------------------------
 * Requires ldA == blockDim, ldA <= ldB == RhsBlockDim.  precision 'c'/'f' -> complex<float>,
 * 'z'/'d' -> complex<double>.  'm' ("start with float and converge double", ref tfqmrgpu.h:72): the reference accepts it here and
 * rejects it at solve (tfqmrgpu.cu:42-44); this library implements it - double data in and out, complex<float> iterations inside an
 * fp64 refinement loop, see tfqmrgpu_b200_ext.h - unless TFQMRGPU_MIXED=0 asks for the reference's PRECISION_MISSMATCH. */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_bufferSize(tfqmrgpuHandle_t handle,
    tfqmrgpuBsrsvPlan_t plan,
    int const ldA, int const blockDim, int const ldB, int const RhsBlockDim,
    char const precision,
    size_t *pBufferSizeInBytes);

/* ref tfqmrgpu.h:79-85, tfqmrgpu.cu:415-462: registers the caller-owned device buffer, fills the
 * random shadow vector (cuRAND XORWOW, seed 1234, like the reference) */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_setBuffer(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, void* const pBuffer);
tfqmrgpuStatus_t tfqmrgpu_bsrsv_getBuffer(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, void* *pBuffer);

/* ---- operand upload / download (ref tfqmrgpu.h:87-105, tfqmrgpu.cu:467-645) ---------------------
 * var 'A' | 'B' | 'X' for set, only 'X' for get.  val is a HOST array of nnzb blocks,
 * A: [ldA][ldA], B/X: [ldA][ldB] complex numbers in `layout`.  trans 'n' | 't' | 'c'/'h' (conjugate
 * transpose) | '*' (conjugate).  ld and d2 are accepted for compatibility and not used. */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_setMatrix(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan,
    char const var, void const *val, char const precision, int const ld, int const d2,
    char const trans, tfqmrgpuDataLayout_t const layout);
tfqmrgpuStatus_t tfqmrgpu_bsrsv_getMatrix(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan,
    char const var, void *val, char const precision, int const ld, int const d2,
    char const trans, tfqmrgpuDataLayout_t const layout);

/* ---- the solver (ref tfqmrgpu.h:107-117, tfqmrgpu.cu:648-679, tfqmrgpu_core.hxx:20-335) --------- */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_solve(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan,
    double const threshold, int const maxIterations);
tfqmrgpuStatus_t tfqmrgpu_bsrsv_getInfo(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan,
    double *residuum_reached, int32_t *iterations_needed,
    double *flops_performed, double *flops_performed_all);

/* ---- quick starters (ref tfqmrgpu.h:138-156, tfqmrgpu.cu:702-821) -------------------------------
 * Host arrays Amat[nnzbA][ldA][ldA][2], Xmat[nnzbX][ldA][ldB][2], Bmat[nnzbB][ldA][ldB][2].
 * *iterations: in max iterations, out iterations needed.  *residual: in threshold, out reached. */
tfqmrgpuStatus_t tfqmrgpu_bsrsv_z(int mb, int ldA, int ldB,
    int32_t const* rowPtrA, int nnzbA, int32_t const* colIndA, double const* Amat, char transA,
    int32_t const* rowPtrX, int nnzbX, int32_t const* colIndX, double      * Xmat, char transX,
    int32_t const* rowPtrB, int nnzbB, int32_t const* colIndB, double const* Bmat, char transB,
    int32_t *iterations, float *residual, int indexOffset, int echo);
tfqmrgpuStatus_t tfqmrgpu_bsrsv_c(int mb, int ldA, int ldB,
    int32_t const* rowPtrA, int nnzbA, int32_t const* colIndA, float const* Amat, char transA,
    int32_t const* rowPtrX, int nnzbX, int32_t const* colIndX, float      * Xmat, char transX,
    int32_t const* rowPtrB, int nnzbB, int32_t const* colIndB, float const* Bmat, char transB,
    int32_t *iterations, float *residual, int indexOffset, int echo);

#ifdef __cplusplus
} /* extern "C" */
#endif

/* ---- constants (ref tfqmrgpu.h:160-191) --------------------------------------------------------- */
#define TFQMRGPU_DECL_CONST static const

TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_SUCCESS           =  0;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_LAUNCH_FAILED     =  2;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_NO_INFO_PASSED    =  3;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_ALLOCATION_FAILED =  4;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_RANDOM_GEN_FAILED =  5;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_BREAKDOWN         =  6;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_POINTER_INVALID          =  7;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_STATUS_MAX_ITERATIONS    =  9;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_B_HAS_A_ZERO_COLUMN      = 11;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_BLOCKSIZE_MISSING        = 12;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_B_IS_NOT_SUBSET_OF_X     = 13;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_UNDOCUMENTED_ERROR       = 14;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_DATALAYOUT_UNKNOWN       = 15;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_PRECISION_MISSMATCH      = 16;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_TANSPOSITION_UNKNOWN     = 17;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_VARIABLENAME_UNKNOWN     = 18;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_NO_IMPLEMENTATION        = 19;
/* payload encoding: lowest 3 decimal digits = code, next 4 = line/payload, upper 3 = an ASCII char */
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_CODE_LINE                = 1000;
TFQMRGPU_DECL_CONST tfqmrgpuStatus_t TFQMRGPU_CODE_CHAR                = 10000*1000;

/* block layouts, written for a 2x2 block: bit = 1 where an imaginary part is stored */
TFQMRGPU_DECL_CONST tfqmrgpuDataLayout_t TFQMRGPU_LAYOUT_RRRRIIII = 0x0f; /* planes: all real, then all imaginary  */
TFQMRGPU_DECL_CONST tfqmrgpuDataLayout_t TFQMRGPU_LAYOUT_RRIIRRII = 0x33; /* per row: real row, imaginary row     */
TFQMRGPU_DECL_CONST tfqmrgpuDataLayout_t TFQMRGPU_LAYOUT_RIRIRIRI = 0x55; /* interleaved (std::complex, Fortran)  */

TFQMRGPU_DECL_CONST size_t TFQMRGPU_MEMORY_ALIGNMENT = 8;         /* log2 of the workspace alignment: 256 bytes */
TFQMRGPU_DECL_CONST int    TFQMRGPU_NUMBER_OF_INSTANCES_OF_X = 7; /* X-shaped vectors kept in the workspace     */

#endif /* TFQMRGPU_H */
