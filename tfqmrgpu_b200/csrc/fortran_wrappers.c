/* Fortran-callable shims: lower-case names with a trailing underscore, every argument by reference,
 * status returned through the last argument.  Same 18 symbols and argument lists as the reference's
 * tfQMRgpu/source/tfqmrgpu_Fortran_wrappers.c:58-187 so that its Fortran module
 * (tfqmrgpu_Fortran_module.F90) and F77 callers link unchanged.  Built into libtfQMRgpu.so and into
 * the static libtfQMRgpu_Fortran.a. */
#include <stddef.h>
#include <stdint.h>

typedef int64_t cudaStream_t; /* no CUDA headers needed here (tfqmrgpu_Fortran_wrappers.c:46) */
#include "../../include/tfqmrgpu.h"

#define H  tfqmrgpuHandle_t
#define PL tfqmrgpuBsrsvPlan_t
#define ST tfqmrgpuStatus_t

void tfqmrgpuprinterror_(ST const *status, ST *stat) { *stat = tfqmrgpuPrintError(*status); }

void tfqmrgpucreatehandle_(H *handle, ST *stat) { *handle = NULL; *stat = tfqmrgpuCreateHandle(handle); }

void tfqmrgpudestroyhandle_(H *handle, ST *stat) { *stat = tfqmrgpuDestroyHandle(*handle); *handle = NULL; }

void tfqmrgpusetstream_(H const *handle, cudaStream_t const *streamId, ST *stat) {
    *stat = tfqmrgpuSetStream(*handle, *streamId);
}

void tfqmrgpugetstream_(H const *handle, cudaStream_t *streamId, ST *stat) {
    *stat = tfqmrgpuGetStream(*handle, streamId);
}

/* Fortran index arrays start at 1 (tfqmrgpu_Fortran_wrappers.c:85) */
void tfqmrgpu_bsrsv_createplan_(H const *handle, PL *plan, int32_t const *mb,
        int32_t const *bsrRowPtrA, int32_t const *nnzbA, int32_t const *bsrColIndA,
        int32_t const *bsrRowPtrX, int32_t const *nnzbX, int32_t const *bsrColIndX,
        int32_t const *bsrRowPtrB, int32_t const *nnzbB, int32_t const *bsrColIndB,
        int32_t const *echo, ST *stat) {
    *plan = NULL;
    *stat = tfqmrgpu_bsrsv_createPlan(*handle, plan, *mb, bsrRowPtrA, *nnzbA, bsrColIndA,
                                      bsrRowPtrX, *nnzbX, bsrColIndX, bsrRowPtrB, *nnzbB, bsrColIndB, 1, *echo);
    if (TFQMRGPU_STATUS_SUCCESS != *stat) tfqmrgpuPrintError(*stat);
}

void tfqmrgpu_bsrsv_destroyplan_(H const *handle, PL *plan, ST *stat) {
    *stat = tfqmrgpu_bsrsv_destroyPlan(*handle, *plan);
    *plan = NULL;
}

void tfqmrgpu_bsrsv_buffersize_(H const *handle, PL const *plan, int32_t const *ldA, int32_t const *blockDim,
        int32_t const *ldB, int32_t const *RhsBlockDim, char const *precision, size_t *pBufferSizeInBytes, ST *stat) {
    *stat = tfqmrgpu_bsrsv_bufferSize(*handle, *plan, *ldA, *blockDim, *ldB, *RhsBlockDim, *precision, pBufferSizeInBytes);
}

/* device memory, like the reference (tfqmrgpu_Fortran_wrappers.c:123) */
void tfqmrgpucreateworkspace_(void **pBuffer, size_t const *pBufferSizeInBytes, ST *stat) {
    *stat = tfqmrgpuCreateWorkspace(pBuffer, *pBufferSizeInBytes, 'd');
}

void tfqmrgpudestroyworkspace_(void **pBuffer, ST *stat) { *stat = tfqmrgpuDestroyWorkspace(*pBuffer); }

void tfqmrgpu_bsrsv_setbuffer_(H const *handle, PL const *plan, void *const *pBuffer, ST *stat) {
    *stat = tfqmrgpu_bsrsv_setBuffer(*handle, *plan, *pBuffer);
}

void tfqmrgpu_bsrsv_getbuffer_(H const *handle, PL const *plan, void **pBuffer, ST *stat) {
    *stat = tfqmrgpu_bsrsv_getBuffer(*handle, *plan, pBuffer);
}

void tfqmrgpu_bsrsv_setmatrix_c_(H const *handle, PL const *plan, char const *var, float const *val,
        int32_t const *ld, int32_t const *d2, char const *trans, tfqmrgpuDataLayout_t const *layout, ST *stat) {
    *stat = tfqmrgpu_bsrsv_setMatrix(*handle, *plan, *var, (void const*)val, 'c', *ld, *d2, *trans, *layout);
}

void tfqmrgpu_bsrsv_setmatrix_z_(H const *handle, PL const *plan, char const *var, double const *val,
        int32_t const *ld, int32_t const *d2, char const *trans, tfqmrgpuDataLayout_t const *layout, ST *stat) {
    *stat = tfqmrgpu_bsrsv_setMatrix(*handle, *plan, *var, (void const*)val, 'z', *ld, *d2, *trans, *layout);
}

void tfqmrgpu_bsrsv_getmatrix_c_(H const *handle, PL const *plan, char const *var, float *val,
        int32_t const *ld, int32_t const *d2, char const *trans, tfqmrgpuDataLayout_t const *layout, ST *stat) {
    *stat = tfqmrgpu_bsrsv_getMatrix(*handle, *plan, *var, (void*)val, 'c', *ld, *d2, *trans, *layout);
}

void tfqmrgpu_bsrsv_getmatrix_z_(H const *handle, PL const *plan, char const *var, double *val,
        int32_t const *ld, int32_t const *d2, char const *trans, tfqmrgpuDataLayout_t const *layout, ST *stat) {
    *stat = tfqmrgpu_bsrsv_getMatrix(*handle, *plan, *var, (void*)val, 'z', *ld, *d2, *trans, *layout);
}

void tfqmrgpu_bsrsv_solve_(H const *handle, PL const *plan, double const *threshold, int32_t const *maxIterations, ST *stat) {
    *stat = tfqmrgpu_bsrsv_solve(*handle, *plan, *threshold, *maxIterations);
}

void tfqmrgpu_bsrsv_getinfo_(H const *handle, PL const *plan, double *residuum_reached, int32_t *iterations_needed,
        double *flops_performed, double *flops_performed_all, ST *stat) {
    *stat = tfqmrgpu_bsrsv_getInfo(*handle, *plan, residuum_reached, iterations_needed, flops_performed, flops_performed_all);
}
