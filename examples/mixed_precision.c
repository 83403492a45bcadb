/*
 * mixed_precision.c - a plain C caller of the tfqmrgpu.h C-ABI that solves the same block-sparse system twice:
 * with precision 'z' (complex double) and with precision 'm' ("start with float and converge double", tfqmrgpu.h:72).
 *
 * With the reference library the second solve ends with TFQMRGPU_PRECISION_MISSMATCH (tfqmrgpu.cu:42-44: the 'm' case of its solver
 * dispatch is commented out).  With this library ONE character changes between the two runs - the precision passed to
 * tfqmrgpu_bsrsv_bufferSize - and both return the same X: the 'm' run iterates in complex fp32 (on the tensor cores for 16, 32 and 64
 * blocks) inside an fp64 refinement loop.  Only tfqmrgpu.h is used, no extension.
 *
 * The system: mb block rows, A block-tridiagonal and periodic (diagonal blocks (4 + sigma)*1 + noise, neighbours -1 + noise),
 * X dense in nc block columns, B = unit blocks on the first nc block rows.
 *
 *   cc -I../include mixed_precision.c -L../tfqmrgpu_b200/lib -ltfQMRgpu -lm -o mixed_precision
 *   ./mixed_precision [mb=64] [block=16] [nc=2] [threshold=1e-10]
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
typedef size_t cudaStream_t; /* no CUDA headers needed, like in the reference's C example */
#include "tfqmrgpu.h"

#define CHECK(call) do { tfqmrgpuStatus_t s_ = (call); if (TFQMRGPU_STATUS_SUCCESS != s_) { \
    printf("%s failed with status %d at line %d\n", #call, (int)s_, __LINE__); tfqmrgpuPrintError(s_); return s_; } } while (0)

static double noise(unsigned *state) { *state = *state*1664525u + 1013904223u; return ((*state >> 8)/16777216.0 - 0.5)*0.1; }

static tfqmrgpuStatus_t run(char precision, int mb, int ld, int nc, double threshold,
                            int32_t const *rpA, int32_t const *ciA, double const *A,
                            int32_t const *rpX, int32_t const *ciX, double *X,
                            int32_t const *rpB, int32_t const *ciB, double const *B,
                            double *residual, int32_t *iterations, size_t *bytes)
{
    tfqmrgpuHandle_t handle = NULL;
    tfqmrgpuBsrsvPlan_t plan = NULL;
    void *buffer = NULL;
    CHECK(tfqmrgpuCreateHandle(&handle));
    CHECK(tfqmrgpuSetStream(handle, 0));
    CHECK(tfqmrgpu_bsrsv_createPlan(handle, &plan, mb, rpA, rpA[mb], ciA, rpX, rpX[mb], ciX, rpB, rpB[mb], ciB, 0, 0));
    CHECK(tfqmrgpu_bsrsv_bufferSize(handle, plan, ld, ld, ld, ld, precision, bytes));     /* <- the one character: 'z' or 'm' */
    CHECK(tfqmrgpuCreateWorkspace(&buffer, *bytes, 'd'));
    CHECK(tfqmrgpu_bsrsv_setBuffer(handle, plan, buffer));
    CHECK(tfqmrgpu_bsrsv_setMatrix(handle, plan, 'A', A, 'z', ld, ld, 'n', TFQMRGPU_LAYOUT_RIRIRIRI));   /* double data in both runs */
    CHECK(tfqmrgpu_bsrsv_setMatrix(handle, plan, 'B', B, 'z', ld, ld, 'n', TFQMRGPU_LAYOUT_RIRIRIRI));
    tfqmrgpuStatus_t const st = tfqmrgpu_bsrsv_solve(handle, plan, threshold, 200);
    CHECK(tfqmrgpu_bsrsv_getInfo(handle, plan, residual, iterations, NULL, NULL));
    CHECK(tfqmrgpu_bsrsv_getMatrix(handle, plan, 'X', X, 'z', ld, ld, 'n', TFQMRGPU_LAYOUT_RIRIRIRI));
    CHECK(tfqmrgpuDestroyWorkspace(buffer));
    CHECK(tfqmrgpu_bsrsv_destroyPlan(handle, plan));
    CHECK(tfqmrgpuDestroyHandle(handle));
    return st;
}

int main(int argc, char **argv)
{
    int const mb = (argc > 1) ? atoi(argv[1]) : 64, ld = (argc > 2) ? atoi(argv[2]) : 16, nc = (argc > 3) ? atoi(argv[3]) : 2;
    double const threshold = (argc > 4) ? atof(argv[4]) : 1e-10, sigma = 1.0;
    if (mb < 3 || nc < 1 || nc > mb || tfqmrgpu_bsrsv_blockSizeMissing(ld, ld)) { printf("usage: %s [mb>=3] [block in 4,8,16,32,64] [nc<=mb] [threshold]\n", argv[0]); return 1; }
    size_t const bs = (size_t)ld*ld*2;     /* doubles per block, RIRIRIRI */
    int32_t *rpA = malloc((mb + 1)*sizeof(int32_t)), *ciA = malloc(3*(size_t)mb*sizeof(int32_t));
    int32_t *rpX = malloc((mb + 1)*sizeof(int32_t)), *ciX = malloc((size_t)mb*nc*sizeof(int32_t));
    int32_t *rpB = malloc((mb + 1)*sizeof(int32_t)), *ciB = malloc((size_t)nc*sizeof(int32_t));
    double *A = calloc(3*(size_t)mb*bs, sizeof(double)), *B = calloc((size_t)nc*bs, sizeof(double));
    double *Xz = calloc((size_t)mb*nc*bs, sizeof(double)), *Xm = calloc((size_t)mb*nc*bs, sizeof(double));
    if (!rpA || !ciA || !rpX || !ciX || !rpB || !ciB || !A || !B || !Xz || !Xm) { printf("out of memory\n"); return 1; }
    unsigned seed = 12345u;
    for (int r = 0; r < mb; ++r) {
        rpA[r] = 3*r; rpX[r] = nc*r; rpB[r] = (r < nc) ? r : nc;
        for (int q = 0; q < 3; ++q) {
            int const c = (r + q - 1 + mb) % mb;
            ciA[3*r + q] = c;
            double *blk = A + (size_t)(3*r + q)*bs;
            for (int i = 0; i < ld; ++i) for (int k = 0; k < ld; ++k) {
                blk[((size_t)i*ld + k)*2 + 0] = noise(&seed) + ((i == k) ? ((c == r) ? 4.0 + sigma : -1.0) : 0.0);
                blk[((size_t)i*ld + k)*2 + 1] = noise(&seed);
            }
        }
        for (int c = 0; c < nc; ++c) ciX[nc*r + c] = c;
    }
    rpA[mb] = 3*mb; rpX[mb] = nc*mb; rpB[mb] = nc;
    for (int c = 0; c < nc; ++c) { ciB[c] = c; for (int i = 0; i < ld; ++i) B[(size_t)c*bs + ((size_t)i*ld + i)*2] = 1.0; }

    double res_z = 0, res_m = 0; int32_t it_z = 0, it_m = 0; size_t bytes_z = 0, bytes_m = 0;
    tfqmrgpuStatus_t const sz = run('z', mb, ld, nc, threshold, rpA, ciA, A, rpX, ciX, Xz, rpB, ciB, B, &res_z, &it_z, &bytes_z);
    tfqmrgpuStatus_t const sm = run('m', mb, ld, nc, threshold, rpA, ciA, A, rpX, ciX, Xm, rpB, ciB, B, &res_m, &it_m, &bytes_m);
    double dev = 0, big = 0;
    for (size_t i = 0; i < (size_t)mb*nc*bs; ++i) { dev = fmax(dev, fabs(Xz[i] - Xm[i])); big = fmax(big, fabs(Xz[i])); }
    printf("# mixed_precision: %d block rows of %dx%d, %d right-hand sides, threshold %.1e\n", mb, ld, ld, nc*ld, threshold);
    printf("# precision z: status %d, %d iterations, residual %.3e, workspace %.1f MB\n", (int)sz, (int)it_z, res_z, bytes_z*1e-6);
    printf("# precision m: status %d, %d iterations, residual %.3e, workspace %.1f MB\n", (int)sm, (int)it_m, res_m, bytes_m*1e-6);
    printf("# max|X_m - X_z| = %.3e at max|X| = %.3e\n", dev, big);
    int const ok = (0 == sz) && (0 == sm) && (res_m <= threshold) && (dev <= 100*threshold*big);
    printf("%s\n", ok ? "mixed_precision: OK" : "mixed_precision: FAILED");
    free(rpA); free(ciA); free(rpX); free(ciX); free(rpB); free(ciB); free(A); free(B); free(Xz); free(Xm);
    return ok ? 0 : 2;
}
