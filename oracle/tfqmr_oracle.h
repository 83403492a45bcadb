/* TEST INFRASTRUCTURE (oracle). Not part of the product; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the tfQMRgpu hot path (plan analysis, BSR block-sparse multiply,
 * per-column tfQMR vector algebra, decisions and the host-side convergence logic).
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Pinned against the unmodified reference built in oracle/_ref (tests/test_oracle_golden.py) and
 * against the committed golden vectors in tests/golden/.
 */
#ifndef TFQMR_ORACLE_H
#define TFQMR_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* arithmetic flavour of the restatement */
enum {
    ORC_MODE_GPU    = 0, /* reference CUDA path: real_t*float products in dotp (linalg.hxx:499-507),
                            one accumulator over all pairs in the block product (blockmult.hxx:28-82) */
    ORC_MODE_CPUREF = 1  /* reference HAS_NO_CUDA path: dotp promotes to double (linalg.hxx:574-583),
                            per-pair partial sums in the block product (blocksparse.hxx:159-186).
                            Used to pin the oracle bit-for-bit against oracle/_ref/libtfqmr_ref_cpu.so */
};

typedef struct orc_plan {
    int32_t  mb, nnzbA, nnzbX, nnzbB;
    uint32_t nCols;
    uint64_t nPairs;
    uint32_t *starts;   /* [nnzbX+1]  */
    uint32_t *pairs;    /* [2*nPairs] (inzA, inzX) */
    uint32_t *subset;   /* [nnzbB]    */
    uint16_t *colindx;  /* [nnzbX]    */
} orc_plan_t;

/* tfqmrgpu.cu:136-351. Returns the reference's status code incl. payload (code + 1000*payload);
 * *plan is NULL on error. Line-number payloads of UNDOCUMENTED_ERROR are returned as payload 0. */
int  orc_create_plan(int mb,
                     const int32_t *rpA, int nnzbA, const int32_t *ciA,
                     const int32_t *rpX, int nnzbX, const int32_t *ciX,
                     const int32_t *rpB, int nnzbB, const int32_t *ciB,
                     int indexOffset, orc_plan_t **plan);
void orc_destroy_plan(orc_plan_t *plan);
/* kind: 0 starts, 1 pairs, 2 subset, 3 colindx(u16); returns element count, copies if out != NULL */
uint64_t orc_plan_array(const orc_plan_t *plan, int kind, void *out);

/* tfqmrgpu.cu:75-106 + allowed_block_sizes.h:4-18 */
int  orc_block_size_allowed(int ldA, int ldB);
/* bufferSize formula of the reference's memcount pass (core.hxx:45-99, blocksparse.hxx:46-51) */
uint64_t orc_ref_buffer_size(const orc_plan_t *plan, int LM, int LN, int is_double);

/* layout conversion of tfqmrgpu.cu:467-603 + linalg.hxx:282-380 (setMatrix direction: host -> internal
 * real_t[nnzb][2][nRows][nCols]; getMatrix direction: internal -> host). var is 'A','B','X' ('A' flips
 * the transposition, tfqmrgpu.cu:509-520). Rectangular 't' uses the mathematically correct stride
 * (the reference indexes out of the block there). Returns 0 or the status code. */
int  orc_import_blocks_d(double *internal, const double *host, uint32_t nnzb, int nRows, int nCols,
                         int layout, char trans, char var);
int  orc_import_blocks_f(float  *internal, const float  *host, uint32_t nnzb, int nRows, int nCols,
                         int layout, char trans, char var);
int  orc_export_blocks_d(double *host, const double *internal, uint32_t nnzb, int nRows, int nCols,
                         int layout, char trans);
int  orc_export_blocks_f(float  *host, const float  *internal, uint32_t nnzb, int nRows, int nCols,
                         int layout, char trans);

/* bench_tfqmrgpu.cu:275-286 fill; cos/sin evaluated in double then cast */
void orc_fill_cos_sin_d(double *c, uint32_t nmat, int LM, int LN);
void orc_fill_cos_sin_f(float  *c, uint32_t nmat, int LM, int LN);

/* Y = A*X with A stored [k][i] (blockmult.hxx:36-92; CPU check loop bench_tfqmrgpu.cu:366-392).
 * nthreads > 1 uses OpenMP over Y blocks like the bench's check loop. */
void orc_multiply_d(double *Y, const double *A, const double *X, const uint32_t *starts,
                    const uint32_t *pairs, uint32_t nnzbY, int LM, int LN, int mode, int nthreads);
void orc_multiply_f(float  *Y, const float  *A, const float  *X, const uint32_t *starts,
                    const uint32_t *pairs, uint32_t nnzbY, int LM, int LN, int mode, int nthreads);

/* glibc rand()-stream v3 of the reference CPU path (linalg.hxx:799-802); reseeds with srand(1) first */
void orc_v3_glibc(float *v3, size_t n);

typedef struct orc_info {
    int32_t status;            /* 0, 6 (breakdown) or 9 (max iterations) - core.hxx:170,258,297 */
    int32_t iterations_needed; /* core.hxx:171,295 */
    int32_t iterations_run;    /* loop trips actually executed */
    int32_t probes;            /* residual probes executed */
    double  residuum_reached;  /* core.hxx:325 */
    double  flops_performed;   /* core.hxx:324 */
    double  last_max_bound2, last_target_bound2;
} orc_info_t;

/* tfqmrgpu::solve (core.hxx:20-335). A internal [nnzbA][2][LM(k)][LM(i)], B internal [nnzbB][2][LM][LN],
 * v3 float[nnzbX][2][LM][LN], X out internal [nnzbX][2][LM][LN]; status_out int8[nCols*LN] may be NULL. */
int  orc_solve_z(const orc_plan_t *plan, int LM, int LN, const double *A, const double *B,
                 const float *v3, double *X, double tolerance, int maxIterations, int mode,
                 orc_info_t *info, int8_t *status_out);
int  orc_solve_c(const orc_plan_t *plan, int LM, int LN, const float *A, const float *B,
                 const float *v3, float *X, double tolerance, int maxIterations, int mode,
                 orc_info_t *info, int8_t *status_out);

#ifdef __cplusplus
}
#endif
#endif /* TFQMR_ORACLE_H */
