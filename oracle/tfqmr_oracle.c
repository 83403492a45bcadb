/* TEST INFRASTRUCTURE (oracle). Not part of the product - see tfqmr_oracle.h.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared tfqmr_oracle.c -o liboracle.so -lm
 * (-ffp-contract=off keeps IEEE operation order so that ORC_MODE_CPUREF is bit-comparable with
 *  the reference's own CPU build, which is compiled without FMA as well.) */
#include "tfqmr_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- plan analysis: tfqmrgpu.cu:136-351 -------------------------------------------------------- */

/* bsr.hxx:27-39: linear search, first match wins */
static int find_in_array(int begin, int end, int value, const int32_t *array) {
    for (int ind = begin; ind < end; ++ind) if (value == array[ind]) return ind;
    return -1;
}

void orc_destroy_plan(orc_plan_t *p) {
    if (!p) return;
    free(p->starts); free(p->pairs); free(p->subset); free(p->colindx); free(p);
}

int orc_create_plan(int mb,
                    const int32_t *rpA, int nnzbA, const int32_t *ciA,
                    const int32_t *rpX, int nnzbX, const int32_t *ciX,
                    const int32_t *rpB, int nnzbB, const int32_t *ciB,
                    int indexOffset, orc_plan_t **plan) {
    *plan = NULL;
    /* tfqmrgpu.cu:166-172 (UNDOCUMENTED_ERROR = 14, line payload dropped) */
    if (mb < 1) return 14;
    if (nnzbX < 1) return 14;
    if (nnzbB > nnzbX) return 14;
    if ((long long)nnzbA > (long long)mb*mb) return 14;
    if (nnzbA != rpA[mb] - rpA[0]) return 14;
    if (nnzbX != rpX[mb] - rpX[0]) return 14;
    if (nnzbB != rpB[mb] - rpB[0]) return 14;

    orc_plan_t *p = calloc(1, sizeof(orc_plan_t));
    p->mb = mb; p->nnzbA = nnzbA; p->nnzbX = nnzbX; p->nnzbB = nnzbB;
    int const C0F1 = indexOffset;

    /* pairs / starts: tfqmrgpu.cu:183-219. Y has the pattern of X; A's given column order is kept */
    size_t cap = ((size_t)nnzbX*(size_t)(nnzbA > 0 ? nnzbA : 1))/(size_t)mb + 16, np = 0;
    p->pairs = malloc(2*cap*sizeof(uint32_t));
    p->starts = malloc(((size_t)nnzbX + 1)*sizeof(uint32_t));
    p->starts[0] = 0;
    for (int irow = 0; irow < mb; ++irow) {
        for (int inzy = rpX[irow] - C0F1; inzy < rpX[irow + 1] - C0F1; ++inzy) {
            int const jcol = ciX[inzy]; /* compared un-shifted, tfqmrgpu.cu:200 */
            for (int inza = rpA[irow] - C0F1; inza < rpA[irow + 1] - C0F1; ++inza) {
                int const krow = ciA[inza] - C0F1;
                int const inzx = find_in_array(rpX[krow] - C0F1, rpX[krow + 1] - C0F1, jcol, ciX);
                if (inzx >= 0) {
                    if (np == cap) { cap *= 2; p->pairs = realloc(p->pairs, 2*cap*sizeof(uint32_t)); }
                    p->pairs[2*np] = (uint32_t)inza; p->pairs[2*np + 1] = (uint32_t)inzx; ++np;
                }
            }
            p->starts[inzy + 1] = (uint32_t)np;
        }
    }
    p->nPairs = np;

    /* subset: tfqmrgpu.cu:233-251 */
    p->subset = malloc(((size_t)nnzbB + 1)*sizeof(uint32_t));
    for (int irow = 0; irow < mb; ++irow) {
        for (int inzb = rpB[irow] - C0F1; inzb < rpB[irow + 1] - C0F1; ++inzb) {
            int const inzx = find_in_array(rpX[irow] - C0F1, rpX[irow + 1] - C0F1, ciB[inzb], ciX);
            if (inzx < 0) { orc_destroy_plan(p); return 13 + 1000*irow; } /* B_IS_NOT_SUBSET_OF_X */
            p->subset[inzb] = (uint32_t)inzx;
        }
    }

    /* colindx / nCols: tfqmrgpu.cu:254-314 */
    int32_t min_c = 2147483647, max_c = -2147483647;
    for (int i = 0; i < nnzbX; ++i) { if (ciX[i] < min_c) min_c = ciX[i]; if (ciX[i] > max_c) max_c = ciX[i]; }
    long long const nc = 1LL + max_c - min_c;
    if (nc < 1) { orc_destroy_plan(p); return 14; }
    uint32_t *rows_per_col = calloc((size_t)nc, sizeof(uint32_t));
    int32_t *jc2jb = malloc((size_t)nc*sizeof(int32_t));
    for (int i = 0; i < nnzbX; ++i) ++rows_per_col[ciX[i] - min_c];
    uint32_t nb = 0;
    for (long long jc = 0; jc < nc; ++jc) jc2jb[jc] = rows_per_col[jc] ? (int32_t)(nb++) : -1;
    p->colindx = malloc((size_t)nnzbX*sizeof(uint16_t));
    for (int i = 0; i < nnzbX; ++i) p->colindx[i] = (uint16_t)jc2jb[ciX[i] - min_c];
    p->nCols = nb;

    /* every X column needs at least one B block: tfqmrgpu.cu:316-337 */
    uint32_t *rows_per_colB = calloc(nb ? nb : 1, sizeof(uint32_t));
    for (int ib = 0; ib < nnzbB; ++ib) ++rows_per_colB[jc2jb[ciX[p->subset[ib]] - min_c]];
    uint32_t nzero = 0;
    for (uint32_t jb = 0; jb < nb; ++jb) nzero += (rows_per_colB[jb] < 1);
    free(rows_per_col); free(jc2jb); free(rows_per_colB);
    if (nzero > 0) { orc_destroy_plan(p); return 11 + 1000*(int)nzero; } /* B_HAS_A_ZERO_COLUMN */

    *plan = p;
    return 0;
}

uint64_t orc_plan_array(const orc_plan_t *p, int kind, void *out) {
    switch (kind) {
        case 0: if (out) memcpy(out, p->starts, ((size_t)p->nnzbX + 1)*4); return (uint64_t)p->nnzbX + 1;
        case 1: if (out) memcpy(out, p->pairs, p->nPairs*8); return 2*p->nPairs;
        case 2: if (out) memcpy(out, p->subset, (size_t)p->nnzbB*4); return (uint64_t)p->nnzbB;
        case 3: if (out) memcpy(out, p->colindx, (size_t)p->nnzbX*2); return (uint64_t)p->nnzbX;
    }
    return 0;
}

/* allowed_block_sizes.h:4-18 */
int orc_block_size_allowed(int ldA, int ldB) {
    static const int list[15][2] = {{4,4},{4,5},{4,8},{4,32},{8,8},{8,9},{8,10},{8,32},{8,64},
                                    {16,16},{16,32},{16,64},{32,32},{32,64},{64,64}};
    for (int q = 0; q < 15; ++q) if (list[q][0] == ldA && list[q][1] == ldB) return 1;
    return 0;
}

/* memcount pass: core.hxx:45-99 + blocksparse.hxx:46-51 + util.hxx:56-86 (256-byte bump allocator) */
static uint64_t take(uint64_t *buf, uint64_t bytes) {
    uint64_t const mask = 255;
    if (*buf & mask) *buf = ((*buf >> 8) + 1) << 8;
    uint64_t const at = *buf;
    *buf += bytes;
    if (*buf & mask) *buf = ((*buf >> 8) + 1) << 8;
    return at;
}
uint64_t orc_ref_buffer_size(const orc_plan_t *p, int LM, int LN, int is_double) {
    uint64_t const s = is_double ? 8 : 4, nX = (uint64_t)p->nnzbX, nB = (uint64_t)p->nnzbB, nC = p->nCols;
    uint64_t buf = 0;
    for (int v = 0; v < 7; ++v) take(&buf, nX*2*LM*LN*s);  /* v1, v4..v9 */
    take(&buf, nX*2*LM*LN*4);                              /* v3 float */
    take(&buf, nB*2*LM*LN*s);                              /* v2 = B */
    for (int q = 0; q < 5; ++q) take(&buf, nC*2*LN*s);     /* rho alfa beta c67 eta */
    unsigned l2 = 0; { uint32_t n = (uint32_t)(nX - 1); while (n > 0) { ++l2; n >>= 1; } } /* highestbit(nnzbX-1)+1 */
    uint64_t const np2 = 1ull << l2;
    take(&buf, np2*nC*2*LN*8);                             /* zvv */
    take(&buf, np2*nC*1*LN*8);                             /* dvv */
    take(&buf, nC*LN*8); take(&buf, nC*LN*8);              /* tau, var */
    take(&buf, nX*2); take(&buf, nB*4); take(&buf, nC*LN); /* colindx, subset, status */
    take(&buf, (nX + 1)*4); take(&buf, p->nPairs*8);       /* starts, pairs */
    take(&buf, (uint64_t)p->nnzbA*2*LM*LM*s);              /* A */
    return buf + 256;
}

/* linalg.hxx:799-802 */
void orc_v3_glibc(float *v3, size_t n) {
    srand(1);
    float const denom = 1./RAND_MAX;
    for (size_t i = 0; i < n; ++i) v3[i] = rand()*denom;
}

#define REAL float
#define SFX(x) x##_f
#include "tfqmr_oracle_typed.inc"
#undef REAL
#undef SFX

#define REAL double
#define SFX(x) x##_d
#include "tfqmr_oracle_typed.inc"
#undef REAL
#undef SFX

int orc_solve_z(const orc_plan_t *plan, int LM, int LN, const double *A, const double *B, const float *v3,
                double *X, double tolerance, int maxIterations, int mode, orc_info_t *info, int8_t *status_out) {
    return orc_solve_impl_d(plan, LM, LN, A, B, v3, X, tolerance, maxIterations, mode, info, status_out);
}
int orc_solve_c(const orc_plan_t *plan, int LM, int LN, const float *A, const float *B, const float *v3,
                float *X, double tolerance, int maxIterations, int mode, orc_info_t *info, int8_t *status_out) {
    return orc_solve_impl_f(plan, LM, LN, A, B, v3, X, tolerance, maxIterations, mode, info, status_out);
}
