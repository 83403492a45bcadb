// Internal declarations of the B200-native tfQMR library (not installed).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/tfqmrgpu_b200_ext.h"

namespace tfq {

// ---- error helpers -------------------------------------------------------------------------------
inline tfqmrgpuStatus_t err_line(tfqmrgpuStatus_t code, int line) { return code + TFQMRGPU_CODE_LINE*(line % 10000); }
inline tfqmrgpuStatus_t err_char(tfqmrgpuStatus_t code, char c, int line) {
    return code + TFQMRGPU_CODE_CHAR*int((unsigned char)c) + TFQMRGPU_CODE_LINE*(line % 10000);
}
#define TFQ_ERR(code)        ::tfq::err_line(code, __LINE__)
#define TFQ_ERRC(code, ch)   ::tfq::err_char(code, ch, __LINE__)
// CUDA runtime failure -> LAUNCH_FAILED with the source line; never exits the process
#define TFQ_CUDA(call) do { cudaError_t e_ = (call); if (cudaSuccess != e_) { \
        if (::tfq::verbosity() > 0) std::fprintf(stderr, "tfQMRgpu: CUDA error \"%s\" at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
        return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED); } } while (0)

int  verbosity();
void set_verbosity(int level);

constexpr uint32_t kNoBlock = 0xffffffffu;   // "no such block" in SpMM entry tables
constexpr int      kPartD   = 4;             // doubles per (tile, j) in the partial-sum scratch

// ---- device-resident solver control (one per plan, lives in the workspace) ------------------------
enum : int { STATE_RUN = 0, STATE_PROBE = 1, STATE_DONE = 2 };
struct Control {
    int      state;              // STATE_*
    int      iteration;          // completed tfQMR iterations (core.hxx:180)
    int      max_iterations;
    int      probes;             // residual probes executed (core.hxx:263)
    int      result;             // 0 converged, 9 max iterations, 6 breakdown (core.hxx:170,258,297)
    int      iterations_needed;  // core.hxx:171,295
    unsigned cols_done;          // ticket of the cross-column finalisation
    int      freeze;             // tfqmrgpux_bsrsv_setEarlyFreeze: a right-hand side whose true residual passed a probe keeps its X (status 2)
    double   tol2;               // core.hxx:129
    double   target_bound2;      // core.hxx:130,290
    double   residual2_reached;  // core.hxx:131,287
    double   max_bound2;         // core.hxx:252
    double   min_norm2, max_norm2; // extrema of |b|^2 (core.hxx:161-167)
};

struct Handle { cudaStream_t stream = nullptr; };

// Column-sharded runs (several GPUs, each solving a range of the right-hand-side block columns): the per-iteration exchange of
// the convergence monitors that keeps the reference's GLOBAL iteration / probe rule (vecops.cu: decide_kernel).
struct Exchange {
    int nshards = 1, shard = 0;
    long long nrhs_global = 0;
    double *slots = nullptr;          // device-accessible [2 kinds][2 parities][nshards][4] doubles; nullptr: a single GPU decides alone
    int parity = 0;                   // of the iteration body being enqueued
    tfqmrgpuxExchange_t hook = nullptr;   // multi-process runs: all-gathers the slots of (kind, parity) on the solver's stream
    void *hook_ctx = nullptr;
};

// one tile of the column-sorted vectors: blocks [b0, b1) all belong to block column `col`
struct Tile { uint32_t col, b0, b1, pad; };

struct Plan {
    // ---- problem sizes -------------------------------------------------------------------------
    int mb = 0, nnzbA = 0, nnzbX = 0, nnzbB = 0, indexOffset = 0;
    uint32_t nCols = 0;
    uint64_t nPairs = 0;
    int LM = 0, LN = 0;
    char precision = 0;               // 'c' | 'z' | 'm' once bufferSize has been called
    int  maxColsPerRow = 1;

    // ---- device index lists owned by the plan (cudaMalloc) -------------------------------------
    // reference-format lists (tfqmrgpu_plan.hxx:20-50), built on the device
    uint32_t *d_starts = nullptr, *d_pairs = nullptr, *d_subset = nullptr;
    uint16_t *d_colindx = nullptr;
    // column-sorted storage order of X-shaped vectors
    uint32_t *d_perm = nullptr;       // caller X index -> storage index
    uint32_t *d_iperm = nullptr;      // storage index -> caller X index
    uint32_t *d_bpos = nullptr;       // storage index of the X block under each B block
    uint32_t *d_blockcol = nullptr;   // block column of every storage-ordered X block
    int32_t  *d_rowptrA = nullptr;    // zero-based copy of bsrRowPtrA (row scales of the A operand)
    std::vector<uint32_t> h_colstart; // [nCols+1] first storage block of every block column
    std::vector<int32_t>  h_rowptrX;  // zero-based copy of bsrRowPtrX
    // zero-based copies of the other index arrays (row ranges of A for chunked / multi-device uploads; sub-plans of column shards)
    std::vector<int32_t>  h_rpA, h_ciA, h_ciX, h_rpB, h_ciB;
    // block-size dependent: vector tiles
    Tile     *d_tiles = nullptr;      uint32_t nTiles = 0;
    uint32_t *d_coltile = nullptr;    // [nCols+1] first tile of every block column
    // block-size dependent: SpMM units (one CTA each): a block row times <= gmax block columns
    uint32_t nUnits = 0, gmax = 1; uint64_t nEntries = 0;
    bool use_tc16 = false;            // block-sparse product on the tensor cores, fp16 operand pairs (spmm_tc16.cu, xop.cu)
    bool tc_planar = false;           // ... in the planar form (spmm_tc16p.cu: four real products, the fast kernel for short rows)
    bool use_dmma = false;            // complex fp64 product on the FP64 tensor pipe (spmm_dmma.cu)
    Exchange exch;
    struct MultiPlan *multi = nullptr;       // in-process multi-GPU plan (multi.cu): this Plan is then only the global analysis
    bool lean_vectors = false;               // workspace layout without v4..v7 (the fp64 side of a mixed plan)
    struct MixedPlan *mixed = nullptr;       // precision 'm' (mixed.cu): this Plan is the fp64 side (precision == 'z') of a refinement around an fp32 plan
    int initial_guess = 0;                  // opt-in: solve starts from v1 instead of zero (tfqmrgpux_bsrsv_setInitialGuess)
    double *d_guess_scratch = nullptr; double guess_flops = 0;
    int early_freeze = 0;                    // opt-in: freeze converged right-hand sides at the probes (tfqmrgpux_bsrsv_setEarlyFreeze)
    unsigned *d_resident_bar = nullptr;      // grid barrier of the resident solver (resident.cu)
    int resident_fits = -1;                  // resident solver: does this configured plan qualify (-1: not yet asked)
    unsigned long long *d_resident_trace = nullptr;   // dev: time breakdown of the resident solver (TFQMRGPU_RESIDENT_TRACE)
    Tile *d_res_tiles = nullptr; uint32_t *d_res_coltile = nullptr; double *d_res_part = nullptr;   // resident solver: its own tiles
    uint32_t *d_unit_of_block = nullptr;     // resident solver: storage index of a Y block -> its unit (plans with gmax == 1)
    size_t tile_blocks = 0;                  // X blocks per vector tile chosen by plan_configure (largest over the block columns)
    int    max_cols_hint = 0;               // ... and choose the product kernel by the unsharded plan's block columns per row
    size_t tile_blocks_hint = 0;             // > 0: forced tile size (dev / experiments); the default rule is per block column
    tfqmrgpuxOperator_t user_op = nullptr;   // user-defined operator instead of the block-sparse product (ext header)
    void *user_ctx = nullptr;
    tfqmrgpuxOperator_t precond = nullptr;   // right preconditioner z = P*x as a C callback (tfqmrgpux_bsrsv_setPreconditioner): the solver iterates on A*P
    void *precond_ctx = nullptr;
    char *d_precond_tmp = nullptr;           // one X-shaped vector owned by the plan: P*v6 / P*v1 (the reference's vP, core.hxx:57)
    bool use_small = false;           // LM <= 8: register-staged batches of entries instead of the bulk-copy ring (spmm.cu)
    uint32_t *d_unit_e0 = nullptr;    // [nUnits+1] first entry of every unit
    uint32_t *d_unit_y = nullptr;     // [nUnits*gmax] storage index of the unit's Y blocks (kNoBlock = none)
    uint32_t *d_ent_a = nullptr;      // [nEntries] A block of the entry
    uint32_t *d_ent_x = nullptr;      // [nEntries*gmax] storage index of X blocks (kNoBlock = structural zero)
    // fp16-pair tensor-core product: units dealt to `tc_grid` persistent CTAs and stored CTA by CTA
    uint32_t *d_cta_u0 = nullptr;     // [tc_grid+1] first unit of every CTA
    uint32_t *d_unit_row = nullptr;   // [nUnits] block row of the unit (scale of the A operand)
    uint32_t tc_grid = 0;
    int      tc_seg = 16;             // entries per accumulation segment

    // ---- caller-owned workspace ------------------------------------------------------------------
    char  *pBuffer = nullptr;
    size_t bufferBytes = 0;
    // byte offsets inside the workspace
    size_t off_v[10] = {0};           // off_v[1], [3..9]: v1 (X), v3 (float), v4..v9
    size_t off_B = 0, off_A = 0, off_zero = 0;
    size_t off_rho = 0, off_alfa = 0, off_beta = 0, off_c67 = 0, off_eta = 0;
    size_t off_tau = 0, off_var = 0, off_invBn2 = 0, off_status = 0, off_snap = 0;
    size_t off_part = 0, off_colmon = 0, off_ticket = 0, off_ctl = 0;
    // fp16-pair operands (xop.cu): X operand, column scales and inverses, row scale inverses of A (directly behind the A
    // blocks, inside the 'A' window), block maxima and row scales of A, tile maxima of the X operand
    size_t off_xop = 0, off_xs = 0, off_xsinv = 0, off_ainv = 0, off_ablkmax = 0, off_arowscale = 0, off_xpart = 0, off_mx = 0;
    size_t vecBytes = 0;              // bytes of one X-shaped vector

    // ---- host side ---------------------------------------------------------------------------
    Control *h_ctl = nullptr;         // pinned ring for control read-backs
    cudaEvent_t ev[8] = {nullptr};
    bool v3_ready = false, solved = false;
    bool xop_of_x = false;            // the tensor-core operand currently holds v1 (made by setMatrix('X'); a solve overwrites it)
    bool configured = false;          // bufferSize succeeded: tiles, units and workspace offsets match LM, LN, precision
    // one tfQMR iteration body (8 iteration kernels + 3 probe kernels) as an instantiated CUDA graph; rebuilt when the
    // workspace or the block configuration changes
    cudaGraphExec_t body_exec = nullptr;
    cudaStream_t    capture_stream = nullptr;
    cudaStream_t    copy_stream = nullptr;     // uploads of large operands, overlapped chunk-wise with their layout conversion
    cudaEvent_t     chunk_ev[3] = {nullptr};   // order the chunk uploads and their conversions (owned by the plan: no leak on error paths)

    // ---- stats (tfqmrgpu_plan.hxx:41-45) -------------------------------------------------------
    double residuum_reached = 0, flops_performed = -1, flops_performed_all = 0;
    int    iterations_needed = -1;
    int    iterations_run = 0;        // iterations the last solve really ran (iterations_needed keeps the reference's meaning)
    double stat_probes = 0, stat_launches = 0, stat_bodies = 0, stat_ms = 0, stat_bound2 = 0, stat_target2 = 0;
    // optional device-side profile of the last solve (tfqmrgpux_bsrsv_setProfiling)
    bool   profile = false;
    std::vector<cudaEvent_t> prof_ev;  // [0],[1] bracket the solve; 4 per iteration body bracket its two A*v6 products
    double prof_solve_ms = 0, prof_spmm_ms = 0, prof_spmm_launches = 0, prof_iterations = 0;
};

// ---- plan analysis (plan.cu) ---------------------------------------------------------------------
tfqmrgpuStatus_t plan_analyse(Plan &p, cudaStream_t stream,
    int32_t const *rpA, int32_t const *ciA, int32_t const *rpX, int32_t const *ciX,
    int32_t const *rpB, int32_t const *ciB, int echo);
tfqmrgpuStatus_t plan_configure(Plan &p, cudaStream_t stream, int LM, int LN, char precision); // tiles, units, offsets
void plan_release(Plan &p);
void plan_drop_graph(Plan &p);   // forget the captured iteration body
// the resident solver (resident.cu): a whole solve of a small system in one cooperative launch
bool resident_supported(Plan &p);
tfqmrgpuStatus_t launch_resident_solve(Plan &p, cudaStream_t stream, int maxIterations);
size_t plan_tile_blocks(size_t nColBlocks, size_t blockBytes, int nsm);   // X blocks per vector tile for a block column of that length

// ---- several devices in one process (multi.cu) ---------------------------------------------------------
tfqmrgpuStatus_t multi_set_devices(Plan &p, int nDevices, int const *devices);
int  multi_get_devices(Plan const &p, int *devices, int arrayLength);
void multi_destroy(Plan &p);
tfqmrgpuStatus_t multi_buffer_size(Plan &p, int LM, int LN, char prec, size_t *bytes);
tfqmrgpuStatus_t multi_set_buffer(Plan &p, void *pBuffer);
tfqmrgpuStatus_t multi_set_matrix(Plan &p, char v, void const *val, char precision, char transposition, tfqmrgpuDataLayout_t layout, bool trans, double scal_imag);
tfqmrgpuStatus_t multi_solve(Plan &p, double tolerance, int maxIterations);
tfqmrgpuStatus_t multi_gather_x(Plan &p, cudaStream_t homeStream);
size_t multi_off_gx(Plan const &p);
size_t multi_off_scratch(Plan const &p);
tfqmrgpuStatus_t multi_rhs_status(Plan &p, int8_t *statusHost);
void multi_set_early_freeze(Plan &p);

// ---- precision 'm': fp64 refinement around the fp32 solver (mixed.cu) -------------------------------------
tfqmrgpuStatus_t mixed_buffer_size(Plan &p, cudaStream_t stream, int LM, int LN, size_t *bytes);
tfqmrgpuStatus_t mixed_set_buffer(Plan &p, cudaStream_t stream);      // after the fp64 side has been attached
tfqmrgpuStatus_t mixed_after_set_a(Plan &p, cudaStream_t stream);     // fp32 operand from the uploaded fp64 operator
tfqmrgpuStatus_t mixed_solve(Plan &p, cudaStream_t stream, double tolerance, int maxIterations);
void  mixed_destroy(Plan &p);
Plan* mixed_inner(Plan const &p);
int   mixed_passes(Plan const &p);
tfqmrgpuStatus_t fill_v3(Plan &p, cudaStream_t stream);               // api.cu: the reference's cuRAND shadow vector

// ---- kernels' host launchers --------------------------------------------------------------------
// block-sparse product y = A*x on storage-ordered vectors; gate: run only if ctl->state == expect (expect < 0: always)
tfqmrgpuStatus_t launch_spmm(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream);
// fp16-pair tensor-core variant (spmm_tc16.cu, the default for complex fp32 with LM, LN in {16, 32, 64}) and its operands (xop.cu)
bool spmm_tc16_supported(int LM, int LN, char precision, int level);
int  spmm_tc16_columns_per_unit(int LM, int LN);
int  spmm_tc16_default_segment(int LM);
tfqmrgpuStatus_t launch_spmm_tc16(Plan const &p, void *y, int expect, cudaStream_t stream);
int  spmm_tc16p_default_segment(int LM);
tfqmrgpuStatus_t launch_spmm_tc16p(Plan const &p, void *y, int expect, cudaStream_t stream);   // planar form
// y = A*x where the X operand of x already exists (written by launch_vecop_xop)
tfqmrgpuStatus_t launch_spmm_operand_ready(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream);
tfqmrgpuStatus_t launch_xop(Plan const &p, void const *x, int expect, cudaStream_t stream);          // X operand from a vector
tfqmrgpuStatus_t launch_aop_blockmax(Plan const &p, uint32_t b0, uint32_t nb, cudaStream_t stream);  // per uploaded chunk of A
tfqmrgpuStatus_t launch_aop_convert(Plan const &p, cudaStream_t stream);                            // after the last chunk
tfqmrgpuStatus_t launch_aop_convert_rows(Plan const &p, int row0, int row1, cudaStream_t stream);    // the same for a range of block rows
// DMMA variant (spmm_dmma.cu): complex fp64, LM and LN in {16, 32, 64}; also switched off by TFQMRGPU_TENSOR=0
bool spmm_dmma_supported(int LM, int LN, char precision);
int  spmm_dmma_columns_per_unit(int LM, int LN);
tfqmrgpuStatus_t launch_spmm_dmma(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream);
// fused vector algebra, see vecops.cu
enum VecOp : int { OP_INIT = 0, OP_K1, OP_E1, OP_K2, OP_K3, OP_E2, OP_K4, OP_N3, OP_COUNT };
tfqmrgpuStatus_t launch_vecop(Plan const &p, int op, cudaStream_t stream);
inline double* exchange_slots(Plan const &p, int kind) { return p.exch.slots + size_t((kind*2 + p.exch.parity)*p.exch.nshards)*4; }
tfqmrgpuStatus_t launch_decide(Plan const &p, int kind, cudaStream_t stream);
// OP_K1 / OP_K3 that also emit the X operand of the fp16-pair tensor-core product from the v6 they write
tfqmrgpuStatus_t launch_vecop_xop(Plan const &p, int op, cudaStream_t stream);
tfqmrgpuStatus_t launch_unit_rhs(Plan const &p, cudaStream_t stream);   // B := unit blocks (the reference's rhs_trivial right-hand sides)
// v[bpos[b]] += scal * B[b]   (linalg.hxx:383-428)
tfqmrgpuStatus_t launch_add_rhs(Plan const &p, void *v, double scal, int expect, cudaStream_t stream);
// host layout <-> internal layout (layout.cu)
tfqmrgpuStatus_t convert_inplace(Plan const &p, void *blocks, uint32_t nnzb, int rows, int cols, bool is_double,
                                 int layout, bool trans, double scal_imag, cudaStream_t stream);
tfqmrgpuStatus_t convert_permuted(Plan const &p, void *dst, void const *src, uint32_t nnzb, int rows, int cols,
                                  bool is_double, int layout, bool trans, double scal_imag, bool to_internal,
                                  cudaStream_t stream);
tfqmrgpuStatus_t permute_v3(Plan const &p, float *dst_storage, float const *src_caller, bool to_storage, cudaStream_t stream);

// ---- solver driver (solver.cu) --------------------------------------------------------------------
tfqmrgpuStatus_t solve(Plan &p, cudaStream_t stream, double tolerance, int maxIterations);
// the pieces of solve(), for the in-process multi-GPU driver (multi.cu) which interleaves them over the shards
tfqmrgpuStatus_t solve_begin(Plan &p, cudaStream_t stream, double tolerance, int maxIterations);   // initial state, v5 := b, INIT
tfqmrgpuStatus_t enqueue_iteration(Plan &p, cudaStream_t stream, cudaEvent_t const *events);      // K1 .. K4
tfqmrgpuStatus_t enqueue_probe(Plan &p, cudaStream_t stream);                                     // A*v1 - b, N3
tfqmrgpuStatus_t solve_finish(Plan &p, Control const &fin, int bodies, double launches);           // bookkeeping from the final control block

// register tile of the SpMM kernel (shared by plan.cu's unit builder and spmm.cu)
constexpr int spmm_tj(bool is_double, int LN) {
    return (LN % 4 == 0) ? (is_double ? 2 : 4) : ((LN % 2 == 0) ? 2 : 1);
}
constexpr int spmm_ti(bool is_double, int LM, int LN) {
    return ((LM/4)*(LN/spmm_tj(is_double, LN)) > 256) ? 8 : 4;
}

bool block_size_allowed(int lm, int ln);
extern const int kAllowedBlockSizes[15][2];

// Dynamic shared memory above 48 KiB is opted in per kernel AND per device: remember the largest request per device
// (one process may drive several GPUs through several handles).
constexpr int kMaxDevices = 64;
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t (&configured)[kMaxDevices]) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (cudaSuccess != err) return err;
    if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (bytes > configured[dev]) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
        if (cudaSuccess == err) configured[dev] = bytes;
    }
    return err;
}

template <typename T> inline T* ws(Plan const &p, size_t off) { return reinterpret_cast<T*>(p.pBuffer + off); }

} // namespace tfq
