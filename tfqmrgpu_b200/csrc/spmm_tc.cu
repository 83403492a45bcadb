// Block-sparse product  Y = A * X  for complex fp32 on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same role as spmm.cu (the reference's blocksparse_action_t::multiply + gemmNxNf,
// tfqmrgpu_blocksparse.hxx:71-199, tfqmrgpu_blockmult.hxx:10-93), used where the SIMT kernel is bound by
// the FP32 pipe (LM = 32: 55 flop per HBM byte).  The complex block product is
// mapped to ONE real GEMM per (block row, A block):
//
//     D[m][n] += sum_k  Xop[m][k] * Aop[n][k]        m = (g, Re|Im of X, j)   -> 128 rows  (G*2*LN)
//                                                    n = (Re|Im of A, i)      -> 2*LM columns
//     Xop[(g,cx,j)][k] = X_g[cx][k][j]               Aop[(ca,i)][k] = A[ca][k][i]
//
// so D holds the four real products Xr*Ar, Xr*Ai, Xi*Ar, Xi*Ai and no tensor flop is wasted; the epilogue
// combines them: Yr = XrAr - XiAi, Yi = XrAi + XiAr.   fp32 accuracy comes from the 3xTF32 split
// x = hi + lo (both rounded to TF32, exact operands):  D += Xhi*Ahi + Xlo*Ahi + Xhi*Alo  (the dropped
// lo*lo term is 2^-24 relative).
//
// Data movement per CTA (one "unit": a block row times G block columns):
//   * X operand: global -> registers (coalesced over j) -> split -> tcgen05.st into TENSOR MEMORY; the MMA
//     reads its 128 x K operand from TMEM, so X never touches shared memory;
//   * A operand: setMatrix('A') stores these blocks in HBM already in the canonical K-major no-swizzle layout
//     of the tcgen05 shared-memory descriptor, so ONE bulk copy (cp.async.bulk, 8 KiB) per block into a
//     6-deep ring delivers the hi operand as is (the tensor core truncates fp32 to TF32); the threads only
//     derive lo = a - trunc(a) into a second buffer;
//   * accumulator D: 128 lanes x 2*LM columns of TMEM; read back once per unit with tcgen05.ld.
// Two operand stages (TMEM columns + lo buffer) are recycled through mbarriers signalled by tcgen05.commit;
// eight converter warps prepare the operands and never synchronise with each other inside the loop (they only
// arrive on the stage's mbarrier), a ninth warp issues the bulk copies and the MMAs from one lane; the X loads
// of entry e+1 are in flight while entry e is split, and two CTAs per SM overlap each other's
// prologue/epilogue.
#include "tfq_internal.hpp"
#include <cstdlib>
#include <algorithm>

namespace tfq {

namespace {

constexpr int kConvWarps = 8, kConvThreads = 32*kConvWarps;   // converter warps; warp 8 issues the MMAs, warp 9 the bulk copies of A
constexpr int kMmaWarp = kConvWarps, kCopyWarp = kConvWarps + 1;
constexpr int kTcThreads = kConvThreads + 64;
// Tensor memory per CTA: [0, 2N): accumulator (main sum | correction sum), then two X operand stages of 2*LM columns (hi | lo).
// LM = 16: 64 + 2*32 = 128 columns, LM = 32: 128 + 2*64 = 256 (two CTAs per SM), LM = 64: 256 + 2*128 = 512 (one CTA per SM)

struct TcArgs {
    float *y; float const *x; float const *A;
    uint32_t const *unit_e0, *unit_y, *ent_a, *ent_x;
    Control const *ctl; int expect; int gstride; uint32_t nUnits;
    int chain;        // entries per accumulation pass
};

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA engine without a tensor map), completion counted on an mbarrier, with an L2 eviction policy
// (createpolicy): the A stream is read once and must not displace the X blocks, which are re-read from L2 (27x for the stencil):
// measured DRAM reads per launch 9.31 -> 8.40 GB, 2.77 -> 2.72 ms
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, void const *src_gmem, unsigned bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return 0 != ok;
}
// bounded: a tensor-core pipeline that never signals is a bug and must fail loudly (launch error), not hang
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
    return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], TF32 inputs, fp32 accumulation; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t const *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, uint32_t const *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 3xTF32 operand split  x ~ hi + lo.   The tensor core TRUNCATES an fp32 bit pattern to TF32 (measured on B200:
// feeding the raw word or the word with its low 13 bits cleared gives bit-identical products), so
//   * for the A blocks hi is the RAW word as the bulk copy delivered it, and lo = a - trunc(a) (exact in fp32);
//   * for X, which passes through registers anyway, hi is rounded to NEAREST (|x - hi| <= 2^-12 |x|) and
//     lo = x - hi (exact); the hardware's truncation of lo costs at most 2^-22 |x|.
__device__ __forceinline__ void split_rn(float v, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ float lo_trunc(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// shared-memory matrix descriptor, no swizzle.  K-major operand: core matrix = 8 rows (m or n) of 16 bytes (4 TF32 k
// values); LBO = byte stride between core matrices along K, SBO = byte stride between 8-row groups along M/N
__device__ __forceinline__ uint64_t smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3fff);
    d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= uint64_t(1) << 46;                          // descriptor version of sm_100
    return d;                                        // base offset 0, layout type 0 = no swizzle
}


// dev-only timing ablations (scripts/dev_ablate.sh builds variants; results are WRONG with any bit set):
//   1 no lo(A) LDS/STS, 2 no split + tcgen05.st, 4 no X loads, 8 no MMAs (commit only), 16 no A bulk copies, 32 no Y stores,
//   128 no X path for every third entry (what sharing one X block between two block rows would save on the 27-point stencil)
#ifndef TFQ_TC_ABLATE
#define TFQ_TC_ABLATE 0
#endif

// dev-only timeline trace of one unit of CTA 0 (scripts/dev_tc_trace.py): clock64 stamps per entry and role
#ifdef TFQ_TC_TRACE
__device__ long long g_tc_trace[64*24];
#define TFQ_TRACE(e, slot) do { if (trace_on) g_tc_trace[(e)*24 + (slot)] = clock64(); } while (0)
#else
#define TFQ_TRACE(e, slot) do { } while (0)
#endif

template <int LM> struct TcShape {
    static constexpr int ring = (64 == LM) ? 3 : 6;                    // A blocks in flight per CTA (bulk-copy ring)
    static constexpr int ctas = (64 == LM) ? 1 : 2;                    // resident CTAs per SM (TMEM columns)
    static constexpr uint32_t acc_cols = 4*LM, stage_cols = 2*LM;      // accumulator (2N), one X stage (hi | lo)
    static constexpr uint32_t tmem_cols = (acc_cols + 2*stage_cols <= 128) ? 128 : ((acc_cols + 2*stage_cols <= 256) ? 256 : 512);
};

template <int LM, int LN>
__global__ void __launch_bounds__(kTcThreads, TcShape<LM>::ctas)
spmm_tc_kernel(TcArgs const a)
{
    static_assert(LM == 16 || LM == 32 || LM == 64, "k range per thread is LM/2 (tcgen05.st .x8 / .x16)");
    static_assert(LN == 16 || LN == 32 || LN == 64, "128 MMA rows = G * 2 * LN");
    constexpr int kRingA = TcShape<LM>::ring;
    constexpr uint32_t kTmemCols = TcShape<LM>::tmem_cols, kTmemStage0 = TcShape<LM>::acc_cols, kStageCols = TcShape<LM>::stage_cols;
    constexpr int KH = LM/2;              // k values per converter thread (two warps share a lane quarter)
    int const kChain = a.chain;           // entries per accumulation pass: LM/8 MMA steps each; default 112 steps per pass (56 at LM = 64)
    constexpr int G  = 64/LN;             // block columns per unit
    constexpr int N  = 2*LM;              // MMA N: (Re|Im of A, i)
    constexpr int KS = LM/8;              // k-steps of 8 (TF32) per entry
    constexpr int ABLK = 2*LM*LM;         // floats of one A block
    constexpr int XBLK = 2*LM*LN;
    // B operand (the A block) in shared memory: K-major, no swizzle, [k/4][n][k%4] - the layout setMatrix('A') stores
    // in HBM for these plans (layout.cu), so one bulk copy per block lands it ready for the MMA.
    // (MN-major TF32 operands return zeros on sm_100a - measured, see DESIGN.md.)
    // A ring slot: per k-quad a slab [hi: N rows x 16 B][lo: N rows x 16 B].  hi arrives by bulk copy (one copy per
    // k-quad slab), lo is written next to it, so ONE descriptor with 2N rows covers [Ahi ; Alo]:
    //   Xhi * [Ahi ; Alo]  -> columns [0,N) (main sum) and [N,2N) (correction sum) of the accumulator, one MMA of N' = 2N
    //   Xlo *  Ahi         -> columns [N,2N), one MMA of N' = N
    // i.e. 2 MMAs per k-step instead of 3 (the per-instruction overhead dominates at these small N).
    constexpr uint32_t SLAB = N*16;       // bytes of one hi (or lo) k-quad slab
    constexpr uint32_t KSB = 4*SLAB;      // bytes of one k-step: 2 k-quads x (hi + lo)
    constexpr uint32_t LBO = 2*SLAB;      // between the two k-quads of a k-step
    constexpr uint32_t SBO = 128;
    constexpr uint32_t SLOT = 2*ABLK*4;   // bytes of a ring slot (hi + lo)
    // instruction descriptors: D fp32, A/B TF32, both K-major, M = 128, N' = 2N and N
    constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(128 >> 4) << 24);
    constexpr uint32_t IDESC_2N = IDESC_BASE | (uint32_t((2*N) >> 3) << 17);
    constexpr uint32_t IDESC_N  = IDESC_BASE | (uint32_t(N >> 3) << 17);

    if (a.expect >= 0 && a.ctl->state != a.expect) return; // device-resident solver control

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *const bar_mma   = reinterpret_cast<uint64_t*>(smem_raw);      // [2]  MMAs of a stage have completed
    uint64_t *const bar_ready = bar_mma + 2;                                // [2]  a stage's operands are in place
    uint64_t *const bar_a     = bar_mma + 4;                                // [kRingA] raw A block has landed
    uint64_t *const bar_free  = bar_a + kRingA;                             // [kRingA] the MMAs that read a ring slot have completed
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 128);
    uint32_t *const s_y = reinterpret_cast<uint32_t*>(smem_raw + 160);      // [G]
    unsigned char *const ring = smem_raw + 1024;                            // [kRingA][SLOT] A operands, hi and lo slabs interleaved

    int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    int const gs = a.gstride;

    if (kMmaWarp == w) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t const tmem_base = *tmem_slot;

    // persistent: two resident CTAs per SM walk the units; TMEM is allocated once, the barriers are re-armed per unit
    // (a gated-off launch of this kernel then costs a handful of CTAs instead of one per unit)
    for (uint32_t u = blockIdx.x; u < a.nUnits; u += gridDim.x) {
    uint32_t const e0u = a.unit_e0[u];
    int const nEu = int(a.unit_e0[u + 1] - e0u);
    // Accumulation chains are cut into passes of at most kChain entries: the tensor core truncates the fp32 accumulator
    // once per MMA, and the attainable tfQMR residual was measured to follow the chain length (LM = 64: 216 MMA steps in one
    // chain 2.3-4x the SIMT floor, 112 steps 1.3-1.7x, 56 steps 0.6-1.0x; LM = 32: 108 steps 0.8x).  A later pass adds to
    // the Y it finds.
    int const nPasses = (nEu + kChain - 1)/kChain, perPass = (nPasses > 0) ? (nEu + nPasses - 1)/nPasses : 0;
    for (int pass = 0; pass < (nPasses > 0 ? nPasses : 1); ++pass) {
    uint32_t const e0 = e0u + uint32_t(pass*perPass);
    int const nE = (nEu - pass*perPass < perPass) ? (nEu - pass*perPass) : perPass;
    bool const first_pass = (0 == pass);
#ifdef TFQ_TC_TRACE
    bool const trace_on = (0 == blockIdx.x) && (u == 3*gridDim.x) && (0 == lane) && (0 == w || 7 == w || kMmaWarp == w);
    int const tslot0 = (0 == w) ? 0 : ((7 == w) ? 7 : 14);
#endif

    if (tid < G) s_y[tid] = (tid < gs) ? a.unit_y[size_t(u)*gs + tid] : kNoBlock;
    if (0 == tid) {
        mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1);
        mbar_init(&bar_ready[0], kConvWarps); mbar_init(&bar_ready[1], kConvWarps);
        #pragma unroll
        for (int r = 0; r < kRingA; ++r) { mbar_init(&bar_a[r], 1); mbar_init(&bar_free[r], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (kCopyWarp == w) {
        // ================= copy warp: bulk copies of the A blocks into the ring ======================================
        // Entry indices are read by the whole warp, 32 entries at a time (lane l holds entry base + l): the elected lane then
        // never waits on an index load in front of a copy (that wait was ~450 cycles per entry when the MMA warp did this).
        // (Requesting the entry's X blocks into L2 here - cp.async.bulk.prefetch.L2 - was measured: no gain, they hit L2 anyway;
        // deriving the lo operands in this warp (on the stage barrier, or decoupled on a barrier per slot) or in two more warps
        // of their own: 3.03-3.10 ms where this version takes 2.82 - the converter warps keep that work.)
        // A slot is free again when the MMAs that read it have completed: its own barrier, because a parity wait cannot look
        // two phases back and this warp may fall behind the stage barriers.
        uint32_t const leader = elect_one_sync();
        uint64_t const stream_once = policy_evict_first();
        uint32_t ia_l = 0;
        for (int e = 0; e < nE; ++e) {
            if (0 == (e & 31)) ia_l = (e + lane < nE) ? a.ent_a[e0 + e + lane] : 0u;
            int const r = e % kRingA;
            uint32_t const ia = __shfl_sync(0xffffffffu, ia_l, e & 31);
            if (e >= kRingA) mbar_wait(&bar_free[r], unsigned(((e / kRingA) - 1) & 1));
            if (leader) {
                if (TFQ_TC_ABLATE & 16) { mbar_arrive(&bar_a[r]); }
                else {
                    mbar_expect_tx(&bar_a[r], unsigned(ABLK*sizeof(float)));
                    unsigned char const *src = reinterpret_cast<unsigned char const*>(a.A + size_t(ia)*ABLK);
                    #pragma unroll
                    for (int kq = 0; kq < LM/4; ++kq)      // one k-quad slab each, leaving room for the lo slab behind it
                        bulk_g2s_hint(ring + size_t(r)*SLOT + size_t(kq)*2*SLAB, src + size_t(kq)*SLAB, SLAB, &bar_a[r], stream_once);
                }
            }
            __syncwarp();
        }
        // every arrival on the slot barriers must have happened before the barriers are invalidated at the end of the unit
        // (the commits complete in order: the last one is enough)
        if (nE > 0) mbar_wait(&bar_free[(nE - 1) % kRingA], unsigned(((nE - 1) / kRingA) & 1));
    } else if (kMmaWarp == w) {
        // ================= MMA warp =================================================================================
        // The whole warp runs the loop converged and ONE elected lane issues: with elect.sync the compiler emits
        // back-to-back UTCHMMA; a lane picked by "if (0 == lane)" costs an elect/branch loop (~125 cycles) per MMA.
        uint32_t const leader = elect_one_sync();
        uint32_t const ring_u32 = smem_u32(ring);
        for (int e = 0; e < nE; ++e) {
            int const s = e & 1, r = e % kRingA;
            TFQ_TRACE(e, 15);
            mbar_wait(&bar_ready[s], unsigned((e >> 1) & 1));   // X in TMEM, lo in shared memory (and the raw A landed)
            tc_fence_after();
            TFQ_TRACE(e, 16);
            if (leader) {
                uint32_t const sa = ring_u32 + uint32_t(r)*SLOT;
                uint32_t const xa = tmem_base + kTmemStage0 + uint32_t(s)*kStageCols;
                #pragma unroll
                for (int ks = 0; ks < ((TFQ_TC_ABLATE & 8) ? 0 : KS); ++ks) {
                    uint64_t const b = smem_desc_noswizzle(sa + ks*KSB, LBO, SBO);
                    uint32_t const first = (e > 0 || ks > 0) ? 1u : 0u;
                    // main sum in columns [0,N), correction sum in [N,2N): the tensor core truncates the fp32 accumulator
                    // once per MMA, so the small correction products must not share the large sum's accumulator
                    mma_tf32_ts(tmem_base,     xa + 8*ks,      b, IDESC_2N, first);   // Xhi * [Ahi ; Alo]
                    mma_tf32_ts(tmem_base + N, xa + LM + 8*ks, b, IDESC_N,  1u);      // Xlo * Ahi
                }
                mma_commit(&bar_mma[s]);       // the X stage is free again
                mma_commit(&bar_free[r]);      // and so is the A slot
            }
            __syncwarp();
            TFQ_TRACE(e, 17);
        }
    } else {
        // ================= converter warps: X -> TMEM, lo(A) -> shared memory, epilogue ==============================
        int const q = w & 3, h = w >> 2;
        int const m = 32*q + lane;
        int const g = m/(2*LN), cx = (m/LN) & 1, j = m % LN;
        uint32_t const iy = s_y[g];
        bool const has_g = (g < gs) && (kNoBlock != iy);
        uint32_t const xoff = uint32_t(cx)*LM*LN + uint32_t(KH*h)*LN + uint32_t(j);
        auto x_index = [&](int e) -> uint32_t { return (has_g && e < nE) ? a.ent_x[size_t(e0 + e)*gs + g] : kNoBlock; };
        auto load_x = [&](uint32_t ix, float (&xr)[KH], bool skip = false) {
            if (kNoBlock != ix && !(TFQ_TC_ABLATE & 4) && !skip) {
                float const *xp = a.x + size_t(ix)*XBLK + xoff;
                #pragma unroll
                for (int r = 0; r < KH; ++r) xr[r] = __ldg(xp + r*LN);
            } else {
                #pragma unroll
                for (int r = 0; r < KH; ++r) xr[r] = 0.f;
            }
        };
        // one entry: xc holds its X values; the loads of entry e+1 go to xn while entry e is split
        auto step = [&](int e, float (&xc)[KH], float (&xn)[KH], uint32_t ix_next, uint32_t &ix_next2) {
            int const s = e & 1, r = e % kRingA;
            TFQ_TRACE(e, tslot0 + 0);
            // the ~100 cycles a try_wait takes on an already completed barrier overlap with the load issue
            bool const stage_free = (e < 2) || mbar_try_wait(&bar_mma[s], unsigned(((e >> 1) - 1) & 1));
            bool const a_landed = mbar_try_wait(&bar_a[r], unsigned((e / kRingA) & 1));
            load_x(ix_next, xn, (TFQ_TC_ABLATE & 128) && ((e + 1) % 3 == 2));
            ix_next2 = x_index(e + 2);
            TFQ_TRACE(e, tslot0 + 1);
            if (e >= 2) { if (!stage_free) mbar_wait(&bar_mma[s], unsigned(((e >> 1) - 1) & 1)); tc_fence_after(); } // stage s is free again
            TFQ_TRACE(e, tslot0 + 2);
            // ---- X operand: split, registers -> tensor memory (lane = m, column = k) ------------------------
            {
                uint32_t const t0 = tmem_base + (uint32_t(32*q) << 16) + kTmemStage0 + uint32_t(s)*kStageCols + uint32_t(KH*h);
                constexpr int W = (KH < 16) ? KH : 16;             // columns per tcgen05.st
                #pragma unroll
                for (int c = 0; c < (((TFQ_TC_ABLATE & 2) || ((TFQ_TC_ABLATE & 128) && e % 3 == 2)) ? 0 : KH/W); ++c) {
                    uint32_t hi[W], lo[W];
                    #pragma unroll
                    for (int t = 0; t < W; ++t) split_rn(xc[W*c + t], hi[t], lo[t]);
                    if (8 == W) { tmem_st8(t0 + W*c, hi); tmem_st8(t0 + LM + W*c, lo); }
                    else        { tmem_st16(t0 + W*c, hi); tmem_st16(t0 + LM + W*c, lo); }
                }
            }
            // ---- A operand: hi = the raw block in the ring, lo = a - trunc(a) -> shared memory ------------------
            TFQ_TRACE(e, tslot0 + 3);
            if (!a_landed) mbar_wait(&bar_a[r], unsigned((e / kRingA) & 1));
            TFQ_TRACE(e, tslot0 + 4);
            {
                unsigned char *const slot = ring + size_t(r)*SLOT;
                #pragma unroll
                for (int c2 = 0; c2 < ((TFQ_TC_ABLATE & 1) ? 0 : (ABLK/4 + kConvThreads - 1)/kConvThreads); ++c2) {
                    int const c = tid + kConvThreads*c2;          // (k-quad, n) chunk of 4 k values
                    if (ABLK/4 % kConvThreads != 0 && c >= ABLK/4) break;
                    int const kq = c / N, n = c % N;
                    float4 v = *reinterpret_cast<float4 const*>(slot + size_t(kq)*2*SLAB + size_t(n)*16);
                    v.x = lo_trunc(v.x); v.y = lo_trunc(v.y); v.z = lo_trunc(v.z); v.w = lo_trunc(v.w);
                    *reinterpret_cast<float4*>(slot + size_t(kq)*2*SLAB + SLAB + size_t(n)*16) = v;
                }
            }
            TFQ_TRACE(e, tslot0 + 5);
            tmem_wait_st();
            fence_proxy_async();       // generic-proxy shared-memory writes -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (0 == lane) mbar_arrive(&bar_ready[s]);
            TFQ_TRACE(e, tslot0 + 6);
        };

        if (nE > 0) {
            float xa_[KH], xb_[KH];
            uint32_t i1 = x_index(1), i2 = kNoBlock;
            load_x(x_index(0), xa_);
            for (int e = 0; e < nE; e += 2) {
                step(e, xa_, xb_, i1, i2);                       // i2 := index of entry e+2
                if (e + 1 < nE) step(e + 1, xb_, xa_, i2, i1);   // i1 := index of entry e+3
                else break;
            }
            // all MMAs complete when the last commit has arrived (they retire in order)
            mbar_wait(&bar_mma[(nE - 1) & 1], unsigned(((nE - 1) >> 1) & 1));
            tc_fence_after();
        }

        // ---- epilogue: D -> registers, combine the four real products, store Y ---------------------------------
        float *const exch = reinterpret_cast<float*>(ring);   // [G][2][EC][LN] floats, aliases the A ring (all copies and MMAs are done)
        constexpr int EC = (LM < 32) ? LM : 32;              // accumulator columns per pass
        #pragma unroll 1
        for (int c = 0; c < LM/EC; ++c) {
            uint32_t d[EC];
            if (nE > 0) {
                uint32_t d2[EC];
                uint32_t const t0 = tmem_base + (uint32_t(32*q) << 16) + uint32_t(h)*LM + uint32_t(EC*c);   // D[m][(ca = h, i)]
                if (16 == EC) { tmem_ld16(t0, d); tmem_ld16(t0 + N, d2); } else { tmem_ld32(t0, d); tmem_ld32(t0 + N, d2); }
                tmem_wait_ld();
                #pragma unroll
                for (int i = 0; i < EC; ++i) d[i] = __float_as_uint(__uint_as_float(d[i]) + __uint_as_float(d2[i]));   // main + correction
            } else {
                #pragma unroll
                for (int i = 0; i < EC; ++i) d[i] = 0u;
            }
            if (1 == cx) {
                #pragma unroll
                for (int i = 0; i < EC; ++i) exch[((g*2 + h)*EC + i)*LN + j] = __uint_as_float(d[i]);
            }
            asm volatile("bar.sync 1, %0;" :: "n"(kConvThreads) : "memory");     // converter warps only
            if (0 == cx && has_g && !(TFQ_TC_ABLATE & 32)) {
                float *const yp = a.y + size_t(iy)*XBLK + size_t(h)*LM*LN + size_t(EC*c)*LN + j;   // plane h: 0 = Re, 1 = Im
                float const sgn = h ? 1.f : -1.f;                                 // Yr = XrAr - XiAi ; Yi = XrAi + XiAr
                #pragma unroll
                for (int i = 0; i < EC; ++i) d[i] = __float_as_uint(__uint_as_float(d[i]) + sgn*exch[((g*2 + (1 - h))*EC + i)*LN + j]);
                if (first_pass) {
                    #pragma unroll
                    for (int i = 0; i < EC; ++i) yp[i*LN] = __uint_as_float(d[i]);
                } else {            // (the same thread wrote this element in the previous pass)
                    #pragma unroll
                    for (int i = 0; i < EC; ++i) yp[i*LN] += __uint_as_float(d[i]);
                }
            }
            if (c + 1 < LM/EC) asm volatile("bar.sync 1, %0;" :: "n"(kConvThreads) : "memory");   // exch is reused
        }
    }
    tc_fence_before();
    __syncthreads();          // everyone is done with this unit's accumulator, ring and barriers
    tc_fence_after();
    if (0 == tid) {
        mbar_inval(&bar_mma[0]); mbar_inval(&bar_mma[1]); mbar_inval(&bar_ready[0]); mbar_inval(&bar_ready[1]);
        #pragma unroll
        for (int r = 0; r < kRingA; ++r) { mbar_inval(&bar_a[r]); mbar_inval(&bar_free[r]); }
    }
    } // passes
    } // units
    tc_fence_before();
    __syncthreads();
    if (kMmaWarp == w) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}


template <int LM, int LN>
tfqmrgpuStatus_t launch_tc(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    constexpr int ring = TcShape<LM>::ring, ctas = TcShape<LM>::ctas;
    constexpr size_t smem = 1024 + ring*2*size_t(2*LM*LM)*sizeof(float) + 1024; // barriers + A ring (hi and lo slabs)
    // `ctas` CTAs per SM by their TMEM columns: pad the request so that one more CTA can never be resident
    constexpr size_t smem_min = (2 == ctas) ? 80*1024 : 120*1024;
    constexpr size_t smem_req = (smem < smem_min) ? smem_min : smem;
    auto kernel = spmm_tc_kernel<LM, LN>;
    static size_t configured[kMaxDevices] = {0}; // per instantiation and device
    TFQ_CUDA(ensure_dynamic_smem(kernel, smem_req, configured));
    TcArgs a;
    a.y = static_cast<float*>(y); a.x = static_cast<float const*>(x); a.A = ws<float const>(p, p.off_A);
    a.unit_e0 = p.d_unit_e0; a.unit_y = p.d_unit_y; a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x;
    a.ctl = ws<Control const>(p, p.off_ctl); a.expect = expect; a.gstride = int(p.gmax); a.nUnits = p.nUnits;
    // TFQMRGPU_TC_CHAIN = entries per accumulation pass (default 896/LM, 7 at LM = 64): shorter chains cut the truncation error of
    // strongly cancelling sums (the accumulator is truncated once per MMA) for one more epilogue per pass
    static int const chain_env = [] { char const *e = std::getenv("TFQMRGPU_TC_CHAIN"); return e ? std::atoi(e) : 0; }();
    a.chain = (chain_env > 0) ? chain_env : ((64 == LM) ? 7 : 896/LM);
    static int sms_of[kMaxDevices] = {0};
    int dev = 0; cudaGetDevice(&dev);
    int &num_sms = sms_of[(dev >= 0 && dev < kMaxDevices) ? dev : 0];
    if (0 == num_sms) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t const grid = std::min<uint32_t>(p.nUnits, uint32_t(ctas)*uint32_t(num_sms));
    if (grid > 0) kernel<<<grid, kTcThreads, smem_req, stream>>>(a);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace

// level = TFQMRGPU_TENSOR (default 1): 0 never.  With the accumulation chains cut into passes (kChain) the attainable
// tfQMR residual is not worse than with the fp32 SIMT product for LM = 16, 32 and 64 (measured).
bool spmm_tc_supported(int LM, int LN, char precision, int level) {
    if (level < 1 || 'c' != precision) return false;
    return (16 == LM || 32 == LM || 64 == LM) && (16 == LN || 32 == LN || 64 == LN) && (LM <= LN);
}
int spmm_tc_columns_per_unit(int LN) { return 64/LN; }

tfqmrgpuStatus_t launch_spmm_tc(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    switch (p.LM*1000 + p.LN) {
#define TFQ_CASE(LM, LN) case LM*1000 + LN: return launch_tc<LM, LN>(p, y, x, expect, stream);
        TFQ_CASE(16, 16) TFQ_CASE(16, 32) TFQ_CASE(16, 64)
        TFQ_CASE(32, 32) TFQ_CASE(32, 64)
        TFQ_CASE(64, 64)
#undef TFQ_CASE
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
}

} // namespace tfq

#ifdef TFQ_TC_TRACE
extern "C" int tfq_tc_trace_dump(long long *host, int n) {
    return int(cudaMemcpyFromSymbol(host, tfq::g_tc_trace, size_t(n)*sizeof(long long)));
}
#endif
