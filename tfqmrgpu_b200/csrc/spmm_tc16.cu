// Block-sparse product  Y = A * X  for complex fp32 on the 5th-generation tensor cores (tcgen05, sm_100a),
// operands split into pairs of HALF-precision numbers ("3xFP16"), accumulators holding Y directly.
//
// Same role as spmm.cu (the reference's blocksparse_action_t::multiply + gemmNxNf,
// tfqmrgpu_blocksparse.hxx:71-199, tfqmrgpu_blockmult.hxx:10-93).  The complex block product is ONE real GEMM per
// (block row, A block) whose K dimension carries Re and Im interleaved, k' = (k, Re|Im):
//
//     D[m][i] += sum_k'  Xop[m][k'] * Aop[i][k']     m = (g, j, t = Re|Im of Y)   -> 128 rows  (G*LN*2)
//     Aop[i][(k, c)]      = A[c][k][i]
//     Xop[(g, j, Re)][(k, c)] = ( Xr[k][j], -Xi[k][j] )        Xop[(g, j, Im)][(k, c)] = ( Xi[k][j],  Xr[k][j] )
//
// so D[(g, j, t)][i] IS Y_t[i][j]: the products Ar*Xr and Ai*Xi cancel inside every MMA like they do in the reference's
// fused multiply-adds, the partial sums stay of the size of Y, and the tensor core's once-per-MMA truncation of the fp32
// accumulator acts on numbers of that size.  (The earlier formulation with the four real products XrAr, XrAi, XiAr, XiAi
// in separate accumulators needed half the MMA instructions, but on strongly cancelling sums - the reference harness's
// cos/sin fill: 700 of |terms| add up to |Y| = 3 - its error could not be brought below 7e-5..2e-4 with accumulation
// chains of any length: the large sums themselves are only fp32 numbers.  Measured, DESIGN.md.)
//
//   * Both operands arrive PRE-SPLIT.  A power-of-two scale per block row of A (per right-hand-side column of X) maps
//     the operand into the fp16 range, then  v*s = hi + lo/2048  with hi = fp16(v*s), lo = fp16((v*s - hi)*2048): two
//     11-bit significands, |v*s - hi - lo/2048| <= 2^-24 |v*s|, i.e. fp32-grade, in the SAME 4 bytes per element as the
//     fp32 value.  setMatrix('A') stores the A blocks like that (xop.cu: [hi ; lo] rows, K-major core matrices, one
//     bulk copy = one MMA-ready operand), and every X-shaped vector that is multiplied is converted ONCE per
//     product (by the vector kernel that writes it, vecops.cu) instead of once per use (27 times on the 27-point stencil).
//   * kind::f16 MMAs, K = 16 per instruction:  D  += Xhi * Ahi,  D' += Xhi * Alo + Xlo * Ahi   (D' carries the factor
//     2048; one MMA of N' = 2*LM writes [D | D'], one of N' = LM adds Xlo * Ahi).
//   * One persistent CTA per SM: a copy warp (A ring, bulk copies), TWO MMA warps that take alternate entries into
//     accumulator sets of their own (a single issuer alternates between waiting for operands and issuing at the pipe's rate;
//     two hide each other's waits, and the two interleaved chains halve the length of every accumulation chain),
//     8 converter warps (X operand: L2 -> registers -> sign / swap -> tcgen05.st -> TMEM) and 4 epilogue warps.
//     All barriers live for the whole launch (running entry / segment counters give slot and parity), so nothing drains
//     at unit boundaries: the copy warp and the converters stream over the CTA's flat entry range.
//   * Four accumulator sets in TMEM (two per MMA warp): the MMAs of the next segment / unit start while the epilogue
//     still reads the previous one; the epilogue adds the chains and segments in fp32 registers.
//
// Block sizes: LM in {16, 32}, LN in {16, 32, 64} natively; 64 x 64 blocks run as 2 x 2 sub-blocks of 32 x 32 through
// index tables (plan.cu) - the "virtual" V64 mode: the two j-halves of an X block are the unit's two block columns.
#include "tfq_internal.hpp"
#include <cstdlib>
#include <algorithm>

namespace tfq {

namespace {

constexpr int kConvWarps = 8, kEpiWarps = 4, kMmaWarps = 2;
constexpr int kMmaWarp0 = kConvWarps + kEpiWarps, kCopyWarp = kMmaWarp0 + kMmaWarps, kXCopyWarp = kCopyWarp + 1;
constexpr int kThreads = 32*(kConvWarps + kEpiWarps + kMmaWarps + 2);
constexpr int kSets = 2*kMmaWarps;                     // accumulator sets in TMEM: two per MMA warp

struct Tc16Args {
    float *y;
    uint4 const *xop;            // X operand blocks (xop.cu)
    unsigned char const *Aop;    // A operand blocks
    float const *a_inv;          // [mb] 1/scale of the block rows of A
    float const *x_inv;          // [nCols*LN] 1/scale of the right-hand-side columns
    uint32_t const *cta_u0;      // [grid+1] unit range of every CTA
    uint32_t const *unit_e0, *unit_y, *unit_row, *ent_a, *ent_x;
    uint32_t const *blockcol;    // block column of a storage-ordered X block
    Control const *ctl; int expect; int gstride;
    int seg;                     // entries per accumulation segment
};

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, void const *src_gmem, unsigned bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, void const *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
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
// Bounded by wall time (%globaltimer, 4 s): a pipeline that never signals is a bug and must end the launch with an error
// instead of hanging the stream; a legitimately slow stage (profiler replay, managed memory migrating) is waited for.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins) {
        if (0 == (spins & 0xfffu)) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (0 == t0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
    return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], fp16 inputs, fp32 accumulation; issued by ONE thread
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, uint32_t const *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, uint32_t const *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows (n) of 16 bytes (8 halves along k);
// LBO = byte stride between core matrices along K, SBO = byte stride between 8-row groups along N
__device__ __forceinline__ uint64_t smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3fff);
    d |= uint64_t((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= uint64_t(1) << 46;                          // descriptor version of sm_100
    return d;
}

// dev-only timing ablations (results are WRONG with any bit set): 1 no X loads, 2 no tcgen05.st, 4 no MMAs (commits only),
// 8 no A bulk copies, 16 no Y stores
#ifndef TFQ_TC16_ABLATE
#define TFQ_TC16_ABLATE 0
#endif

// dev-only per-role cycle accounting of CTA 0 (scripts/dev_tc16_trace.py): where every role's time goes, summed over the launch
#ifdef TFQ_TC16_TRACE
__device__ long long g_tc16_trace[8*8];
#define TR_DECL long long tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64()
#define TR_LAP(k) do { long long const n_ = clock64(); tr_[k] += n_ - tr_t; tr_t = n_; } while (0)
#define TR_DUMP(role) do { if (0 == blockIdx.x && 0 == lane) { for (int k_ = 0; k_ < 8; ++k_) g_tc16_trace[(role)*8 + k_] = tr_[k_]; } } while (0)
#else
#define TR_DECL do { } while (0)
#define TR_LAP(k) do { } while (0)
#define TR_DUMP(role) do { } while (0)
#endif

template <int LM, int LN> struct Tc16Shape {
    static constexpr int KS  = LM/8;                    // MMA k-steps per entry: K' = 2*LM halves, 16 per instruction
    static constexpr int N   = LM;                      // accumulator columns of D (i); D' behind it
    static constexpr int NB  = 2*LM;                    // rows of [Ahi ; Alo] = accumulator columns [D | D'] of one set
    static constexpr uint32_t ABYTES = 8u*LM*LM;        // one A operand block: [LM/4 k'-octets][NB rows][8 halves]
    static constexpr uint32_t SLAB   = 16u*NB;          // bytes of one k'-octet slab
    static constexpr int NCHH = LM/4;                   // 16-byte chunks (4 complex k) of the hi half of an X operand row; as many lo
    static constexpr uint32_t XROWS = LN;               // rows (j) of one X operand block
    static constexpr uint32_t XCH   = 2u*NCHH*XROWS;    // uint4 elements of one X operand block
    static constexpr int SC  = 2*LM;                    // TMEM columns of one X stage (hi: LM, lo: LM)
    static constexpr int NS  = (32 == LM) ? 4 : 8;      // X stages in TMEM
    static constexpr int RA  = (32 == LM) ? 16 : 32;    // A blocks in flight (bulk-copy ring)
    static constexpr uint32_t ACC0 = 0, STAGE0 = kSets*NB;  // TMEM columns: the accumulator sets, then the X stages
    static constexpr uint32_t TMEM_COLS = 512;
    static constexpr int G   = 64/LN;                   // block columns per unit: 128 MMA rows = G * LN * 2
    static constexpr uint32_t XBYTES = 8u*LM*LN;        // one X operand block
    static constexpr uint32_t XSLOT  = G*XBYTES;        // the X operand blocks of one entry
    static constexpr int XR  = (32 == LM) ? 4 : 8;      // entries whose X blocks are in flight / staged in shared memory
    static constexpr size_t smem = 2048 + size_t(RA)*ABYTES + size_t(XR)*XSLOT;
    static_assert(STAGE0 + NS*SC <= TMEM_COLS, "tensor memory");
    static_assert((RA & (RA - 1)) == 0 && (NS & (NS - 1)) == 0 && (XR & (XR - 1)) == 0, "ring sizes are powers of two");
};

// own entries of MMA warp `mw` among the unit's entries (the CTA's entries [nb, nb + nE)): every other entry, counted from the
// unit's first one - the two accumulation chains of a Y block then do not depend on which CTA (or GPU) works on the unit
__device__ __forceinline__ void own_range(uint32_t nb, uint32_t nE, uint32_t mw, uint32_t &first, uint32_t &count) {
    first = nb + mw;
    count = (nE + 1u - mw)/2u;       // mw = 0: ceil(nE/2), mw = 1: floor(nE/2)
}

template <int LM, int LN, bool V64>
__global__ void __launch_bounds__(kThreads, 1)
spmm_tc16_kernel(Tc16Args const a)
{
    static_assert(LM == 16 || LM == 32, "k-steps of 16");
    static_assert(LN == 16 || LN == 32 || LN == 64, "128 MMA rows = G * 2 * LN");
    static_assert(!V64 || (32 == LM && 32 == LN), "64 x 64 blocks run as 32 x 32 sub-blocks");
    using S = Tc16Shape<LM, LN>;
    constexpr int KS = S::KS, N = S::N, NB = S::NB, NCHH = S::NCHH, NS = S::NS, RA = S::RA, SC = S::SC, XR = S::XR, G = S::G;
    // instruction descriptor: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N' >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t IDESC_BASE = (1u << 4) | (uint32_t(128 >> 4) << 24);
    constexpr uint32_t IDESC_NB = IDESC_BASE | (uint32_t(NB >> 3) << 17);
    constexpr uint32_t IDESC_N  = IDESC_BASE | (uint32_t(N >> 3) << 17);

    if (a.expect >= 0 && a.ctl->state != a.expect) return; // device-resident solver control

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *const bar_a_full   = reinterpret_cast<uint64_t*>(smem_raw);   // [RA] A block has landed
    uint64_t *const bar_done     = bar_a_full + RA;                          // [RA] the MMAs of the entry have completed: its A slot
                                                                             //      (entry n + RA) and its X stage (entry n + NS) are free
    uint64_t *const bar_x_full   = bar_done + RA;                            // [NS] X operand of the stage is in TMEM
    uint64_t *const bar_acc_full = bar_x_full + NS;                          // [kSets] a segment's MMAs have completed
    uint64_t *const bar_acc_free = bar_acc_full + kSets;                     // [kSets] the epilogue has read the set
    uint64_t *const bar_xs_full  = bar_acc_free + kSets;                     // [XR] the entry's X operand blocks have landed in shared memory
    uint64_t *const bar_xs_free  = bar_xs_full + XR;                         // [XR] the converter warps have read them
    uint32_t *const s_xvalid = reinterpret_cast<uint32_t*>(smem_raw + 1024); // [XR][4] 1: block column g of the entry has an X block
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 1536);
    unsigned char *const ring = smem_raw + 2048;
    unsigned char *const xring = ring + size_t(RA)*S::ABYTES;
    static_assert((2*RA + NS + 2*kSets + 2*XR)*8 <= 1024 && XR*16 <= 512, "barrier area");
    static_assert(RA >= NS, "one ring of completion barriers serves the A slots and the X stages");

    int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    int const gs = a.gstride;

    if (kMmaWarp0 == w) tmem_alloc(tmem_slot, S::TMEM_COLS);
    if (0 == tid) {
        for (int r = 0; r < RA; ++r) { mbar_init(&bar_a_full[r], 1); mbar_init(&bar_done[r], 1); }
        for (int s = 0; s < NS; ++s) mbar_init(&bar_x_full[s], kConvWarps);
        for (int c = 0; c < kSets; ++c) { mbar_init(&bar_acc_full[c], 1); mbar_init(&bar_acc_free[c], kEpiWarps); }
        for (int x = 0; x < XR; ++x) { mbar_init(&bar_xs_full[x], 1); mbar_init(&bar_xs_free[x], kConvWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t const tmem_base = *tmem_slot;

    // this CTA's units [u0, u1) are contiguous in the unit and entry tables (plan.cu deals the units of consecutive
    // block rows round the CTAs, so that at any time the CTAs work on neighbouring rows: the X blocks stay in L2)
    uint32_t const u0 = a.cta_u0[blockIdx.x], u1 = a.cta_u0[blockIdx.x + 1];
    uint32_t const E0 = a.unit_e0[u0], E1 = a.unit_e0[u1];
    uint32_t const total = E1 - E0;
    uint32_t const seg = uint32_t(a.seg);

    if (kCopyWarp == w) {
        // ================= copy warp: the A blocks of the CTA's entries, in order, into the ring ======================
        uint32_t const leader = elect_one_sync();
        uint64_t const stream_once = policy_evict_first();   // A is read once per product: keep it from displacing X in L2
        uint32_t n = 0;
        TR_DECL;
        for (uint32_t eb = 0; eb < total; eb += 32) {
            uint32_t const ia_l = (eb + lane < total) ? a.ent_a[E0 + eb + lane] : 0u;
            uint32_t const cnt = (total - eb < 32u) ? (total - eb) : 32u;
            for (uint32_t t = 0; t < cnt; ++t, ++n) {
                uint32_t const ia = __shfl_sync(0xffffffffu, ia_l, int(t));
                uint32_t const r = n & (RA - 1), use = n / RA;
                TR_LAP(1);
                if (use > 0) mbar_wait(&bar_done[r], (use - 1) & 1);
                TR_LAP(0);
                if (leader) {
                    if (TFQ_TC16_ABLATE & 8) mbar_arrive(&bar_a_full[r]);
                    else {
                        mbar_expect_tx(&bar_a_full[r], S::ABYTES);
                        bulk_g2s_hint(ring + size_t(r)*S::ABYTES, a.Aop + size_t(ia)*S::ABYTES, S::ABYTES, &bar_a_full[r], stream_once);
                    }
                }
                __syncwarp();
            }
        }
        TR_LAP(1);
        TR_DUMP(0);
    } else if (kXCopyWarp == w) {
        // ================= X copy warp: the X operand blocks of the CTA's entries, in order, into their shared-memory ring ====
        // (they come from L2 - 27 uses per block on the stencil -; the bulk-copy engine keeps several entries in flight without
        // holding registers or load slots of the converter warps)
        uint32_t const leader = elect_one_sync();
        uint32_t n = 0;
        for (uint32_t eb = 0; eb < total; eb += 32) {
            uint32_t ix_l[G];
            #pragma unroll
            for (int g = 0; g < G; ++g) ix_l[g] = (eb + lane < total && g < gs) ? a.ent_x[size_t(E0 + eb + lane)*gs + g] : kNoBlock;
            uint32_t const cnt = (total - eb < 32u) ? (total - eb) : 32u;
            for (uint32_t t = 0; t < cnt; ++t, ++n) {
                uint32_t ix[G];
                #pragma unroll
                for (int g = 0; g < G; ++g) ix[g] = __shfl_sync(0xffffffffu, ix_l[g], int(t));
                uint32_t const x = n & (XR - 1), use = n / XR;
                if (use > 0) mbar_wait(&bar_xs_free[x], (use - 1) & 1);
                if (leader) {
                    uint32_t nvalid = 0;
                    #pragma unroll
                    for (int g = 0; g < G; ++g) {
                        bool const v = (kNoBlock != ix[g]) && !(TFQ_TC16_ABLATE & 1);
                        s_xvalid[x*4 + g] = v ? 1u : 0u;
                        nvalid += v ? 1u : 0u;
                    }
                    if (0 == nvalid) mbar_arrive(&bar_xs_full[x]);
                    else {
                        mbar_expect_tx(&bar_xs_full[x], nvalid*S::XBYTES);
                        #pragma unroll
                        for (int g = 0; g < G; ++g)
                            if (kNoBlock != ix[g] && !(TFQ_TC16_ABLATE & 1))
                                bulk_g2s(xring + size_t(x)*S::XSLOT + size_t(g)*S::XBYTES, a.xop + size_t(ix[g])*S::XCH, S::XBYTES, &bar_xs_full[x]);
                    }
                }
                __syncwarp();
            }
        }
    } else if (w >= kMmaWarp0) {
        // ================= MMA warps: one elected lane of each issues; warp mw takes every other entry of a unit ===========
        uint32_t const mw = uint32_t(w - kMmaWarp0);
        uint32_t const leader = elect_one_sync();
        uint32_t const ring_u32 = smem_u32(ring);
        uint32_t nb = 0, sgw = 0;                                   // entries before this unit; this warp's segment counter
        uint32_t e_begin = E0;
        uint32_t e_end_next = (u0 < u1) ? a.unit_e0[u0 + 1] : E0;
        TR_DECL;
        for (uint32_t u = u0; u < u1; ++u) {
            uint32_t const e_end = e_end_next;
            if (u + 1 < u1) e_end_next = a.unit_e0[u + 2];          // one unit ahead: not on the critical path
            uint32_t const nE = e_end - e_begin;
            e_begin = e_end;
            uint32_t first, cnt;
            own_range(nb, nE, mw, first, cnt);
            nb += nE;
            uint32_t const nSeg = (cnt + seg - 1)/seg;
            uint32_t n = first;
            for (uint32_t s = 0; s < nSeg; ++s, ++sgw) {
                uint32_t const len = (cnt*(s + 1))/nSeg - (cnt*s)/nSeg;
                uint32_t const c = 2*(sgw & 1) + mw, cuse = sgw >> 1;
                TR_LAP(3);
                if (cuse > 0) { mbar_wait(&bar_acc_free[c], (cuse - 1) & 1); tc_fence_after(); }
                TR_LAP(2);
                uint32_t const acc = tmem_base + S::ACC0 + c*NB;
                for (uint32_t t = 0; t < len; ++t, n += 2) {
                    uint32_t const r = n & (RA - 1), st = n & (NS - 1);
                    bool const a_ok = mbar_try_wait(&bar_a_full[r], (n / RA) & 1);      // both tests in flight together
                    bool const x_ok = mbar_try_wait(&bar_x_full[st], (n / NS) & 1);
                    if (!a_ok) mbar_wait(&bar_a_full[r], (n / RA) & 1);
                    if (!x_ok) mbar_wait(&bar_x_full[st], (n / NS) & 1);
                    tc_fence_after();
                    TR_LAP(0);
                    if (leader) {
                        uint32_t const sa = ring_u32 + r*S::ABYTES;
                        uint32_t const xa = tmem_base + S::STAGE0 + st*SC;
                        #pragma unroll
                        for (int ks = 0; ks < ((TFQ_TC16_ABLATE & 4) ? 0 : KS); ++ks) {
                            uint64_t const b = smem_desc_noswizzle(sa + ks*2*S::SLAB, S::SLAB, 128);
                            mma_f16_ts(acc,     xa + 8*ks,      b, IDESC_NB, (t > 0 || ks > 0) ? 1u : 0u);   // Xhi * [Ahi ; Alo] -> [D | D']
                            mma_f16_ts(acc + N, xa + LM + 8*ks, b, IDESC_N,  1u);                              // Xlo * Ahi        ->      D'
                        }
                        mma_commit(&bar_done[r]);
                        if (t == len - 1) mma_commit(&bar_acc_full[c]);
                    }
                    __syncwarp();
                    TR_LAP(1);
                }
            }
        }
        TR_DUMP(1 + mw);
    } else if (w < kConvWarps) {
        // ================= converter warps: X operand rows, shared memory -> registers -> tensor memory ===================
        // Warp (q4, h): TMEM lanes 32 q4 .. +31 = operand rows m = (g, j, t); h = 0 writes the hi half of the row, h = 1 the lo
        // half.  An operand row holds (Re, Im) pairs of halves along k; the row of Y's real part needs (Re, -Im), the row of
        // its imaginary part (Im, Re): one byte permutation and one XOR per word.  The two rows of a column j are neighbouring
        // lanes, so they read the same 16 bytes of shared memory in the same instruction (a broadcast).
        int const q4 = w & 3, h = w >> 2;
        int const m = 32*q4 + lane;
        int const g = m/(2*LN), t = m & 1;               // g is warp-uniform for LN = 16, 32, 64
        uint32_t const row = uint32_t((m % (2*LN)) >> 1);
        uint32_t const lane_base = tmem_base + (uint32_t(32*q4) << 16) + S::STAGE0 + uint32_t(h)*LM;
        uint32_t const sel = t ? 0x1032u : 0x3210u, flip = t ? 0u : 0x80000000u;
        uint4 const *const xrow = reinterpret_cast<uint4 const*>(xring + size_t(g)*S::XBYTES) + size_t(h)*NCHH*S::XROWS + row;
        TR_DECL;
        for (uint32_t n = 0; n < total; ++n) {
            uint32_t const x = n & (XR - 1), st = n & (NS - 1);
            mbar_wait(&bar_xs_full[x], (n / XR) & 1);
            TR_LAP(0);
            uint32_t r[4*NCHH];
            bool const valid = (g < gs) && (0u != *reinterpret_cast<uint32_t const volatile*>(&s_xvalid[x*4 + g]));
            if (valid) {
                uint4 const *src = xrow + size_t(x)*(S::XSLOT/16);
                #pragma unroll
                for (int q = 0; q < NCHH; ++q) {
                    uint4 const b = src[q*S::XROWS];
                    r[4*q + 0] = __byte_perm(b.x, 0u, sel) ^ flip; r[4*q + 1] = __byte_perm(b.y, 0u, sel) ^ flip;
                    r[4*q + 2] = __byte_perm(b.z, 0u, sel) ^ flip; r[4*q + 3] = __byte_perm(b.w, 0u, sel) ^ flip;
                }
            } else {
                #pragma unroll
                for (int q = 0; q < 4*NCHH; ++q) r[q] = 0u;
            }
            __syncwarp();
            if (0 == lane) mbar_arrive(&bar_xs_free[x]);          // the words are in registers: the slot may be refilled
            TR_LAP(1);
            if (n >= uint32_t(NS)) { mbar_wait(&bar_done[(n - NS) & (RA - 1)], ((n - NS) / RA) & 1); tc_fence_after(); }   // the stage's previous entry
            TR_LAP(2);
            if (!(TFQ_TC16_ABLATE & 2)) {
                if (32 == LM) tmem_st32(lane_base + st*SC, r); else tmem_st16(lane_base + st*SC, r);
                tmem_wait_st();
            }
            tc_fence_before();
            __syncwarp();
            if (0 == lane) mbar_arrive(&bar_x_full[st]);
            TR_LAP(3);
        }
        if (0 == q4) TR_DUMP(3 + h);
    } else {
        // ================= epilogue warps: accumulator sets -> registers (fp32 sums of chains and segments) -> Y ===========
        int const q4 = (w - kConvWarps) & 3;
        int const m = 32*q4 + lane;
        int const g = m/(2*LN), t = m & 1, j = (m % (2*LN)) >> 1;
        uint32_t const lane_base = tmem_base + (uint32_t(32*q4) << 16) + S::ACC0;
        uint32_t nb = 0, sg[2] = {0u, 0u};
        uint32_t e_begin = E0;
        uint32_t e_end_next = (u0 < u1) ? a.unit_e0[u0 + 1] : E0;
        TR_DECL;
        for (uint32_t u = u0; u < u1; ++u) {
            uint32_t const e_end = e_end_next;
            if (u + 1 < u1) e_end_next = a.unit_e0[u + 2];
            uint32_t const nE = e_end - e_begin;
            e_begin = e_end;
            uint32_t first, cnt0, cnt1;
            own_range(nb, nE, 0u, first, cnt0);
            own_range(nb, nE, 1u, first, cnt1);
            nb += nE;
            uint32_t const nSeg0 = (cnt0 + seg - 1)/seg, nSeg1 = (cnt1 + seg - 1)/seg;
            uint32_t const iy = (g < gs) ? a.unit_y[size_t(u)*gs + g] : kNoBlock;
            float acc[LM];
            #pragma unroll
            for (int i = 0; i < LM; ++i) acc[i] = 0.f;
            // the sets in the order in which the MMA warps complete them: segment s of warp 0, segment s of warp 1, ...
            for (uint32_t s = 0; s < ((nSeg0 > nSeg1) ? nSeg0 : nSeg1); ++s) {
                #pragma unroll
                for (uint32_t mw = 0; mw < 2; ++mw) {
                    if (s >= (mw ? nSeg1 : nSeg0)) continue;
                    uint32_t const c = 2*(sg[mw] & 1) + mw;
                    TR_LAP(2);
                    mbar_wait(&bar_acc_full[c], (sg[mw] >> 1) & 1);
                    tc_fence_after();
                    TR_LAP(0);
                    #pragma unroll
                    for (int ch = 0; ch < LM/16; ++ch) {
                        uint32_t d[16], d2[16];
                        tmem_ld16(lane_base + c*NB + 16*ch, d);
                        tmem_ld16(lane_base + c*NB + N + 16*ch, d2);
                        tmem_wait_ld();
                        #pragma unroll
                        for (int i = 0; i < 16; ++i)   // main + correction/2048, then the running sum: both rounded to nearest
                            acc[16*ch + i] += fmaf(__uint_as_float(d2[i]), 1.f/2048.f, __uint_as_float(d[i]));
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (0 == lane) mbar_arrive(&bar_acc_free[c]);
                    sg[mw] += 1;
                    TR_LAP(1);
                }
            }
            if (kNoBlock != iy && !(TFQ_TC16_ABLATE & 16)) {     // lane (g, j, t) holds Y_t[i][j] of block column g for all i
                float scale;
                float *yp;
                if (V64) {   // iy = 2*(Y block) + (row half); g = column half
                    uint32_t const yb = iy >> 1, ih = iy & 1;
                    scale = a.a_inv[a.unit_row[u]]*a.x_inv[size_t(a.blockcol[yb])*64 + 32*g + j];
                    yp = a.y + size_t(yb)*8192 + size_t(t)*4096 + size_t(32*ih)*64 + 32*g + j;
                } else {
                    scale = a.a_inv[a.unit_row[u]]*a.x_inv[size_t(a.blockcol[iy])*LN + j];
                    yp = a.y + size_t(iy)*(2*LM*LN) + size_t(t)*LM*LN + j;                 // plane t: 0 = Re, 1 = Im
                }
                constexpr int YS = V64 ? 64 : LN;
                #pragma unroll
                for (int i = 0; i < LM; ++i) yp[i*YS] = acc[i]*scale;
            }
            TR_LAP(2);
        }
        if (0 == q4) TR_DUMP(5);
    }
    tc_fence_before();
    __syncthreads();
    if (kMmaWarp0 == w) { tc_fence_after(); tmem_dealloc(tmem_base, S::TMEM_COLS); }
}

template <int LM, int LN, bool V64>
tfqmrgpuStatus_t launch_tc16(Plan const &p, void *y, int expect, cudaStream_t stream)
{
    using S = Tc16Shape<LM, LN>;
    auto kernel = spmm_tc16_kernel<LM, LN, V64>;
    static size_t configured[kMaxDevices] = {0}; // per instantiation and device
    TFQ_CUDA(ensure_dynamic_smem(kernel, S::smem, configured));
    Tc16Args a;
    a.y = static_cast<float*>(y);
    a.xop = ws<uint4 const>(p, p.off_xop); a.Aop = ws<unsigned char const>(p, p.off_A);
    a.a_inv = ws<float const>(p, p.off_ainv); a.x_inv = ws<float const>(p, p.off_xsinv);
    a.cta_u0 = p.d_cta_u0; a.unit_e0 = p.d_unit_e0; a.unit_y = p.d_unit_y; a.unit_row = p.d_unit_row;
    a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x; a.blockcol = p.d_blockcol;
    a.ctl = ws<Control const>(p, p.off_ctl); a.expect = expect; a.gstride = int(p.gmax);
    a.seg = p.tc_seg;
    if (p.tc_grid > 0 && p.nUnits > 0) kernel<<<p.tc_grid, kThreads, S::smem, stream>>>(a);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace

bool spmm_tc16_supported(int LM, int LN, char precision, int level) {
    if (1 != level || 'c' != precision) return false;
    return (16 == LM || 32 == LM || 64 == LM) && (16 == LN || 32 == LN || 64 == LN) && (LM <= LN);
}
// block columns per unit (the virtual 64 x 64 mode has one real block column = two virtual ones)
int spmm_tc16_columns_per_unit(int LM, int LN) { return (64 == LM) ? 1 : 64/LN; }
// entries per accumulation segment of ONE of the two interleaved chains (the accumulators hold numbers of the size of Y:
// long chains are harmless; the cap only bounds what a very long row can pile up before the fp32 sum in registers)
int spmm_tc16_default_segment(int LM) { return (16 == LM) ? 128 : 64; }

// the product proper; the X operand (xop, scales) must have been produced from x by launch_xop (xop.cu)
tfqmrgpuStatus_t launch_spmm_tc16(Plan const &p, void *y, int expect, cudaStream_t stream)
{
    switch (p.LM*1000 + p.LN) {
        case 16016: return launch_tc16<16, 16, false>(p, y, expect, stream);
        case 16032: return launch_tc16<16, 32, false>(p, y, expect, stream);
        case 16064: return launch_tc16<16, 64, false>(p, y, expect, stream);
        case 32032: return launch_tc16<32, 32, false>(p, y, expect, stream);
        case 32064: return launch_tc16<32, 64, false>(p, y, expect, stream);
        case 64064: return launch_tc16<32, 32, true >(p, y, expect, stream);
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
}

} // namespace tfq

#ifdef TFQ_TC16_TRACE
extern "C" int tfq_tc16_trace_dump(long long *host, int n) {
    return int(cudaMemcpyFromSymbol(host, tfq::g_tc16_trace, size_t(n)*sizeof(long long)));
}
#endif
