// Block-sparse product  Y = A * X  for complex fp32 on the 5th-generation tensor cores (tcgen05, sm_100a), fp16 operand
// pairs, PLANAR formulation: the four real products in separate accumulators.
//
// Same role and same operand number format as spmm_tc16.cu; what differs is the GEMM the complex block product is mapped to:
//
//     D[m][n] += sum_k  Xop[m][k] * Aop[n][k]        m = (g, Re|Im of X, j)   -> 128 rows  (G*2*LN)
//                                                    n = (Re|Im of A, i)      -> N = 2*LM columns
//
// D holds XrAr, XrAi, XiAr, XiAi and the epilogue forms Yr = XrAr - XiAi, Yi = XrAi + XiAr.  Every X element is one MMA row
// entry (spmm_tc16.cu needs it in two rows) and N is twice as large: half the MMA instructions and half the operand bytes
// through the converter warps - this is the FAST kernel (config 3: see DESIGN.md).  Its price: XrAr and XiAi are
// accumulated separately, so a sum that cancels between them (the reference harness's cos/sin fill) is formed from two large
// fp32 numbers, and the tensor core truncates the accumulator once per MMA.  The row's entries are therefore cut into short
// SEGMENTS (4 entries at LM = 32) that alternate between two accumulator sets and are added in fp32 by the epilogue warps;
// measured error on the harness's fill: <= 9e-5 up to 27 entries of 32 x 32 (bar 1e-4, bench_tfqmrgpu.cu:414), growing with
// entries * LM.  plan.cu selects this kernel for plans whose longest row has entries * LM <= 864 and spmm_tc16.cu (whose
// accumulators hold Y itself) for longer rows; TFQMRGPU_TC_FORM=direct|planar overrides.
//
// Roles of the persistent CTA (one per SM): copy warp (A ring, bulk copies), TWO MMA warps that take alternate segments (each
// owns one accumulator set: one waits for its operands while the other issues), 8 converter warps in two groups on alternate
// entries (X operand: L2 -> registers -> tcgen05.st -> TMEM) and 8 epilogue warps.
#include "tfq_internal.hpp"
#include <cstdlib>
#include <algorithm>

namespace tfq {

namespace {

constexpr int kConvWarps = 8, kEpiWarps = 8, kMmaWarps = 2;
constexpr int kMmaWarp0 = kConvWarps + kEpiWarps, kCopyWarp = kMmaWarp0 + kMmaWarps;
constexpr int kThreads = 32*(kConvWarps + kEpiWarps + kMmaWarps + 1);
constexpr int kConvGroups = 2;                         // converter groups working on alternate entries
constexpr int kWarpsPerGroup = kConvWarps/kConvGroups; // = 4: one warp per TMEM lane quarter

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
#ifndef TFQ_TC16P_ABLATE
#define TFQ_TC16P_ABLATE 0
#endif

// dev-only per-role cycle accounting of CTA 0 (scripts/dev_tc16_trace.py): where every role's time goes, summed over the launch
#ifdef TFQ_TC16P_TRACE
__device__ long long g_tc16p_trace[8*8];
#define TR_DECL long long tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64()
#define TR_LAP(k) do { long long const n_ = clock64(); tr_[k] += n_ - tr_t; tr_t = n_; } while (0)
#define TR_DUMP(role) do { if (0 == blockIdx.x && 0 == lane) { for (int k_ = 0; k_ < 8; ++k_) g_tc16p_trace[(role)*8 + k_] = tr_[k_]; } } while (0)
#else
#define TR_DECL do { } while (0)
#define TR_LAP(k) do { } while (0)
#define TR_DUMP(role) do { } while (0)
#endif

template <int LM, int LN> struct Tc16Shape {
    static constexpr int G   = 64/LN;                   // block columns per unit: 128 MMA rows = G * 2 * LN
    static constexpr int KS  = LM/16;                   // MMA k-steps per entry
    static constexpr int N   = 2*LM;                    // (Re|Im of A, i)
    static constexpr int NB  = 4*LM;                    // rows of [Ahi ; Alo] = accumulator columns [D | D']
    static constexpr uint32_t ABYTES = 8u*LM*LM;        // one A operand block: [LM/8 k-octets][NB rows][8 halves]
    static constexpr uint32_t SLAB   = 16u*NB;          // bytes of one k-octet slab
    static constexpr int NCH = LM/4;                    // 16-byte chunks of one X operand row (hi chunks, then lo chunks)
    static constexpr uint32_t XROWS = 2u*LN;            // rows (Re|Im, j) of one X operand block
    static constexpr uint32_t XCH   = NCH*XROWS;        // uint4 elements of one X operand block
    static constexpr int SC  = LM;                      // TMEM columns of one X stage (hi: LM/2, lo: LM/2)
    static constexpr int NS  = (32 == LM) ? 8 : 16;     // X stages in TMEM
    static constexpr int RA  = (32 == LM) ? 16 : 32;    // A blocks in flight (bulk-copy ring)
    static constexpr uint32_t ACC0 = 0, STAGE0 = 2*NB;  // TMEM columns: two accumulator sets, then the X stages
    static constexpr uint32_t TMEM_COLS = 512;
    static constexpr uint32_t EXCH = uint32_t(G)*2*LM*LN*4;   // bytes of one exchange buffer of the epilogue
    static constexpr size_t smem = 1024 + size_t(RA)*ABYTES + 2*size_t(EXCH);
    static_assert(STAGE0 + NS*SC <= TMEM_COLS, "tensor memory");
    static_assert((RA & (RA - 1)) == 0 && (NS & (NS - 1)) == 0, "ring sizes are powers of two");
};

template <int LM, int LN, bool V64>
__global__ void __launch_bounds__(kThreads, 1)
spmm_tc16p_kernel(Tc16Args const a)
{
    static_assert(LM == 16 || LM == 32, "k-steps of 16");
    static_assert(LN == 16 || LN == 32 || LN == 64, "128 MMA rows = G * 2 * LN");
    static_assert(!V64 || (32 == LM && 32 == LN), "64 x 64 blocks run as 32 x 32 sub-blocks");
    using S = Tc16Shape<LM, LN>;
    constexpr int KS = S::KS, N = S::N, NB = S::NB, NCH = S::NCH, NS = S::NS, RA = S::RA, SC = S::SC;
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
    uint64_t *const bar_acc_full = bar_x_full + NS;                          // [2]  a segment's MMAs have completed
    uint64_t *const bar_acc_free = bar_acc_full + 2;                         // [2]  the epilogue has read the set
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 1008);
    unsigned char *const ring = smem_raw + 1024;
    float *const exch0 = reinterpret_cast<float*>(smem_raw + 1024 + size_t(RA)*S::ABYTES);
    static_assert((2*RA + NS + 4)*8 <= 1008, "barrier area");
    static_assert(RA >= NS, "one ring of completion barriers serves the A slots and the X stages");

    int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    int const gs = a.gstride;

    if (kMmaWarp0 == w) tmem_alloc(tmem_slot, S::TMEM_COLS);
    if (0 == tid) {
        for (int r = 0; r < RA; ++r) { mbar_init(&bar_a_full[r], 1); mbar_init(&bar_done[r], 1); }
        for (int s = 0; s < NS; ++s) mbar_init(&bar_x_full[s], kWarpsPerGroup);
        for (int c = 0; c < 2; ++c)  { mbar_init(&bar_acc_full[c], 1); mbar_init(&bar_acc_free[c], kEpiWarps); }
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
    int const seg = a.seg;

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
                    if (TFQ_TC16P_ABLATE & 8) mbar_arrive(&bar_a_full[r]);
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
    } else if (w >= kMmaWarp0) {
        // ================= MMA warps: one elected lane of each issues; warp mw takes the segments with set mw ================
        uint32_t const mw = uint32_t(w - kMmaWarp0);
        uint32_t const leader = elect_one_sync();
        uint32_t const ring_u32 = smem_u32(ring);
        uint32_t n = 0, sg = 0;
        uint32_t e_begin = E0;
        uint32_t e_end_next = (u0 < u1) ? a.unit_e0[u0 + 1] : E0;
        TR_DECL;
        for (uint32_t u = u0; u < u1; ++u) {
            uint32_t const e_end = e_end_next;
            if (u + 1 < u1) e_end_next = a.unit_e0[u + 2];          // one unit ahead: not on the critical path
            int const nE = int(e_end - e_begin);
            e_begin = e_end;
            int const nSeg = (nE + seg - 1)/seg;
            for (int s = 0; s < nSeg; ++s, ++sg) {
                int const len = (nE*(s + 1))/nSeg - (nE*s)/nSeg;
                uint32_t const c = sg & 1, cuse = sg >> 1;
                if (c != mw) { n += uint32_t(len); continue; }          // the other MMA warp's segment
                TR_LAP(3);
                if (cuse > 0) { mbar_wait(&bar_acc_free[c], (cuse - 1) & 1); tc_fence_after(); }
                TR_LAP(2);
                uint32_t const acc = tmem_base + S::ACC0 + c*NB;
                // The correction sum D' of a set lives for the whole unit (its truncation errors carry the factor 1/2048); the main
                // sum D starts afresh with every segment.  First use of the set in this unit: one MMA of N' = 2N initialises both.
                bool const fresh = (s < 2);
                for (int t = 0; t < len; ++t, ++n) {
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
                        for (int ks = 0; ks < ((TFQ_TC16P_ABLATE & 4) ? 0 : KS); ++ks) {
                            uint64_t const b = smem_desc_noswizzle(sa + ks*2*S::SLAB, S::SLAB, 128);
                            if (0 == ks && 0 == t && !fresh) {
                                uint64_t const blo = smem_desc_noswizzle(sa + N*16, S::SLAB, 128);           // rows [N, 2N): Alo
                                mma_f16_ts(acc,     xa, b,   IDESC_N, 0u);                                    // Xhi * Ahi  -> D (restart)
                                mma_f16_ts(acc + N, xa, blo, IDESC_N, 1u);                                    // Xhi * Alo  -> D' (continue)
                            } else {
                                mma_f16_ts(acc, xa + 8*ks, b, IDESC_NB, (t > 0 || ks > 0) ? 1u : 0u);         // Xhi * [Ahi ; Alo] -> [D | D']
                            }
                            mma_f16_ts(acc + N, xa + LM/2 + 8*ks, b, IDESC_N, 1u);                            // Xlo * Ahi         ->      D'
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
        // ================= converter warps: X operand rows -> registers -> tensor memory =================================
        // Group grp takes the entries n = grp (mod 2) of the CTA's flat entry range; its 4 warps cover the 128 TMEM lanes
        // (one operand row each: NCH 16-byte chunks = hi | lo).  Loads run two own entries (four entries) ahead.
        int const grp = w / kWarpsPerGroup, q4 = w & 3;
        int const m = 32*q4 + lane;
        int const g = m/(2*LN);                          // warp-uniform for LN = 16, 32, 64
        uint32_t const row = uint32_t(m % (2*LN));
        bool const has_g = (g < gs);
        uint32_t const lane_base = tmem_base + (uint32_t(32*q4) << 16) + S::STAGE0;
        uint32_t const nOwn = (total > uint32_t(grp)) ? (total - grp + kConvGroups - 1)/kConvGroups : 0u;

        TR_DECL;
        auto index_of = [&](uint32_t k) -> uint32_t {    // X block of own entry k (kNoBlock: structural zero)
            return (has_g && k < nOwn) ? a.ent_x[size_t(E0 + grp + kConvGroups*k)*gs + g] : kNoBlock;
        };
        auto load = [&](uint32_t ix, uint4 (&b)[NCH]) {
            if (kNoBlock != ix && !(TFQ_TC16P_ABLATE & 1)) {
                uint4 const *src = a.xop + size_t(ix)*S::XCH + row;
                #pragma unroll
                for (int q = 0; q < NCH; ++q) b[q] = __ldg(src + q*S::XROWS);
            } else {
                #pragma unroll
                for (int q = 0; q < NCH; ++q) b[q] = make_uint4(0u, 0u, 0u, 0u);
            }
        };
        auto put = [&](uint32_t k, uint4 const (&b)[NCH]) {
            uint32_t const n = grp + kConvGroups*k;      // entry number within the CTA
            uint32_t const st = n & (NS - 1);
            TR_LAP(3);
            if (n >= uint32_t(NS)) { mbar_wait(&bar_done[(n - NS) & (RA - 1)], ((n - NS) / RA) & 1); tc_fence_after(); }   // the stage's previous entry
            TR_LAP(0);
            if (!(TFQ_TC16P_ABLATE & 2)) {
                uint32_t const *r = reinterpret_cast<uint32_t const*>(&b[0]);
                if (32 == LM) tmem_st32(lane_base + st*SC, r); else tmem_st16(lane_base + st*SC, r);
                tmem_wait_st();
            }
            tc_fence_before();
            __syncwarp();
            if (0 == lane) mbar_arrive(&bar_x_full[st]);
            TR_LAP(1);
        };

        if (nOwn > 0) {
            uint4 b0[NCH], b1[NCH];
            uint32_t i2 = index_of(2), i3 = index_of(3);
            load(index_of(0), b0);
            load(index_of(1), b1);
            for (uint32_t k = 0; k < nOwn; k += 2) {
                put(k, b0);
                { uint32_t const ix = i2; i2 = index_of(k + 4); load(ix, b0); }      // own entry k+2
                if (k + 1 < nOwn) {
                    put(k + 1, b1);
                    { uint32_t const ix = i3; i3 = index_of(k + 5); load(ix, b1); }  // own entry k+3
                }
            }
        }
        TR_LAP(3);
        if (0 == q4) TR_DUMP(3 + grp);
    } else {
        // ================= epilogue warps: accumulator segments -> registers (fp32 sums) -> Y ==========================
        int const ew = w - kConvWarps, q4 = ew & 3, h = ew >> 2;      // h: Re|Im of A = half of the accumulator columns
        int const m = 32*q4 + lane;
        int const g = m/(2*LN), cx = (m/LN) & 1, j = m % LN;
        uint32_t const lane_base = tmem_base + (uint32_t(32*q4) << 16) + S::ACC0 + uint32_t(h)*LM;
        uint32_t sg = 0;
        uint32_t e_begin = E0;
        uint32_t e_end_next = (u0 < u1) ? a.unit_e0[u0 + 1] : E0;
        TR_DECL;
        for (uint32_t u = u0; u < u1; ++u) {
            uint32_t const e_end = e_end_next;
            if (u + 1 < u1) e_end_next = a.unit_e0[u + 2];
            int const nE = int(e_end - e_begin);
            e_begin = e_end;
            int const nSeg = (nE + seg - 1)/seg;
            uint32_t const iy = (g < gs) ? a.unit_y[size_t(u)*gs + g] : kNoBlock;
            float acc[LM];
            #pragma unroll
            for (int i = 0; i < LM; ++i) acc[i] = 0.f;
            for (int s = 0; s < nSeg; ++s, ++sg) {
                uint32_t const c = sg & 1;
                TR_LAP(2);
                mbar_wait(&bar_acc_full[c], (sg >> 1) & 1);
                tc_fence_after();
                TR_LAP(0);
                bool const with_corr = (s + 2 >= nSeg);          // last segment of this unit in set c: its D' is complete
                #pragma unroll
                for (int ch = 0; ch < LM/16; ++ch) {
                    uint32_t d[16], d2[16];
                    tmem_ld16(lane_base + c*NB + 16*ch, d);
                    if (with_corr) tmem_ld16(lane_base + c*NB + N + 16*ch, d2);
                    tmem_wait_ld();
                    if (with_corr) {
                        #pragma unroll
                        for (int i = 0; i < 16; ++i)   // main + correction/2048, then the running sum: both rounded to nearest
                            acc[16*ch + i] += fmaf(__uint_as_float(d2[i]), 1.f/2048.f, __uint_as_float(d[i]));
                    } else {
                        #pragma unroll
                        for (int i = 0; i < 16; ++i) acc[16*ch + i] += __uint_as_float(d[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (0 == lane) mbar_arrive(&bar_acc_free[c]);
                TR_LAP(1);
            }
            // combine the four real products: Yr = XrAr - XiAi ; Yi = XrAi + XiAr.  Threads with Im(X) rows hand their sums
            // to the threads with the Re(X) rows of the same column j through shared memory (two buffers alternate by unit,
            // so one barrier per unit is enough)
            float *const exch = exch0 + size_t(u & 1)*(S::EXCH/4);
            if (1 == cx) {
                #pragma unroll
                for (int i = 0; i < LM; ++i) exch[((g*2 + h)*LM + i)*LN + j] = acc[i];
            }
            asm volatile("bar.sync 1, %0;" :: "n"(32*kEpiWarps) : "memory");
            if (0 == cx && kNoBlock != iy && !(TFQ_TC16P_ABLATE & 16)) {
                float const sgn = h ? 1.f : -1.f;
                float scale;
                float *yp;
                if (V64) {   // iy = 2*(Y block) + (row half); g = column half
                    uint32_t const yb = iy >> 1, ih = iy & 1;
                    scale = a.a_inv[a.unit_row[u]]*a.x_inv[size_t(a.blockcol[yb])*64 + 32*g + j];
                    yp = a.y + size_t(yb)*8192 + size_t(h)*4096 + size_t(32*ih)*64 + 32*g + j;
                } else {
                    scale = a.a_inv[a.unit_row[u]]*a.x_inv[size_t(a.blockcol[iy])*LN + j];
                    yp = a.y + size_t(iy)*(2*LM*LN) + size_t(h)*LM*LN + j;                 // plane h: 0 = Re, 1 = Im
                }
                constexpr int YS = V64 ? 64 : LN;
                #pragma unroll
                for (int i = 0; i < LM; ++i)
                    yp[i*YS] = (acc[i] + sgn*exch[((g*2 + (1 - h))*LM + i)*LN + j])*scale;
            }
            TR_LAP(2);
        }
        if (0 == ew) TR_DUMP(5);
    }
    tc_fence_before();
    __syncthreads();
    if (kMmaWarp0 == w) { tc_fence_after(); tmem_dealloc(tmem_base, S::TMEM_COLS); }
}

template <int LM, int LN, bool V64>
tfqmrgpuStatus_t launch_tc16p(Plan const &p, void *y, int expect, cudaStream_t stream)
{
    using S = Tc16Shape<LM, LN>;
    auto kernel = spmm_tc16p_kernel<LM, LN, V64>;
    static size_t configured[kMaxDevices] = {0}; // per instantiation and device
    TFQ_CUDA(ensure_dynamic_smem(kernel, S::smem, configured));
    Tc16Args a;
    a.y = static_cast<float*>(y);
    a.xop = ws<uint4 const>(p, p.off_xop); a.Aop = ws<unsigned char const>(p, p.off_A);
    a.a_inv = ws<float const>(p, p.off_ainv); a.x_inv = ws<float const>(p, p.off_xsinv);
    a.cta_u0 = p.d_cta_u0; a.unit_e0 = p.d_unit_e0; a.unit_y = p.d_unit_y; a.unit_row = p.d_unit_row;
    a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x; a.blockcol = p.d_blockcol;
    a.ctl = ws<Control const>(p, p.off_ctl); a.expect = expect; a.gstride = int(p.gmax);
    // Two MMA warps share the operand barriers: a warp that starts its segment tests barriers up to one foreign segment ahead of
    // what it has observed itself.  An mbarrier parity test cannot tell phase k from phase k + 2, so that look-ahead must stay
    // below the depth of the X stage ring (then the phase before the tested one is known to be complete): segment <= NS - 1.
    a.seg = std::min(p.tc_seg, S::NS - 1);
    if (p.tc_grid > 0 && p.nUnits > 0) kernel<<<p.tc_grid, kThreads, S::smem, stream>>>(a);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace

// entries per accumulation segment: 8 MMA k-steps (K = 16 each) per accumulator before the fp32 sum in registers
int spmm_tc16p_default_segment(int LM) { return (16 == LM) ? 8 : 4; }

// the product proper; the X operand (xop, scales) must have been produced from x by launch_xop (xop.cu)
tfqmrgpuStatus_t launch_spmm_tc16p(Plan const &p, void *y, int expect, cudaStream_t stream)
{
    switch (p.LM*1000 + p.LN) {
        case 16016: return launch_tc16p<16, 16, false>(p, y, expect, stream);
        case 16032: return launch_tc16p<16, 32, false>(p, y, expect, stream);
        case 16064: return launch_tc16p<16, 64, false>(p, y, expect, stream);
        case 32032: return launch_tc16p<32, 32, false>(p, y, expect, stream);
        case 32064: return launch_tc16p<32, 64, false>(p, y, expect, stream);
        case 64064: return launch_tc16p<32, 32, true >(p, y, expect, stream);
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
}

} // namespace tfq

#ifdef TFQ_TC16P_TRACE
extern "C" int tfq_tc16p_trace_dump(long long *host, int n) {
    return int(cudaMemcpyFromSymbol(host, tfq::g_tc16p_trace, size_t(n)*sizeof(long long)));
}
#endif
