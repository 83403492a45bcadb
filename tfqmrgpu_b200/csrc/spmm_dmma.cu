// Block-sparse product  Y = A * X  for complex fp64 on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64).
//
// Same role as spmm.cu (the reference's blocksparse_action_t::multiply + gemmNxNf,
// tfqmrgpu_blocksparse.hxx:71-199, tfqmrgpu_blockmult.hxx:10-93) for LM, LN in {16, 32, 64}.  tcgen05 has no fp64
// kind, and on B200 DMMA has the same peak as the FP64 FMA pipe - what it buys is operand reuse: the SIMT kernel
// needs 96 bytes of shared-memory operands per 32 FMAs and thread and is shared-memory bound at ~60 % of the fp64
// peak (measured 22.8 TFLOP/s at 32x32), a warp tile of 2 x 4 DMMA tiles needs 96 bytes per 256.
//
// One CTA per unit (a block row times G block columns, G*LN = 64 ... 256).  Per pipeline step one k-slab (KC rows) of
// the A block and of the unit's X blocks is staged by the bulk-copy engine, one copy per (operand, Re|Im) plane.
// The 64-bit fragment loads (4 k-rows x 8 columns per warp) are 4-way bank-conflicted in this unpadded layout; a
// padded layout needs one bulk copy per ROW, which was measured to be copy-engine bound and slower (47.5 ms vs
// 31.0 ms per product at 32x32, 128 RHS; the SIMT kernel: 40.6 ms).  A swizzled 2-D TMA box is the next step.
// Complex arithmetic: Yr += Ar*Xr + (-Ai)*Xi ; Yi += Ar*Xi + Ai*Xr  -> 4 DMMAs per (8x8 tile, 4 k).
// Fragments (PTX ISA, m8n8k4 .row.col): a = Amma[lane/4][lane%4], b = Bmma[lane%4][lane/4],
// c0,c1 = C[lane/4][2*(lane%4) + {0,1}] with Amma[i][k] = A[k][i] (A is stored transposed) and Bmma[k][j] = X[k][j].
#include "tfq_internal.hpp"
#include <algorithm>

namespace tfq {

namespace {

constexpr int kKC = 8;          // k rows per pipeline step
constexpr int kMaxStagesD = 4;

struct DmmaArgs {
    double *y; double const *x; double const *A; double const *zero;
    uint32_t const *unit_e0, *unit_y, *ent_a, *ent_x;
    Control const *ctl; int expect;
    int gmax, stages;
};

__device__ __forceinline__ uint32_t smem_u32d(void const *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init_d(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32d(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_d(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32d(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_d(uint64_t *bar, unsigned parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32d(bar)), "r"(parity) : "memory");
    return 0 != ok;
}
// bounded by wall time (%globaltimer, 4 s): a pipeline that never signals is a bug and must end the launch with an error instead of
// hanging the stream; a legitimately slow stage (profiler replay, managed memory migrating) is waited for
__device__ __forceinline__ void mbar_wait_d(uint64_t *bar, unsigned parity) {
    if (mbar_try_wait_d(bar, parity)) return;
    uint64_t t0 = 0;
    for (uint32_t spins = 1; !mbar_try_wait_d(bar, parity); ++spins) {
        if (0 == (spins & 0xfffu)) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (0 == t0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive_d(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32d(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_d(void *dst_smem, void const *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32d(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32d(bar)) : "memory");
}
// D(8x8) += A(8x4) * B(4x8), fp64
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int LM, int LN>
__global__ void __launch_bounds__(256)
spmm_dmma_kernel(DmmaArgs const a)
{
    static_assert(LM % 16 == 0 && LN % 16 == 0, "warp tile = 16 rows x 32 columns");
    constexpr int WR = LM/16;                       // warps along the rows
    constexpr int CH = LM/kKC;                      // pipeline steps per entry
    constexpr int SA = LM, SX = LN;                 // row strides of the staged slabs (doubles)

    if (a.expect >= 0 && a.ctl->state != a.expect) return; // device-resident solver control

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *const bars = reinterpret_cast<uint64_t*>(smem_raw);          // [stages] slab landed
    uint64_t *const bars_free = bars + kMaxStagesD;                        // [stages] all warps of the CTA are done reading it
    double *const stage0 = reinterpret_cast<double*>(smem_raw + 128);
    __shared__ uint32_t s_y[16];
    __shared__ int s_ng;
    constexpr int kEntCache = 48;
    __shared__ uint32_t s_ent_a[kEntCache];
    __shared__ uint32_t s_ent_x[kEntCache*16];

    int const G = a.gmax, nStages = a.stages;
    int const stageElems = 2*kKC*SA + G*2*kKC*SX;   // [A re rows][A im rows][g: X re rows, X im rows]
    uint32_t const u = blockIdx.x;
    uint32_t const e0 = a.unit_e0[u];
    int const nE = int(a.unit_e0[u + 1] - e0);
    int const nSteps = nE*CH;
    int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    int const wr = w % WR, wc = w / WR;             // this warp: rows [16 wr, +16), columns [32 wc, +32) of the unit tile

    if (tid < 16) s_y[tid] = (tid < G) ? a.unit_y[size_t(u)*G + tid] : kNoBlock;
    {
        int const nC = (nE < kEntCache) ? nE : kEntCache;
        for (int q = tid; q < nC; q += blockDim.x) s_ent_a[q] = a.ent_a[e0 + q];
        for (int q = tid; q < nC*G; q += blockDim.x) s_ent_x[q] = a.ent_x[size_t(e0)*G + q];
    }
    if (0 == tid) {
        for (int s = 0; s < nStages; ++s) { mbar_init_d(&bars[s], 1); mbar_init_d(&bars_free[s], blockDim.x >> 5); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (0 == tid) { int n = 0; for (int q = 0; q < G; ++q) n += (s_y[q] != kNoBlock); s_ng = n; }
    __syncthreads();
    int const ng = s_ng;

    // producer: warp 0
    auto issue = [&](int st) {
        int const s = st % nStages;
        int const e = st / CH, ch = st - e*CH;
        double *const dst = stage0 + size_t(s)*stageElems;
        uint32_t const ia = (e < kEntCache) ? s_ent_a[e] : a.ent_a[e0 + e];
        if (0 == lane) mbar_expect_tx_d(&bars[s], unsigned((2*kKC*LM + ng*2*kKC*LN)*sizeof(double)));
        __syncwarp();
        for (int c = lane; c < 2 + 2*ng; c += 32) {   // contiguous slabs: one copy per (operand, Re|Im) plane
            if (c < 2) {
                double const *src = a.A + (size_t(ia)*2 + c)*LM*LM + size_t(ch)*kKC*LM;
                bulk_g2s_d(dst + c*kKC*SA, src, unsigned(kKC*LM*sizeof(double)), &bars[s]);
            } else {
                int const gg = (c - 2) >> 1, ri = (c - 2) & 1;
                uint32_t const ix = (e < kEntCache) ? s_ent_x[e*G + gg] : a.ent_x[size_t(e0 + e)*G + gg];
                double const *base = (kNoBlock == ix) ? a.zero : a.x + size_t(ix)*2*LM*LN;
                double const *src = base + size_t(ri)*LM*LN + size_t(ch)*kKC*LN;
                bulk_g2s_d(dst + 2*kKC*SA + (gg*2 + ri)*kKC*SX, src, unsigned(kKC*LN*sizeof(double)), &bars[s]);
            }
        }
    };

    // accumulators: 2 row tiles x 4 column tiles, Re and Im
    double cre[2][4][2], cim[2][4][2];
    #pragma unroll
    for (int m = 0; m < 2; ++m)
        #pragma unroll
        for (int n = 0; n < 4; ++n) { cre[m][n][0] = cre[m][n][1] = 0; cim[m][n][0] = cim[m][n][1] = 0; }

    // column tiles of this warp: tile n covers unit columns [32 wc + 8 n, +8) -> block column gt[n], j offset jt[n]
    int gt[4], jt[4];
    #pragma unroll
    for (int n = 0; n < 4; ++n) { int const col = 32*wc + 8*n; gt[n] = col / LN; jt[n] = col % LN; }
    bool const warp_active = (gt[0] < ng);          // (a 32-column group never straddles the end of the unit's columns)

    if (w == 0) for (int st = 0; st < nStages - 1 && st < nSteps; ++st) issue(st);

    int const fr = lane >> 2, fk = lane & 3;        // fragment coordinates of this lane
    // LM >= 32: no CTA-wide barrier per step (it was 15 % of the stall samples at 32x32; config-4 shard 30.97 -> 29.59 ms per
    // product); at LM = 16 the steps are too short for the extra arrive/wait to pay (262 -> 279 us on the sweep): __syncthreads
    constexpr bool kFreeBars = (LM >= 32);
    for (int st = 0; st < nSteps; ++st) {
        int const s = st % nStages;
        if (!kFreeBars && w == 0 && st + nStages - 1 < nSteps) issue(st + nStages - 1);   // that slot was drained at the end of step st-1
        // every warp waits for the slab, also one without columns: it must not run ahead and report a slot free twice in one phase
        mbar_wait_d(&bars[s], unsigned((st/nStages) & 1));
        if (warp_active) {
            double const *const As_re = stage0 + size_t(s)*stageElems;
            double const *const As_im = As_re + kKC*SA;
            double const *const Xs = As_re + 2*kKC*SA;
            #pragma unroll
            for (int k4 = 0; k4 < kKC/4; ++k4) {
                int const kr = 4*k4 + fk;
                double ar[2], ai[2], nai[2];
                #pragma unroll
                for (int m = 0; m < 2; ++m) {
                    int const i = 16*wr + 8*m + fr;
                    ar[m] = As_re[kr*SA + i]; ai[m] = As_im[kr*SA + i]; nai[m] = -ai[m];
                }
                #pragma unroll
                for (int n = 0; n < 4; ++n) {
                    if (gt[n] < ng) {
                        double const *const xg = Xs + size_t(gt[n])*2*kKC*SX;
                        double const xr = xg[kr*SX + jt[n] + fr], xi = xg[(kKC + kr)*SX + jt[n] + fr];
                        #pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            dmma(cre[m][n], ar[m],  xr);
                            dmma(cre[m][n], nai[m], xi);
                            dmma(cim[m][n], ar[m],  xi);
                            dmma(cim[m][n], ai[m],  xr);
                        }
                    }
                }
            }
        }
        if (kFreeBars) {
            // every warp reports the stage free, and the producer warp - after its own share of step st - refills the slot of
            // step st-1 once all warps have left it
            __syncwarp();
            if (0 == lane) mbar_arrive_d(&bars_free[s]);
            if (w == 0 && st + nStages - 1 < nSteps) {
                if (st >= 1) mbar_wait_d(&bars_free[(st - 1) % nStages], unsigned(((st - 1)/nStages) & 1));
                issue(st + nStages - 1);
            }
        } else {
            __syncthreads(); // everyone is done with stage s before it is refilled
        }
    }

    if (warp_active) {
        #pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (gt[n] < ng) {
                uint32_t const iy = s_y[gt[n]];
                double *const yre = a.y + size_t(iy)*2*LM*LN, *const yim = yre + LM*LN;
                #pragma unroll
                for (int m = 0; m < 2; ++m) {
                    int const i = 16*wr + 8*m + fr, j = jt[n] + 2*fk;
                    *reinterpret_cast<double2*>(yre + i*LN + j) = make_double2(cre[m][n][0], cre[m][n][1]);
                    *reinterpret_cast<double2*>(yim + i*LN + j) = make_double2(cim[m][n][0], cim[m][n][1]);
                }
            }
        }
    }
}

template <int LM, int LN>
tfqmrgpuStatus_t launch_d(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    int const G = int(p.gmax);
    size_t const stageBytes = (2*size_t(kKC)*LM + size_t(G)*2*kKC*LN)*sizeof(double);
    int const stages = int(std::min<size_t>(kMaxStagesD, std::max<size_t>(2, (96*1024)/stageBytes)));
    size_t const smem = 128 + stages*stageBytes;
    int const warps = (LM/16)*((G*LN + 31)/32);
    if (warps > 8 || warps < 1) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    auto kernel = spmm_dmma_kernel<LM, LN>;
    static size_t configured[kMaxDevices] = {0}; // per instantiation and device
    TFQ_CUDA(ensure_dynamic_smem(kernel, smem, configured));
    DmmaArgs a;
    a.y = static_cast<double*>(y); a.x = static_cast<double const*>(x);
    a.A = ws<double const>(p, p.off_A); a.zero = ws<double const>(p, p.off_zero);
    a.unit_e0 = p.d_unit_e0; a.unit_y = p.d_unit_y; a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x;
    a.ctl = ws<Control const>(p, p.off_ctl); a.expect = expect; a.gmax = G; a.stages = stages;
    if (p.nUnits > 0) kernel<<<p.nUnits, 32*warps, smem, stream>>>(a);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace

// measured against the SIMT kernel on the 12^3 stencil with 64 RHS: 16x32/16x64 1.2x, 32xN 1.4x, 64x64 1.5x - but 0.93x at
// 16x16 (0.7x on the ragged rows of the reference's plan_unordered.14-287-16), which therefore stays on the SIMT kernel
bool spmm_dmma_supported(int LM, int LN, char precision) {
    return ('z' == precision) && (16 == LM || 32 == LM || 64 == LM) && (16 == LN || 32 == LN || 64 == LN) && (LM <= LN)
        && !(16 == LM && 16 == LN);
}
// block columns per unit: 8 warps of 16 x 32 outputs -> G*LN = 256 / (LM/16), G*LN a multiple of 32
int spmm_dmma_columns_per_unit(int LM, int LN) {
    int const cols = 256/(LM/16);
    return std::max(1, cols/LN);
}

tfqmrgpuStatus_t launch_spmm_dmma(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    switch (p.LM*1000 + p.LN) {
#define TFQ_CASE(LM, LN) case LM*1000 + LN: return launch_d<LM, LN>(p, y, x, expect, stream);
        TFQ_CASE(16, 16) TFQ_CASE(16, 32) TFQ_CASE(16, 64)
        TFQ_CASE(32, 32) TFQ_CASE(32, 64)
        TFQ_CASE(64, 64)
#undef TFQ_CASE
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
}

} // namespace tfq
