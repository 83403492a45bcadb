// Block-sparse product  Y = A * X  over complex LM x LM A-blocks and LM x LN X-blocks.
//
// Role of the reference's blocksparse_action_t::multiply + gemmNxNf (tfqmrgpu_blocksparse.hxx:71-199,
// tfqmrgpu_blockmult.hxx:10-93), re-designed for sm_100a:
//   * one CTA per "unit" = one block row of A times up to gmax of that row's block columns, so every
//     A block is fetched ONCE per unit and reused across all its right-hand-side columns (the
//     reference re-reads A for every Y block);
//   * A and X k-slabs are staged in shared memory by the bulk-copy engine (cp.async.bulk, i.e. TMA
//     without a tensor map; SASS: UBLKCP) into a 3-stage ring completed through mbarriers, so copy and
//     FMA work overlap and there are no per-k __syncthreads (the reference has two per k);
//   * each thread owns a TI x TJ register tile of complex accumulators and reads its operands with
//     128-bit shared-memory loads (A broadcast across the j-threads, X conflict-free).
// Internal layouts: A[nnzbA][2][LM(k)][LM(i)] (transposed), X,Y[nnzb][2][LM][LN] in storage order.
// Arithmetic: accumulators in real_t over all pairs and k like blockmult.hxx:28-82.
#include "spmm_small.cuh"
#include <type_traits>
#include <algorithm>
#include <cstdlib>

namespace tfq {

namespace {

constexpr int kMaxStages = 8;   // ring depth is chosen per block size (small blocks are latency-bound: deeper ring)

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes: multiple of 16, 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, void const *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename real_t, int LM, int LN, int TI, int TJ>
__global__ void __launch_bounds__(256)
spmm_unit_kernel(SpmmArgs<real_t> const a)
{
    if (a.expect >= 0 && a.ctl->state != a.expect) return; // device-resident solver control

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *const bars = reinterpret_cast<uint64_t*>(smem_raw);
    real_t *const stage0 = reinterpret_cast<real_t*>(smem_raw + 128);
    __shared__ uint32_t s_y[16];
    __shared__ int s_ng;
    // entry indices of the unit, fetched once with coalesced loads: a per-step global index load would put an
    // L2/HBM round trip in front of every bulk copy (measured: 1.4 us per entry on the 8x8 FD example)
    constexpr int kEntCache = 48;
    __shared__ uint32_t s_ent_a[kEntCache];
    __shared__ uint32_t s_ent_x[kEntCache*16];

    int const G = a.gmax, KC = a.kc, kStages = a.stages;
    int const CH = LM/KC;                            // k-chunks per entry
    int const stageElems = 2*KC*(LM + G*LN);         // [A re][A im][g: X re, X im]
    uint32_t const u = blockIdx.x;
    uint32_t const e0 = a.unit_e0[u];
    int const nE = int(a.unit_e0[u + 1] - e0);
    int const nSteps = nE*CH;

    int const tid = threadIdx.x;

    if (tid < 16) s_y[tid] = (tid < G) ? a.unit_y[size_t(u)*G + tid] : kNoBlock;
    {
        int const nC = (nE < kEntCache) ? nE : kEntCache;
        for (int q = tid; q < nC; q += blockDim.x) s_ent_a[q] = a.ent_a[e0 + q];
        for (int q = tid; q < nC*G; q += blockDim.x) s_ent_x[q] = a.ent_x[size_t(e0)*G + q];
    }
    if (0 == tid) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (0 == tid) { int n = 0; for (int q = 0; q < G; ++q) n += (s_y[q] != kNoBlock); s_ng = n; }
    __syncthreads();
    int const ng = s_ng;
    // thread tile mapping over the ng block columns this unit really has (ragged rows: whole warps idle, not lanes)
    int const NTJ = (ng*LN)/TJ;
    int const tj = (NTJ > 0) ? tid % NTJ : 0, ti = (NTJ > 0) ? tid / NTJ : LM;
    int const g = (tj*TJ)/LN, j0 = (tj*TJ) % LN, i0 = ti*TI;
    bool const active = (ti < LM/TI) && (g < ng);

    // producer: warp 0 issues the bulk copies of one pipeline step
    auto issue = [&](int st) {
        int const s = st % kStages;
        int const e = st / CH, ch = st - e*CH;
        real_t *const dst = stage0 + size_t(s)*stageElems;
        uint32_t const ia = (e < kEntCache) ? s_ent_a[e] : a.ent_a[e0 + e];
        int const lane = tid;
        if (0 == lane) mbar_expect_tx(&bars[s], unsigned(2*KC*(LM + ng*LN)*sizeof(real_t)));
        __syncwarp();
        if (KC == LM) {
            // the whole block per step: its Re and Im planes are contiguous in memory AND in the stage, so ONE copy per
            // operand (the copy engine retires small copies at a fixed ~50 cycles each; this halves their number)
            for (int c = lane; c < 1 + ng; c += 32) {
                if (0 == c) {
                    bulk_g2s(dst, a.A + size_t(ia)*2*LM*LM, unsigned(2*LM*LM*sizeof(real_t)), &bars[s]);
                } else {
                    int const gg = c - 1;
                    uint32_t const ix = (e < kEntCache) ? s_ent_x[e*G + gg] : a.ent_x[size_t(e0 + e)*G + gg];
                    real_t const *src = (kNoBlock == ix) ? a.zero : a.x + size_t(ix)*2*LM*LN;   // (the zero block is a full block)
                    bulk_g2s(dst + 2*KC*LM + gg*2*KC*LN, src, unsigned(2*LM*LN*sizeof(real_t)), &bars[s]);
                }
            }
            return;
        }
        for (int c = lane; c < 2 + 2*ng; c += 32) {
            if (c < 2) {
                real_t const *src = a.A + (size_t(ia)*2 + c)*LM*LM + size_t(ch)*KC*LM;
                bulk_g2s(dst + c*KC*LM, src, unsigned(KC*LM*sizeof(real_t)), &bars[s]);
            } else {
                int const gg = (c - 2) >> 1, ri = (c - 2) & 1;
                uint32_t const ix = (e < kEntCache) ? s_ent_x[e*G + gg] : a.ent_x[size_t(e0 + e)*G + gg];
                real_t const *base = (kNoBlock == ix) ? a.zero : a.x + size_t(ix)*2*LM*LN;
                real_t const *src = base + size_t(ri)*LM*LN + size_t(ch)*KC*LN;
                bulk_g2s(dst + 2*KC*LM + (gg*2 + ri)*KC*LN, src, unsigned(KC*LN*sizeof(real_t)), &bars[s]);
            }
        }
    };

    real_t acc_re[TI][TJ], acc_im[TI][TJ];
    #pragma unroll
    for (int ii = 0; ii < TI; ++ii) {
        #pragma unroll
        for (int jj = 0; jj < TJ; ++jj) { acc_re[ii][jj] = 0; acc_im[ii][jj] = 0; }
    }

    if (tid < 32) {
        for (int st = 0; st < kStages - 1 && st < nSteps; ++st) issue(st);
    }

    for (int st = 0; st < nSteps; ++st) {
        int const s = st % kStages;
        if (tid < 32 && st + kStages - 1 < nSteps) issue(st + kStages - 1); // the slot was drained at the end of step st-1
        if (active) {
            mbar_wait(&bars[s], unsigned((st/kStages) & 1));
            real_t const *const As_re = stage0 + size_t(s)*stageElems;
            real_t const *const As_im = As_re + KC*LM;
            real_t const *const Xs_re = As_re + 2*KC*LM + (g*2)*KC*LN;
            real_t const *const Xs_im = Xs_re + KC*LN;
            #pragma unroll 4
            for (int kk = 0; kk < KC; ++kk) {
                real_t ar[TI], ai[TI], xr[TJ], xi[TJ];
                load_vec<real_t, TI>(ar, As_re + kk*LM + i0);
                load_vec<real_t, TI>(ai, As_im + kk*LM + i0);
                load_vec<real_t, TJ>(xr, Xs_re + kk*LN + j0);
                load_vec<real_t, TJ>(xi, Xs_im + kk*LN + j0);
                #pragma unroll
                for (int ii = 0; ii < TI; ++ii) {
                    #pragma unroll
                    for (int jj = 0; jj < TJ; ++jj) {
                        // complex multiply-accumulate, 8 flop (blockmult.hxx:76-77)
                        acc_re[ii][jj] = fma( ar[ii], xr[jj], acc_re[ii][jj]);
                        acc_re[ii][jj] = fma(-ai[ii], xi[jj], acc_re[ii][jj]);
                        acc_im[ii][jj] = fma( ar[ii], xi[jj], acc_im[ii][jj]);
                        acc_im[ii][jj] = fma( ai[ii], xr[jj], acc_im[ii][jj]);
                    }
                }
            }
        }
        __syncthreads(); // everyone is done with stage s before it is refilled
    }

    if (active) {
        uint32_t const iy = s_y[g];
        real_t *const yre = a.y + size_t(iy)*2*LM*LN;
        real_t *const yim = yre + LM*LN;
        #pragma unroll
        for (int ii = 0; ii < TI; ++ii) {
            store_vec<real_t, TJ>(yre + (i0 + ii)*LN + j0, acc_re[ii]);
            store_vec<real_t, TJ>(yim + (i0 + ii)*LN + j0, acc_im[ii]);
        }
    }
}

template <typename real_t, int LM, int LN>
tfqmrgpuStatus_t launch_typed(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    constexpr bool is_double = std::is_same<real_t, double>::value;
    constexpr int TI = spmm_ti(is_double, LM, LN);
    constexpr int TJ = spmm_tj(is_double, LN);
    int const G = int(p.gmax);
    // k-chunk: largest power-of-two slab with a stage of at most 24 KiB
    int kc = LM;
    while (kc > 4 && 2*size_t(kc)*(LM + size_t(G)*LN)*sizeof(real_t) > 24*1024) kc >>= 1;
    size_t const stageBytes = 2*size_t(kc)*(LM + size_t(G)*LN)*sizeof(real_t);
    int const stages = int(std::min<size_t>(kMaxStages, std::max<size_t>(3, (48*1024)/stageBytes)));
    size_t const smem = 128 + stages*stageBytes;
    int threads = (LM/TI)*((G*LN)/TJ);
    threads = ((threads + 31)/32)*32;
    if (threads > 256) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);

    auto kernel = spmm_unit_kernel<real_t, LM, LN, TI, TJ>;
    static size_t configured[kMaxDevices] = {0}; // per instantiation and device
    TFQ_CUDA(ensure_dynamic_smem(kernel, smem, configured));
    SpmmArgs<real_t> a;
    a.y = static_cast<real_t*>(y); a.x = static_cast<real_t const*>(x);
    a.A = ws<real_t const>(p, p.off_A); a.zero = ws<real_t const>(p, p.off_zero);
    a.unit_e0 = p.d_unit_e0; a.unit_y = p.d_unit_y; a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x;
    a.ctl = ws<Control const>(p, p.off_ctl); a.expect = expect;
    a.gmax = G; a.kc = kc; a.stages = stages;
    if (p.nUnits > 0) kernel<<<p.nUnits, threads, smem, stream>>>(a);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}


template <typename real_t, int LM, int LN, int TI, int TJ>
__global__ void __launch_bounds__(kSmallThreads, 4)
spmm_small_kernel(SpmmArgs<real_t> const a)
{
    if (a.expect >= 0 && a.ctl->state != a.expect) return; // device-resident solver control
    extern __shared__ __align__(128) unsigned char smem_raw[];
    spmm_small_unit<real_t, LM, LN, TI, TJ, false>(a, blockIdx.x, smem_raw);
}

template <typename real_t, int LM, int LN>
tfqmrgpuStatus_t launch_small(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    constexpr bool is_double = std::is_same<real_t, double>::value;
    constexpr int TI = spmm_ti(is_double, LM, LN);
    constexpr int TJ = spmm_tj(is_double, LN);
    int const G = int(p.gmax);
    int eb = 0; size_t smem = 0;
    // entries too large to batch (or units made for the ring kernel's thread count): the ring kernel
    if (!spmm_small_config<real_t, LM, LN>(G, eb, smem)) return launch_typed<real_t, LM, LN>(p, y, x, expect, stream);
    auto kernel = spmm_small_kernel<real_t, LM, LN, TI, TJ>;
    static size_t configured[kMaxDevices] = {0}; // per instantiation and device
    TFQ_CUDA(ensure_dynamic_smem(kernel, smem, configured));
    SpmmArgs<real_t> a;
    a.y = static_cast<real_t*>(y); a.x = static_cast<real_t const*>(x);
    a.A = ws<real_t const>(p, p.off_A); a.zero = ws<real_t const>(p, p.off_zero);
    a.unit_e0 = p.d_unit_e0; a.unit_y = p.d_unit_y; a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x;
    a.ctl = ws<Control const>(p, p.off_ctl); a.expect = expect;
    a.gmax = G; a.kc = eb; a.stages = 2;
    if (p.nUnits > 0) kernel<<<p.nUnits, kSmallThreads, smem, stream>>>(a);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

template <int LM, int LN>
tfqmrgpuStatus_t launch_sized(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream) {
    if constexpr (LM <= 8) {
        if (p.use_small) {         // chosen with the unit size by plan_configure
            if ('z' == p.precision) return launch_small<double, LM, LN>(p, y, x, expect, stream);
            if ('c' == p.precision) return launch_small<float,  LM, LN>(p, y, x, expect, stream);
        }
    }
    if ('z' == p.precision) return launch_typed<double, LM, LN>(p, y, x, expect, stream);
    if ('c' == p.precision) return launch_typed<float,  LM, LN>(p, y, x, expect, stream);
    return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, p.precision); // 'm' is not implemented (tfqmrgpu.cu:42-44)
}

} // namespace

tfqmrgpuStatus_t launch_spmm_operand_ready(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    if (p.use_tc16 && nullptr == p.user_op) return p.tc_planar ? launch_spmm_tc16p(p, y, expect, stream) : launch_spmm_tc16(p, y, expect, stream);
    return launch_spmm(p, y, x, expect, stream);
}

tfqmrgpuStatus_t launch_spmm(Plan const &p, void *y, void const *x, int expect, cudaStream_t stream)
{
    if (p.user_op) {             // tfqmrgpux_bsrsv_setOperator: the caller's Y = A*X
        int32_t const st = p.user_op(p.user_ctx, y, x, reinterpret_cast<int32_t const*>(&ws<Control const>(p, p.off_ctl)->state), expect, stream);
        return st ? tfqmrgpuStatus_t(st) : TFQMRGPU_STATUS_SUCCESS;
    }
    if (p.use_tc16) {            // X operand once per product (scales + half pairs, xop.cu), then the tcgen05 product
        static bool const skip_xop = (nullptr != std::getenv("TFQMRGPU_DEV_SKIP_XOP"));   // dev-only: time the product kernel alone
        tfqmrgpuStatus_t const st = skip_xop ? TFQMRGPU_STATUS_SUCCESS : launch_xop(p, x, expect, stream);
        return st ? st : (p.tc_planar ? launch_spmm_tc16p(p, y, expect, stream) : launch_spmm_tc16(p, y, expect, stream));
    }
    if (p.use_dmma) return launch_spmm_dmma(p, y, x, expect, stream);
    switch (p.LM*1000 + p.LN) {
#define TFQ_CASE(LM, LN) case LM*1000 + LN: return launch_sized<LM, LN>(p, y, x, expect, stream);
        TFQ_CASE( 4,  4) TFQ_CASE( 4,  5) TFQ_CASE( 4,  8) TFQ_CASE( 4, 32)
        TFQ_CASE( 8,  8) TFQ_CASE( 8,  9) TFQ_CASE( 8, 10) TFQ_CASE( 8, 32) TFQ_CASE( 8, 64)
        TFQ_CASE(16, 16) TFQ_CASE(16, 32) TFQ_CASE(16, 64)
        TFQ_CASE(32, 32) TFQ_CASE(32, 64)
        TFQ_CASE(64, 64)
#undef TFQ_CASE
        default: return TFQMRGPU_BLOCKSIZE_MISSING + TFQMRGPU_CODE_CHAR*p.LM + TFQMRGPU_CODE_LINE*p.LN; // tfqmrgpu.cu:70
    }
}

} // namespace tfq
