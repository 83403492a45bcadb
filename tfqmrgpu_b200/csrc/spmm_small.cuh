// Device code of the small-block product (LM <= 8) shared by the per-kernel launch (spmm.cu) and the resident solver
// (resident.cu), and the arguments of the SIMT product kernels.
#pragma once
#include "tfq_internal.hpp"
#include <type_traits>
#include <algorithm>

namespace tfq {

namespace {

template <typename real_t> struct SpmmArgs {
    real_t *y; real_t const *x; real_t const *A; real_t const *zero;
    uint32_t const *unit_e0, *unit_y, *ent_a, *ent_x;
    Control const *ctl; int expect;
    int gmax, kc, stages;
};

template <typename T, int N>
__device__ __forceinline__ void load_vec(T (&r)[N], T const *p) {
    if constexpr ((sizeof(T)*N) % 16 == 0) {
        #pragma unroll
        for (int q = 0; q < int(sizeof(T)*N/16); ++q) {
            float4 const v = reinterpret_cast<float4 const*>(p)[q];
            reinterpret_cast<float4*>(r)[q] = v;
        }
    } else if constexpr ((sizeof(T)*N) % 8 == 0) {
        #pragma unroll
        for (int q = 0; q < int(sizeof(T)*N/8); ++q) {
            float2 const v = reinterpret_cast<float2 const*>(p)[q];
            reinterpret_cast<float2*>(r)[q] = v;
        }
    } else {
        #pragma unroll
        for (int q = 0; q < N; ++q) r[q] = p[q];
    }
}
template <typename T, int N>
__device__ __forceinline__ void store_vec(T *p, T const (&r)[N]) {
    if constexpr ((sizeof(T)*N) % 16 == 0) {
        #pragma unroll
        for (int q = 0; q < int(sizeof(T)*N/16); ++q) reinterpret_cast<float4*>(p)[q] = reinterpret_cast<float4 const*>(r)[q];
    } else if constexpr ((sizeof(T)*N) % 8 == 0) {
        #pragma unroll
        for (int q = 0; q < int(sizeof(T)*N/8); ++q) reinterpret_cast<float2*>(p)[q] = reinterpret_cast<float2 const*>(r)[q];
    } else {
        #pragma unroll
        for (int q = 0; q < N; ++q) p[q] = r[q];
    }
}

// ---- small blocks (LM <= 8) ----------------------------------------------------------------------------------------
// Blocks of 128 B ... 1 KiB: the bulk-copy engine retires a copy every ~50 cycles whatever its size, so a ring of per-block
// copies is copy-engine bound (12 CTAs x 17 copies x 27 entries per SM = 143 us for the 4x4 sweep case, measured 97-150 us
// per product).  Here the threads stage the operands themselves: 128-bit coalesced global loads of a whole BATCH of
// entries into registers, one shared-memory store + one __syncthreads per batch (double buffer), and the entries of a batch
// are shared out over kSplit thread groups that each keep the usual TI x TJ accumulator tile.  Group p always takes the
// unit's entries e = p (mod kSplit) and the partial sums are added in the order ((S0 + S1) + S2) + S3, whatever the
// batch size or the number of block columns of the unit: a column's bits do not depend on how the columns are sharded.
// 128 threads and <= 40 KiB per CTA: 4-5 CTAs per SM (a first version with 256 threads, 159 registers and 87 KiB ran ONE
// CTA per SM and lost to the ring kernel).
constexpr int kSplit = 4;              // entry groups per CTA
constexpr int kSmallThreads = 128;
constexpr int kSmallNV = 10;           // 128-bit loads in flight per thread and batch

// One unit of the product.  u: the CTA index of the per-kernel launch (spmm.cu); the resident solver (resident.cu) walks its CTAs
// over the units and, because X is written by other CTAs of the SAME launch there, reads X past the L1 (COHERENT).
template <typename real_t, int LM, int LN, int TI, int TJ, bool COHERENT>
__device__ __forceinline__ void spmm_small_unit(SpmmArgs<real_t> const &a, uint32_t const u, unsigned char *const smem_raw)
{
    __shared__ uint32_t s_y[16];
    __shared__ int s_ng;
    constexpr int kEntCache = 64;
    __shared__ uint32_t s_ent_a[kEntCache];
    __shared__ uint32_t s_ent_x[kEntCache*16];

    int const G = a.gmax, EB = a.kc;                 // kc carries the entries per batch here
    constexpr int aF4 = int(2*LM*LM*sizeof(real_t)/16), xF4 = int(2*LM*LN*sizeof(real_t)/16);
    static_assert((2*LM*LM*sizeof(real_t)) % 16 == 0 && (2*LM*LN*sizeof(real_t)) % 16 == 0, "blocks are whole float4s");
    static_assert(xF4 < 65536, "offset field of the load code");
    // an X block of 128 ... 512 B per thread column: one float4 of padding per block keeps the 128-bit operand loads of the
    // threads of a group (one block column each when LN <= TJ) off each other's banks
    constexpr int xSlot = xF4 + 1;
    uint32_t const e0 = a.unit_e0[u];
    int const nE = int(a.unit_e0[u + 1] - e0);
    int const tid = threadIdx.x;

    if (tid < 16) s_y[tid] = (tid < G) ? a.unit_y[size_t(u)*G + tid] : kNoBlock;
    {
        int const nC = (nE < kEntCache) ? nE : kEntCache;
        for (int q = tid; q < nC; q += kSmallThreads) s_ent_a[q] = a.ent_a[e0 + q];
        for (int q = tid; q < nC*G; q += kSmallThreads) s_ent_x[q] = a.ent_x[size_t(e0)*G + q];
    }
    __syncthreads();
    if (0 == tid) { int n = 0; for (int q = 0; q < G; ++q) n += (s_y[q] != kNoBlock); s_ng = n; }
    __syncthreads();
    int const ng = s_ng;
    int const entryF4 = aF4 + ng*xF4;                // float4s to load per entry
    int const entrySlot = aF4 + G*xSlot;             // float4s of one staged entry: [A re, im][g: X re, im, pad]
    int const bufF4 = EB*entrySlot;
    float4 *const stage0 = reinterpret_cast<float4*>(smem_raw);   // two buffers

    // what this thread loads in every batch: float4 number q = tid + v*threads of the batch, coded as
    // entry-in-batch (8 bits) | block column + 1, 0 = the A block (8 bits) | float4 inside the block (16 bits); ~0: nothing
    uint32_t code[kSmallNV];
    #pragma unroll
    for (int v = 0; v < kSmallNV; ++v) {
        int const q = tid + v*kSmallThreads;
        int const eb = q / entryF4, r = q - eb*entryF4;
        int const gg = (r < aF4) ? -1 : (r - aF4)/xF4;
        int const rr = (r < aF4) ? r : (r - aF4) - gg*xF4;
        code[v] = (eb < EB) ? ((uint32_t(eb) << 24) | (uint32_t(gg + 1) << 16) | uint32_t(rr)) : ~0u;
    }
    float4 const *const A4 = reinterpret_cast<float4 const*>(a.A);
    float4 const *const X4 = reinterpret_cast<float4 const*>(a.x);

    // thread tile over the ng block columns this unit really has, times kSplit entry groups
    int const NTJ = (ng*LN)/TJ;
    int const T = (LM/TI)*((G*LN)/TJ);               // threads of one entry group (kSplit*T <= 128)
    int const grp = tid / T, t = tid - grp*T;
    int const tj = (NTJ > 0) ? t % NTJ : 0, ti = (NTJ > 0) ? t / NTJ : LM;
    int const g = (tj*TJ)/LN, j0 = (tj*TJ) % LN, i0 = ti*TI;
    bool const active = (grp < kSplit) && (ti < LM/TI) && (g < ng);

    real_t acc_re[TI][TJ], acc_im[TI][TJ];
    #pragma unroll
    for (int ii = 0; ii < TI; ++ii) {
        #pragma unroll
        for (int jj = 0; jj < TJ; ++jj) { acc_re[ii][jj] = 0; acc_im[ii][jj] = 0; }
    }

    int const nBatches = (nE + EB - 1)/EB;
    float4 reg[kSmallNV];
#define TFQ_SMALL_FETCH(batch)                                                                                        \
    _Pragma("unroll")                                                                                                 \
    for (int v = 0; v < kSmallNV; ++v) {                                                                              \
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);                                                                 \
        int const e = (batch)*EB + int(code[v] >> 24);                                                                \
        if (~0u != code[v] && e < nE) {                                                                               \
            int const gg = int((code[v] >> 16) & 0xff) - 1, rr = int(code[v] & 0xffff);                               \
            if (gg < 0) {                                                                                             \
                uint32_t const ia = (e < kEntCache) ? s_ent_a[e] : a.ent_a[e0 + e];                                   \
                val = __ldg(A4 + size_t(ia)*aF4 + rr);                                                                \
            } else {                                                                                                  \
                uint32_t const ix = (e < kEntCache) ? s_ent_x[e*G + gg] : a.ent_x[size_t(e0 + e)*G + gg];             \
                if (kNoBlock != ix) val = COHERENT ? __ldcg(X4 + size_t(ix)*xF4 + rr) : __ldg(X4 + size_t(ix)*xF4 + rr);      \
            }                                                                                                         \
        }                                                                                                             \
        reg[v] = val;                                                                                                 \
    }
    if (nBatches > 0) { TFQ_SMALL_FETCH(0) }
    for (int b = 0; b < nBatches; ++b) {
        float4 *const buf = stage0 + (b & 1)*bufF4;
        #pragma unroll
        for (int v = 0; v < kSmallNV; ++v) {
            if (~0u != code[v]) {
                int const eb = int(code[v] >> 24), gg = int((code[v] >> 16) & 0xff) - 1, rr = int(code[v] & 0xffff);
                buf[eb*entrySlot + ((gg < 0) ? rr : aF4 + gg*xSlot + rr)] = reg[v];
            }
        }
        __syncthreads();                             // batch b is in place; everyone has left batch b-1 (other buffer: b-2)
        if (b + 1 < nBatches) { TFQ_SMALL_FETCH(b + 1) }   // in flight during the FMAs
        if (active) {
            int const nb = (nE - b*EB < EB) ? (nE - b*EB) : EB;
            // entries e = grp (mod kSplit) of the UNIT; EB is a multiple of kSplit, so eb = grp, grp + kSplit, ...
            for (int eb = grp; eb < nb; eb += kSplit) {
                real_t const *const As_re = reinterpret_cast<real_t const*>(buf + eb*entrySlot);
                real_t const *const As_im = As_re + LM*LM;
                real_t const *const Xs_re = reinterpret_cast<real_t const*>(buf + eb*entrySlot + aF4 + g*xSlot);
                real_t const *const Xs_im = Xs_re + LM*LN;
                #pragma unroll
                for (int kk = 0; kk < LM; ++kk) {
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
        }
    }
#undef TFQ_SMALL_FETCH
    __syncthreads();                                 // the stages are free: they now take the partial sums of groups 1..3
    real_t *const part = reinterpret_cast<real_t*>(smem_raw);     // [kSplit - 1][T][2][TI][TJ]
    if (active && grp > 0) {
        real_t *const dst = part + (size_t(grp - 1)*T + t)*2*TI*TJ;
        #pragma unroll
        for (int ii = 0; ii < TI; ++ii) {
            #pragma unroll
            for (int jj = 0; jj < TJ; ++jj) { dst[ii*TJ + jj] = acc_re[ii][jj]; dst[TI*TJ + ii*TJ + jj] = acc_im[ii][jj]; }
        }
    }
    __syncthreads();
    if (active && 0 == grp) {
        #pragma unroll
        for (int pp = 1; pp < kSplit; ++pp) {
            real_t const *const src = part + (size_t(pp - 1)*T + t)*2*TI*TJ;
            #pragma unroll
            for (int ii = 0; ii < TI; ++ii) {
                #pragma unroll
                for (int jj = 0; jj < TJ; ++jj) { acc_re[ii][jj] += src[ii*TJ + jj]; acc_im[ii][jj] += src[TI*TJ + ii*TJ + jj]; }
            }
        }
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


// entries per batch and dynamic shared memory of the small-block product for units of up to G block columns; false: units too
// large to batch (the ring kernel takes them)
template <typename real_t, int LM, int LN>
inline bool spmm_small_config(int G, int &eb, size_t &smem) {
    constexpr bool is_double = std::is_same<real_t, double>::value;
    constexpr int TI = spmm_ti(is_double, LM, LN);
    constexpr int TJ = spmm_tj(is_double, LN);
    int const T = (LM/TI)*((G*LN)/TJ);
    if (T < 1) return false;
    // what the threads can hold in kSmallNV 128-bit registers each, a multiple of kSplit
    size_t const entryF4 = (2*size_t(LM)*LM + size_t(G)*2*LM*LN)*sizeof(real_t)/16;
    eb = int((size_t(kSmallThreads)*kSmallNV)/entryF4);
    eb = std::min((eb/kSplit)*kSplit, 32);
    if (eb < kSplit || kSplit*T > kSmallThreads || G > 16) return false;
    size_t const entrySlot = 2*size_t(LM)*LM*sizeof(real_t)/16 + size_t(G)*(2*size_t(LM)*LN*sizeof(real_t)/16 + 1);
    size_t const part = size_t(kSplit - 1)*T*2*TI*TJ*sizeof(real_t);
    smem = std::max(2*size_t(eb)*entrySlot*16, part);
    return true;
}

} // namespace
} // namespace tfq
