// Half-precision operand pairs for the tcgen05 block-sparse product (spmm_tc16.cu).
//
// The product D = sum_k Xop[m][k] * Aop[n][k] runs on fp16 tensor-core inputs with fp32 accumulation.  An fp32 value v
// is carried as  v * s = hi + lo/2048  with a power-of-two scale s, hi = fp16(v*s) and lo = fp16((v*s - hi) * 2048)
// (round to nearest twice: |v*s - hi - lo/2048| <= 2^-24 |v*s|, the size of ONE fp32 rounding), in the same 4 bytes per
// element.  The scale brings the largest magnitude of
//   * a block row of A          (all blocks of the row share the accumulators of its Y blocks), resp.
//   * a right-hand-side column  (block column c, lane j: all blocks of X in that column feed the same accumulator lane)
// into [2^14, 2^15), so nothing overflows in fp16 and elements down to 2^-29 of the largest keep their full 22 bits
// (below that they lose bits gradually - an absolute error of 2^-50 of the row/column maximum).
//
// Layouts (what the MMA and the converter warps of spmm_tc16.cu read).  The K dimension of the product carries Re and Im
// interleaved, k' = (k, Re|Im), so that the accumulators hold Y itself (spmm_tc16.cu):
//   A operand block  [LM/4 k'-octets][2*LM rows n = (hi|lo, i)][8 halves = (Re, Im) of 4 consecutive k]  = the K-major
//                    no-swizzle core-matrix layout of a tcgen05 shared-memory descriptor: ONE bulk copy per block, no work
//                    in the product loop;
//   X operand block  [LM/2 chunks q][LN rows j][8 halves = (Re, Im) of 4 consecutive k]  chunk q < LM/4: hi of k = 4q .. 4q+3,
//                    then the lo chunks: a warp reads 32 consecutive rows of one chunk = 512 contiguous bytes per instruction.
//   64 x 64 blocks are stored as 2 x 2 sub-blocks of 32 x 32 in these layouts: A sub-block (ia, i/32, k/32),
//   X sub-block (ix, k/32, j/32).
// The PLANAR form of the product (spmm_tc16p.cu, selected per plan) keeps Re and Im in separate operand rows instead:
//   A operand block  [LM/8 k-octets][4*LM rows n = (hi|lo, Re|Im, i)][8 halves k%8],
//   X operand block  [LM/4 chunks q][2*LN rows (Re|Im, j)][8 halves]  (hi chunks, then lo chunks).
// Role in the reference: none - its product reads fp32 blocks directly (tfqmrgpu_blockmult.hxx:10-93).
#include "tfq_internal.hpp"
#include <cuda_fp16.h>
#include <cfloat>

namespace tfq {

namespace {

// power-of-two scale that maps a maximum magnitude mx into [2^14, 2^15); 1 for mx = 0, inf or nan
__device__ __forceinline__ float scale_for(float mx) {
    if (!(mx > 0.f) || !(mx <= FLT_MAX)) return 1.f;
    int ex;
    frexpf(mx, &ex);                       // mx = f * 2^ex, f in [0.5, 1)
    int e = 15 - ex;
    e = (e > 120) ? 120 : ((e < -120) ? -120 : e);   // scale and its inverse stay normal fp32 numbers
    return ldexpf(1.f, e);
}

__device__ __forceinline__ void split_half(float vs, __half &hi, __half &lo) {
    hi = __float2half_rn(vs);
    lo = __float2half_rn((vs - __half2float(hi))*2048.f);
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
    return uint32_t(__half_as_ushort(a)) | (uint32_t(__half_as_ushort(b)) << 16);
}

// ---- X operand -------------------------------------------------------------------------------------------------
// pass 1: per-column maximum magnitude.  One CTA per vector tile (a contiguous range of blocks of ONE block column);
// the last tile of a column (ticket counter) folds the tile maxima and writes the scale and its inverse.
__global__ void __launch_bounds__(256)
xop_absmax_kernel(float const *__restrict__ x, Tile const *__restrict__ tiles, uint32_t const *__restrict__ coltile,
                  float *__restrict__ part, unsigned *__restrict__ ticket, float *__restrict__ xs, float *__restrict__ xsinv,
                  int LM, int LN, Control const *ctl, int expect)
{
    if (expect >= 0 && ctl->state != expect) return;
    __shared__ float red[256];
    __shared__ int s_last;
    Tile const t = tiles[blockIdx.x];
    int const tid = threadIdx.x;
    int const LNV = LN/4;                             // LN in {16, 32, 64}: float4 loads, 256 % LNV == 0
    int const jv = tid % LNV, r0 = tid / LNV, rstep = 256/LNV;
    size_t const base = size_t(t.b0)*2*LM*LN;
    int const nrows = int(t.b1 - t.b0)*2*LM;           // rows of LN numbers (Re and Im planes alike)
    float mx[4] = {0.f, 0.f, 0.f, 0.f};
    #pragma unroll 4
    for (int rr = r0; rr < nrows; rr += rstep) {
        float4 const v = *reinterpret_cast<float4 const*>(x + base + size_t(rr)*LN + 4*jv);
        mx[0] = fmaxf(mx[0], fabsf(v.x)); mx[1] = fmaxf(mx[1], fabsf(v.y));
        mx[2] = fmaxf(mx[2], fabsf(v.z)); mx[3] = fmaxf(mx[3], fabsf(v.w));
    }
    // threads with the same jv: tid = r0*LNV + jv
    #pragma unroll
    for (int v = 0; v < 4; ++v) {
        red[tid] = mx[v];
        __syncthreads();
        for (int half = rstep >> 1; half > 0; half >>= 1) {
            if (r0 < half) red[tid] = fmaxf(red[tid], red[tid + half*LNV]);
            __syncthreads();
        }
        if (0 == r0) part[size_t(blockIdx.x)*64 + 4*jv + v] = red[jv];
        __syncthreads();
    }
    uint32_t const c = t.col, t0 = coltile[c], t1 = coltile[c + 1];
    __threadfence();
    __syncthreads();
    if (0 == tid) s_last = (atomicAdd(&ticket[c], 1u) == (t1 - t0) - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (0 == tid) ticket[c] = 0;
    if (tid < LN) {
        float m = 0.f;
        for (uint32_t tt = t0; tt < t1; ++tt) m = fmaxf(m, __ldcg(&part[size_t(tt)*64 + tid]));
        float const s = scale_for(m);
        xs[size_t(c)*LN + tid] = s;
        xsinv[size_t(c)*LN + tid] = 1.f/s;            // exact: a power of two
    }
}

// pass 2: fp32 block [Re|Im][k][j] -> operand block.  One thread per operand row j: reads are coalesced over j,
// every 16-byte chunk (Re and Im of 4 k values) is written next to the chunks of the neighbouring rows.
template <int LM, int LN>
__global__ void __launch_bounds__(256)
xop_convert_kernel(float const *__restrict__ x, uint4 *__restrict__ xop, float const *__restrict__ xs,
                   uint32_t const *__restrict__ blockcol, uint32_t nnzb, Control const *ctl, int expect)
{
    if (expect >= 0 && ctl->state != expect) return;
    constexpr int BPC = 256/LN;                        // blocks per CTA
    uint32_t const b = blockIdx.x*BPC + threadIdx.x/LN;
    if (b >= nnzb) return;
    int const j = threadIdx.x % LN;
    float const s = xs[size_t(blockcol[b])*LN + j];
    float const *src = x + size_t(b)*2*LM*LN + j;
    constexpr size_t plane = size_t(LM)*LN;
    #pragma unroll 4
    for (int kq = 0; kq < LM/4; ++kq) {                // 4 consecutive k
        __half hr[4], lr[4], hi[4], li[4];
        #pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            split_half(src[size_t(4*kq + kk)*LN]*s, hr[kk], lr[kk]);
            split_half(src[plane + size_t(4*kq + kk)*LN]*s, hi[kk], li[kk]);
        }
        uint4 const vh = make_uint4(pack2(hr[0], hi[0]), pack2(hr[1], hi[1]), pack2(hr[2], hi[2]), pack2(hr[3], hi[3]));
        uint4 const vl = make_uint4(pack2(lr[0], li[0]), pack2(lr[1], li[1]), pack2(lr[2], li[2]), pack2(lr[3], li[3]));
        if (64 == LM) {    // 2 x 2 sub-blocks (k/32, j/32), each in the 32 x 32 layout: 16 chunks x 32 rows
            int const kh = kq >> 3, q = kq & 7, jh = j >> 5;
            uint4 *dst = xop + (size_t(b)*4 + kh*2 + jh)*(16*32) + size_t(q)*32 + (j & 31);
            dst[0] = vh; dst[8*32] = vl;
        } else {
            uint4 *dst = xop + size_t(b)*((LM/2)*LN) + size_t(kq)*LN + j;
            dst[0] = vh; dst[(LM/4)*LN] = vl;
        }
    }
}

// pass 2, PLANAR form (spmm_tc16p.cu): rows (Re|Im, j), chunks of 8 k values of one plane; fp32 block [Re|Im][k][j] -> operand block.  One thread per operand row (Re|Im, j): reads are coalesced over j,
// every 16-byte chunk (8 k values) is written next to the chunks of the neighbouring rows.
template <int LM, int LN>
__global__ void __launch_bounds__(256)
xop_convert_planar_kernel(float const *__restrict__ x, uint4 *__restrict__ xop, float const *__restrict__ xs,
                   uint32_t const *__restrict__ blockcol, uint32_t nnzb, Control const *ctl, int expect)
{
    if (expect >= 0 && ctl->state != expect) return;
    constexpr int ROWS = 2*LN, BPC = 256/ROWS;        // operand rows per block, blocks per CTA
    uint32_t const b = blockIdx.x*BPC + threadIdx.x/ROWS;
    if (b >= nnzb) return;
    int const r = threadIdx.x % ROWS, c = r / LN, j = r % LN;
    float const s = xs[size_t(blockcol[b])*LN + j];
    float const *src = x + size_t(b)*2*LM*LN + size_t(c)*LM*LN + j;
    if (64 == LM) {
        // 2 x 2 sub-blocks (k/32, j/32), each in the 32 x 32 layout: 8 chunks x 64 rows
        int const jh = j >> 5, r32 = c*32 + (j & 31);
        #pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
            uint4 *dst = xop + (size_t(b)*4 + kh*2 + jh)*(8*64) + r32;
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                __half hi[8], lo[8];
                #pragma unroll
                for (int kk = 0; kk < 8; ++kk) split_half(src[size_t(32*kh + 8*q + kk)*LN]*s, hi[kk], lo[kk]);
                dst[q*64]       = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
                dst[(q + 4)*64] = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
            }
        }
    } else {
        uint4 *dst = xop + size_t(b)*((LM/4)*ROWS) + r;
        #pragma unroll
        for (int q = 0; q < LM/8; ++q) {
            __half hi[8], lo[8];
            #pragma unroll
            for (int kk = 0; kk < 8; ++kk) split_half(src[size_t(8*q + kk)*LN]*s, hi[kk], lo[kk]);
            dst[q*ROWS]          = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
            dst[(q + LM/8)*ROWS] = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
        }
    }
}

// ---- A operand -------------------------------------------------------------------------------------------------
// maximum magnitude of every block (internal fp32 layout; runs behind the layout conversion of each uploaded chunk)
__global__ void __launch_bounds__(256)
aop_blockmax_kernel(float const *__restrict__ A, float *__restrict__ blockmax, int blockElems)
{
    __shared__ float red[8];
    float const *src = A + size_t(blockIdx.x)*blockElems;
    float m = 0.f;
    for (int q = 4*threadIdx.x; q < blockElems; q += 4*256) {
        float4 const v = *reinterpret_cast<float4 const*>(src + q);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (0 == (threadIdx.x & 31)) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (0 == threadIdx.x) {
        #pragma unroll
        for (int q = 1; q < 8; ++q) m = fmaxf(m, red[q]);
        blockmax[blockIdx.x] = m;
    }
}

// scale of every block row: 1/scale goes behind the A blocks (read by the product's epilogue), the scale itself into a scratch
__global__ void aop_rowscale_kernel(float const *__restrict__ blockmax, int32_t const *__restrict__ rowptr, int row0, int row1,
                                    float *__restrict__ rowscale, float *__restrict__ ainv)
{
    int const r = row0 + blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= row1) return;
    float m = 0.f;
    for (int i = rowptr[r]; i < rowptr[r + 1]; ++i) m = fmaxf(m, blockmax[i]);
    float const s = scale_for(m);
    rowscale[r] = s;
    ainv[r] = 1.f/s;
}

// fp32 block [Re|Im][k][i] (the reference's transposed internal layout) -> operand block, in place (one CTA per block,
// staged in shared memory)
template <int LM>
__global__ void __launch_bounds__(256)
aop_convert_kernel(float *__restrict__ A, int32_t const *__restrict__ rowptr, int mb, float const *__restrict__ rowscale, uint32_t b0)
{
    extern __shared__ __align__(16) float tmp[];                 // [2][LM][LM]
    __shared__ float s_scale;
    uint32_t const b = b0 + blockIdx.x;
    float *const blk = A + size_t(b)*2*LM*LM;
    for (int q = 4*threadIdx.x; q < 2*LM*LM; q += 4*256)
        *reinterpret_cast<float4*>(tmp + q) = *reinterpret_cast<float4 const*>(blk + q);
    if (0 == threadIdx.x) {                                       // block row of this block: last r with rowptr[r] <= b
        int lo = 0, hi = mb;
        while (hi - lo > 1) { int const mid = (lo + hi) >> 1; if (uint32_t(rowptr[mid]) <= b) lo = mid; else hi = mid; }
        s_scale = rowscale[lo];
    }
    __syncthreads();
    float const s = s_scale;
    uint4 *const dst = reinterpret_cast<uint4*>(blk);
    constexpr int SUB = (64 == LM) ? 32 : LM;                     // edge of a (sub-)block
    constexpr int NSUB = LM/SUB;
    // one item = (k-quad, i): Re and Im of 4 k values -> one hi chunk and one lo chunk of operand row i
    for (int it = threadIdx.x; it < LM*(LM/4); it += 256) {
        int const i = it % LM, kq = it / LM;                      // consecutive threads: consecutive i
        __half hr[4], lr[4], hi[4], li[4];
        #pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            split_half(tmp[(4*kq + kk)*LM + i]*s, hr[kk], lr[kk]);
            split_half(tmp[(LM + 4*kq + kk)*LM + i]*s, hi[kk], li[kk]);
        }
        int const ih = i / SUB, kh = (4*kq) / SUB, o = kq % (SUB/4), is = i % SUB;
        size_t const sub = (size_t(ih)*NSUB + kh)*(size_t(SUB/4)*2*SUB);     // uint4 elements of one sub-block: k'-octets x rows
        size_t const at = sub + size_t(o)*2*SUB + is;                        // rows: [hi: i][lo: i]
        dst[at]       = make_uint4(pack2(hr[0], hi[0]), pack2(hr[1], hi[1]), pack2(hr[2], hi[2]), pack2(hr[3], hi[3]));
        dst[at + SUB] = make_uint4(pack2(lr[0], li[0]), pack2(lr[1], li[1]), pack2(lr[2], li[2]), pack2(lr[3], li[3]));
    }
}

// PLANAR form (spmm_tc16p.cu): rows n = (hi|lo, Re|Im, i), chunks of 8 k values of one plane.
// fp32 block [Re|Im][k][i] (the reference's transposed internal layout) -> operand block, in place (one CTA per block,
// staged in shared memory)
template <int LM>
__global__ void __launch_bounds__(256)
aop_convert_planar_kernel(float *__restrict__ A, int32_t const *__restrict__ rowptr, int mb, float const *__restrict__ rowscale, uint32_t b0)
{
    extern __shared__ __align__(16) float tmp[];                 // [2][LM][LM]
    __shared__ float s_scale;
    uint32_t const b = b0 + blockIdx.x;
    float *const blk = A + size_t(b)*2*LM*LM;
    for (int q = 4*threadIdx.x; q < 2*LM*LM; q += 4*256)
        *reinterpret_cast<float4*>(tmp + q) = *reinterpret_cast<float4 const*>(blk + q);
    if (0 == threadIdx.x) {                                       // block row of this block: last r with rowptr[r] <= b
        int lo = 0, hi = mb;
        while (hi - lo > 1) { int const mid = (lo + hi) >> 1; if (uint32_t(rowptr[mid]) <= b) lo = mid; else hi = mid; }
        s_scale = rowscale[lo];
    }
    __syncthreads();
    float const s = s_scale;
    uint4 *const dst = reinterpret_cast<uint4*>(blk);
    constexpr int SUB = (64 == LM) ? 32 : LM;                     // edge of a (sub-)block
    constexpr int NSUB = LM/SUB;
    // one item = (sub-block, k-octet, Re|Im, i): 8 k values -> one hi chunk and one lo chunk
    for (int it = threadIdx.x; it < 2*LM*(LM/8); it += 256) {
        int const i = it % LM, c = (it / LM) & 1, ko = it / (2*LM);          // consecutive threads: consecutive i
        __half hi[8], lo[8];
        #pragma unroll
        for (int kk = 0; kk < 8; ++kk) split_half(tmp[(c*LM + 8*ko + kk)*LM + i]*s, hi[kk], lo[kk]);
        int const ih = i / SUB, kh = (8*ko) / SUB, kos = ko % (SUB/8), is = i % SUB;
        size_t const sub = (size_t(ih)*NSUB + kh)*(size_t(SUB/8)*4*SUB);     // uint4 elements of one sub-block: k-octets x rows
        size_t const at = sub + size_t(kos)*4*SUB + size_t(c)*SUB + is;      // rows: [hi: Re i, Im i][lo: Re i, Im i]
        dst[at]         = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
        dst[at + 2*SUB] = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
    }
}

} // namespace

// X operand of the tensor-core product from the storage-ordered fp32 vector x: scales, then the half pairs
tfqmrgpuStatus_t launch_xop(Plan const &p, void const *x, int expect, cudaStream_t stream)
{
    if (p.nnzbX < 1) return TFQMRGPU_STATUS_SUCCESS;
    Control const *ctl = ws<Control const>(p, p.off_ctl);
    float const *xf = static_cast<float const*>(x);
    xop_absmax_kernel<<<p.nTiles, 256, 0, stream>>>(xf, p.d_tiles, p.d_coltile, ws<float>(p, p.off_xpart),
        ws<unsigned>(p, p.off_ticket), ws<float>(p, p.off_xs), ws<float>(p, p.off_xsinv), p.LM, p.LN, ctl, expect);
    uint4 *xop = ws<uint4>(p, p.off_xop);
    float const *xs = ws<float const>(p, p.off_xs);
    uint32_t const n = uint32_t(p.nnzbX);
#define TFQ_XOP(LM, LN) case LM*1000 + LN: if (p.tc_planar) { constexpr int BPC = 256/(2*LN); \
        xop_convert_planar_kernel<LM, LN><<<(n + BPC - 1)/BPC, 256, 0, stream>>>(xf, xop, xs, p.d_blockcol, n, ctl, expect); \
    } else { constexpr int BPC = 256/LN; \
        xop_convert_kernel<LM, LN><<<(n + BPC - 1)/BPC, 256, 0, stream>>>(xf, xop, xs, p.d_blockcol, n, ctl, expect); } break;
    switch (p.LM*1000 + p.LN) {
        TFQ_XOP(16, 16) TFQ_XOP(16, 32) TFQ_XOP(16, 64) TFQ_XOP(32, 32) TFQ_XOP(32, 64) TFQ_XOP(64, 64)
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
#undef TFQ_XOP
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

// block maxima of `nb` freshly uploaded and layout-converted A blocks starting at block `b0`
tfqmrgpuStatus_t launch_aop_blockmax(Plan const &p, uint32_t b0, uint32_t nb, cudaStream_t stream)
{
    if (nb < 1) return TFQMRGPU_STATUS_SUCCESS;
    int const blockElems = 2*p.LM*p.LM;
    aop_blockmax_kernel<<<nb, 256, 0, stream>>>(ws<float const>(p, p.off_A) + size_t(b0)*blockElems,
                                                 ws<float>(p, p.off_ablkmax) + b0, blockElems);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

// row scales from the block maxima, then the A blocks of the block rows [row0, row1) -> operand blocks in place
tfqmrgpuStatus_t launch_aop_convert_rows(Plan const &p, int row0, int row1, cudaStream_t stream)
{
    if (p.nnzbA < 1 || row1 <= row0) return TFQMRGPU_STATUS_SUCCESS;
    if (p.h_rpA.size() != size_t(p.mb) + 1) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    uint32_t const b0 = uint32_t(p.h_rpA[row0]), nb = uint32_t(p.h_rpA[row1]) - b0;
    float *const rowscale = ws<float>(p, p.off_arowscale);
    aop_rowscale_kernel<<<(row1 - row0 + 255)/256, 256, 0, stream>>>(ws<float const>(p, p.off_ablkmax), p.d_rowptrA, row0, row1,
                                                                      rowscale, ws<float>(p, p.off_ainv));
    if (nb < 1) { TFQ_CUDA(cudaGetLastError()); return TFQMRGPU_STATUS_SUCCESS; }
    float *const A = ws<float>(p, p.off_A);
    size_t const smem = 2*size_t(p.LM)*p.LM*sizeof(float);
    switch (p.LM + (p.tc_planar ? 1000 : 0)) {
        case 16: aop_convert_kernel<16><<<nb, 256, smem, stream>>>(A, p.d_rowptrA, p.mb, rowscale, b0); break;
        case 32: aop_convert_kernel<32><<<nb, 256, smem, stream>>>(A, p.d_rowptrA, p.mb, rowscale, b0); break;
        case 64: aop_convert_kernel<64><<<nb, 256, smem, stream>>>(A, p.d_rowptrA, p.mb, rowscale, b0); break;
        case 1016: aop_convert_planar_kernel<16><<<nb, 256, smem, stream>>>(A, p.d_rowptrA, p.mb, rowscale, b0); break;
        case 1032: aop_convert_planar_kernel<32><<<nb, 256, smem, stream>>>(A, p.d_rowptrA, p.mb, rowscale, b0); break;
        case 1064: aop_convert_planar_kernel<64><<<nb, 256, smem, stream>>>(A, p.d_rowptrA, p.mb, rowscale, b0); break;
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t launch_aop_convert(Plan const &p, cudaStream_t stream) { return launch_aop_convert_rows(p, 0, p.mb, stream); }

} // namespace tfq
