// Fused per-right-hand-side-column tfQMR vector algebra.
//
// Role of the reference's dotp/nrm2/col_inner/col_reduction, axpy/xpay/col_axpay, tfQMRdec35/34/T and
// of the host-side convergence logic (tfqmrgpu_linalg.hxx:34-254,480-704, tfqmrgpu_core.hxx:189-304).
// The reference runs 26 + 4*ceil(log2 nnzbX) launches and two blocking device->host copies per
// iteration; here one iteration is 6 streaming kernels (plus 2 block-sparse products):
//
//   K1  v6 = v5 + beta v6
//   E1  v4 = v9 + beta (v8 + beta v4)              z34 = v3.v4  -> dec34 (alfa, c67)
//   K2  v7 = v6 + c67 v7 ; v5 += alfa v9           d55 = |v5|^2 -> decT  (var, tau, eta, c67 := r67)
//   K3  v1 += eta v7 ; v6 += alfa v4 ; v7 = v6 + c67 v7
//   E2  v5 += alfa v8                              d55 = |v5|^2 -> decT  (var, tau, eta)
//   K4  v1 += eta v7                               z35 = v3.v5  -> convergence monitor, dec35 (beta, rho)
//   N3  (probe only)                               |A v1 - b|^2 -> probe evaluation
//
// X-shaped vectors are stored column-sorted, so a CTA ("tile") streams a contiguous range of blocks of
// ONE block column: its coefficients are CTA-uniform registers, the per-column sums are reduced in
// registers -> shared memory -> one partial per tile, and the LAST tile of a column (ticket counter)
// sums the partials in a fixed order and evaluates the scalar recurrences.  The last column of K4/N3
// evaluates the reference's probe rule on the device, so no host round trip is needed per iteration.
// Everything is deterministic (no floating-point atomics).
#include "vec_body.cuh"

namespace tfq {

namespace {


// Column-sharded runs (several GPUs, each with a range of the right-hand-side block columns): the iteration count and the probe
// schedule of the reference are GLOBAL (one maximum over all right-hand sides, core.hxx:239-299).  K4 / N3 of a shard then only
// export their partial (max, count, count) and, once every shard's triple is there, every shard takes the same decision.
// kind 0: after K4, kind 1: after N3.  slots: [nShards][4] doubles of this kind and iteration parity.
__global__ void decide_kernel(Control *ctl, double const *slots, int nShards, long long nRHS, int kind) {
    if (ctl->state != ((0 == kind) ? STATE_RUN : STATE_PROBE)) return;
    double m0 = 0, m1 = 0, m2 = 0;
    for (int s = 0; s < nShards; ++s) {
        double const x0 = *reinterpret_cast<double const volatile*>(&slots[4*s + 0]);
        m0 = (m0 < x0) ? x0 : m0;
        m1 += *reinterpret_cast<double const volatile*>(&slots[4*s + 1]);
        m2 += *reinterpret_cast<double const volatile*>(&slots[4*s + 2]);
    }
    if (0 == kind) decide_iteration(*ctl, m0, m1, m2, nRHS); else decide_probe(*ctl, m0, m1);
}

template <typename real_t, int VEC, int OP, bool XM>
__global__ void __launch_bounds__(256)
vec_kernel(VecArgs<real_t> const a)
{
    if (OpTraits<OP>::expect >= 0 && a.ctl->state != OpTraits<OP>::expect) return; // device-resident control
    extern __shared__ double red[];
    vec_tile<real_t, VEC, OP, XM>(a, blockIdx.x, red);
}

// ---- K1 / K3 with the tensor-core operand -----------------------------------------------------------
// Both write v6, the vector the next block-sparse product multiplies.  For plans on the fp16-pair tensor-core product they
// also emit v6 as that product's X operand (layout and number format: xop.cu), so that the operand costs one more vector
// WRITE per product instead of a pass of its own.  The column scale comes from a bound on the result's magnitude,
//   K1: max|v6'| <= max|v5| + (|Re beta| + |Im beta|) max|v6|      K3: max|v6'| <= max|v6| + (|Re alfa| + |Im alfa|) max|v4|
// (maxima over Re and Im parts, kept per column by the kernels that wrote those vectors), so no second pass is needed; fp16
// being a floating-point format, a bound that is loose by a few binades costs range (29 binades are there), not precision.
// Thread = (k-octet, lane j): its 8 k values of Re and Im are four 16-byte chunks (two hi, two lo) of operand row j;
// all global accesses are coalesced over j.  The arithmetic statements are those of vec_kernel (reference: core.hxx:194,216-220).
__device__ __forceinline__ float xop_scale_for_bound(float bound) {
    float const b = bound*1.0001f;                       // (the bound itself was rounded)
    if (!(b > 0.f) || !(b <= 3.0e38f)) return 1.f;
    int ex;
    frexpf(b, &ex);
    int e = 15 - ex;
    e = (e > 120) ? 120 : ((e < -120) ? -120 : e);
    return ldexpf(1.f, e);
}
__device__ __forceinline__ uint32_t xop_pack2(__half a, __half b) {
    return uint32_t(__half_as_ushort(a)) | (uint32_t(__half_as_ushort(b)) << 16);
}

template <int OP, int LM, int LN, bool PLANAR>
__global__ void __launch_bounds__(256)
vec_xop_kernel(VecArgs<float> const a)
{
    static_assert(OP == OP_K1 || OP == OP_K3, "the kernels that write v6");
    if (a.ctl->state != STATE_RUN) return;
    constexpr int SUBS = 256/LN;                 // items in flight per CTA
    constexpr int KO = LM/8;                     // k-octets per block
    __shared__ float s_max[256];
    __shared__ int s_last;

    Tile const t = a.tiles[blockIdx.x];
    uint32_t const c = t.col;
    int const tid = threadIdx.x, j = tid % LN, sub = tid / LN;
    size_t const sj = (size_t(c)*2 + 0)*LN + j, sji = (size_t(c)*2 + 1)*LN + j;

    float pr = 0, pi = 0, qr = 0, qi = 0, sr = 0, si = 0, bound;
    if (OP == OP_K1) {
        pr = a.beta[sj]; pi = a.beta[sji];
        bound = a.mx5[size_t(c)*LN + j] + (fabsf(pr) + fabsf(pi))*a.mx6[size_t(c)*LN + j];
    } else {
        pr = a.eta[sj]; pi = a.eta[sji]; qr = a.alfa[sj]; qi = a.alfa[sji]; sr = a.c67[sj]; si = a.c67[sji];
        bound = a.mx6[size_t(c)*LN + j] + (fabsf(qr) + fabsf(qi))*a.mx4[size_t(c)*LN + j];
    }
    (void)qr; (void)qi; (void)sr; (void)si;
    float const scale = xop_scale_for_bound(bound);
    if (blockIdx.x == a.coltile[c] && 0 == sub) { a.xs[size_t(c)*LN + j] = scale; a.xsinv[size_t(c)*LN + j] = 1.f/scale; }

    float vmax = 0.f;
    int const nItems = int(t.b1 - t.b0)*KO;
    constexpr size_t plane = size_t(LM)*LN;
    for (int it = sub; it < nItems; it += SUBS) {
        uint32_t const blk = t.b0 + uint32_t(it / KO);
        int const ko = it % KO;
        size_t const ore = (size_t(blk)*2*LM + 8*ko)*LN + j, oim = ore + plane;
        float zr[8], zi[8];                       // the new v6
        if (OP == OP_K1) {    // v6 := v5 + beta*v6   (core.hxx:194, linalg.hxx:660-661)
            float xr[8], xi[8];
            #pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                xr[kk] = a.v5[ore + kk*LN]; xi[kk] = a.v5[oim + kk*LN];
                zr[kk] = a.v6[ore + kk*LN]; zi[kk] = a.v6[oim + kk*LN];
            }
            #pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                float const nr = xr[kk] + pr*zr[kk] - pi*zi[kk];
                float const ni = xi[kk] + pi*zr[kk] + pr*zi[kk];
                zr[kk] = nr; zi[kk] = ni;
            }
        } else {              // v1 += eta*v7 ; v6 += alfa*v4 ; v7 := v6 + c67*v7  (core.hxx:216-220)
            float yr[8], yi[8], xr[8], xi[8], br[8], bi[8];
            #pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                yr[kk] = a.v7[ore + kk*LN]; yi[kk] = a.v7[oim + kk*LN];
                xr[kk] = a.v1[ore + kk*LN]; xi[kk] = a.v1[oim + kk*LN];
                br[kk] = a.v4[ore + kk*LN]; bi[kk] = a.v4[oim + kk*LN];
                zr[kk] = a.v6[ore + kk*LN]; zi[kk] = a.v6[oim + kk*LN];
            }
            #pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                xr[kk] = pr*yr[kk] - pi*yi[kk] + xr[kk];
                xi[kk] = pi*yr[kk] + pr*yi[kk] + xi[kk];
                float const mr = qr*br[kk] - qi*bi[kk] + zr[kk];
                float const mi = qi*br[kk] + qr*bi[kk] + zi[kk];
                zr[kk] = mr; zi[kk] = mi;
                float const nr = mr + sr*yr[kk] - si*yi[kk];
                float const ni = mi + si*yr[kk] + sr*yi[kk];
                yr[kk] = nr; yi[kk] = ni;
            }
            #pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                a.v1[ore + kk*LN] = xr[kk]; a.v1[oim + kk*LN] = xi[kk];
                a.v7[ore + kk*LN] = yr[kk]; a.v7[oim + kk*LN] = yi[kk];
            }
        }
        if (PLANAR) {
        uint32_t hr[4], lr[4], hi_[4], li_[4];
        #pragma unroll
        for (int kk = 0; kk < 8; kk += 2) {
            a.v6[ore + kk*LN] = zr[kk]; a.v6[oim + kk*LN] = zi[kk];
            a.v6[ore + (kk + 1)*LN] = zr[kk + 1]; a.v6[oim + (kk + 1)*LN] = zi[kk + 1];
            vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(zr[kk]), fabsf(zi[kk])), fmaxf(fabsf(zr[kk + 1]), fabsf(zi[kk + 1]))));
            __half h0, l0, h1, l1;
            float v0 = zr[kk]*scale, v1 = zr[kk + 1]*scale;
            h0 = __float2half_rn(v0); l0 = __float2half_rn((v0 - __half2float(h0))*2048.f);
            h1 = __float2half_rn(v1); l1 = __float2half_rn((v1 - __half2float(h1))*2048.f);
            hr[kk/2] = xop_pack2(h0, h1); lr[kk/2] = xop_pack2(l0, l1);
            v0 = zi[kk]*scale; v1 = zi[kk + 1]*scale;
            h0 = __float2half_rn(v0); l0 = __float2half_rn((v0 - __half2float(h0))*2048.f);
            h1 = __float2half_rn(v1); l1 = __float2half_rn((v1 - __half2float(h1))*2048.f);
            hi_[kk/2] = xop_pack2(h0, h1); li_[kk/2] = xop_pack2(l0, l1);
        }
        uint4 *dre, *dim_; int lo_off;
        if (64 == LM) {       // 2 x 2 sub-blocks (k/32, j/32) in the 32 x 32 layout: 8 chunks x 64 rows each
            int const kh = ko >> 2, q = ko & 3, jh = j >> 5;
            uint4 *const base = a.xop + (size_t(blk)*4 + kh*2 + jh)*(8*64) + size_t(q)*64;
            dre = base + (j & 31); dim_ = base + 32 + (j & 31); lo_off = 4*64;
        } else {
            uint4 *const base = a.xop + size_t(blk)*(2*KO*2*LN) + size_t(ko)*(2*LN);
            dre = base + j; dim_ = base + LN + j; lo_off = KO*2*LN;
        }
        dre[0] = make_uint4(hr[0], hr[1], hr[2], hr[3]);       dim_[0] = make_uint4(hi_[0], hi_[1], hi_[2], hi_[3]);
        dre[lo_off] = make_uint4(lr[0], lr[1], lr[2], lr[3]);  dim_[lo_off] = make_uint4(li_[0], li_[1], li_[2], li_[3]);
            } else {
        // operand row j: (Re, Im) pairs of halves along k; this thread's 8 k values = two hi chunks and two lo chunks
        uint32_t hw[8], lw[8];
        #pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            a.v6[ore + kk*LN] = zr[kk]; a.v6[oim + kk*LN] = zi[kk];
            vmax = fmaxf(vmax, fmaxf(fabsf(zr[kk]), fabsf(zi[kk])));
            float const vr = zr[kk]*scale, vi = zi[kk]*scale;
            __half const hr = __float2half_rn(vr), hi = __float2half_rn(vi);
            __half const lr = __float2half_rn((vr - __half2float(hr))*2048.f), li = __float2half_rn((vi - __half2float(hi))*2048.f);
            hw[kk] = xop_pack2(hr, hi); lw[kk] = xop_pack2(lr, li);
        }
        uint4 *dst; int lo_off, rows;
        if (64 == LM) {       // 2 x 2 sub-blocks (k/32, j/32) in the 32 x 32 layout: 16 chunks x 32 rows each
            int const kh = ko >> 2, q = 2*(ko & 3), jh = j >> 5;
            dst = a.xop + (size_t(blk)*4 + kh*2 + jh)*(16*32) + size_t(q)*32 + (j & 31); lo_off = 8*32; rows = 32;
        } else {
            dst = a.xop + size_t(blk)*((LM/2)*LN) + size_t(2*ko)*LN + j; lo_off = (LM/4)*LN; rows = LN;
        }
        dst[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);             dst[rows] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        dst[lo_off] = make_uint4(lw[0], lw[1], lw[2], lw[3]);        dst[lo_off + rows] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
            }
    }

    // ---- max|v6| of this column for the next bound: tile maximum, the last tile folds ---------------------------------
    s_max[tid] = vmax;
    __syncthreads();
    for (int half = SUBS >> 1; half > 0; half >>= 1) {
        if (sub < half) s_max[tid] = fmaxf(s_max[tid], s_max[tid + half*LN]);
        __syncthreads();
    }
    if (0 == sub) a.partmax[size_t(blockIdx.x)*64 + j] = s_max[j];
    uint32_t const t0 = a.coltile[c], t1 = a.coltile[c + 1];
    __threadfence();
    __syncthreads();
    if (0 == tid) s_last = (atomicAdd(&a.ticket[c], 1u) == (t1 - t0) - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (0 == tid) a.ticket[c] = 0;
    if (tid < LN) {
        float m = 0.f;
        for (uint32_t tt = t0; tt < t1; ++tt) m = fmaxf(m, __ldcg(&a.partmax[size_t(tt)*64 + tid]));
        a.mx6[size_t(c)*LN + tid] = m;
    }
}

// B := unit blocks: b[Re][j % LM][j] = 1, everything else 0 (set_unit_blocks, linalg.hxx:432-455: the right-hand sides of the
// reference's `rhs_trivial` mode, core.hxx:140-147, "columns of the unit matrix")
template <typename real_t>
__global__ void unit_blocks_kernel(real_t *__restrict__ B, int LM, int LN) {
    size_t const base = size_t(blockIdx.x)*2*LM*LN;
    for (int q = threadIdx.x; q < 2*LM*LN; q += blockDim.x) {
        int const c = q/(LM*LN), i = (q/LN) % LM, j = q % LN;
        B[base + q] = real_t((0 == c && i == j % LM) ? 1 : 0);
    }
}

// v[bpos[b]] += scal*B[b]  (add_RHS, linalg.hxx:383-404)
template <typename real_t>
__global__ void add_rhs_kernel(real_t *__restrict__ v, real_t const *__restrict__ B, real_t scal,
                               uint32_t const *__restrict__ bpos, int blockElems, Control const *ctl, int expect) {
    if (expect >= 0 && ctl->state != expect) return;
    size_t const src = size_t(blockIdx.x)*blockElems, dst = size_t(bpos[blockIdx.x])*blockElems;
    for (int q = threadIdx.x; q < blockElems; q += blockDim.x) v[dst + q] += scal*B[src + q];
}

template <typename real_t, int VEC, int OP>
void launch_one(Plan const &p, cudaStream_t stream) {
    int const LNV = p.LN/VEC;
    int const threads = LNV*(256/LNV);
    size_t const smem = vec_smem_bytes<OP>(threads, LNV, p.LN);
    constexpr bool kWritesV45 = (OP == OP_INIT || OP == OP_E1 || OP == OP_K2 || OP == OP_E2);
    VecArgs<real_t> args = make_args<real_t>(p);
    if ((OP == OP_K4 || OP == OP_N3) && p.exch.slots) args.slot_out = exchange_slots(p, (OP == OP_K4) ? 0 : 1) + 4*p.exch.shard;
    if (std::is_same<real_t, float>::value && kWritesV45 && p.use_tc16)
        vec_kernel<real_t, VEC, OP, std::is_same<real_t, float>::value && kWritesV45><<<p.nTiles, threads, smem, stream>>>(args);
    else
        vec_kernel<real_t, VEC, OP, false><<<p.nTiles, threads, smem, stream>>>(args);
}

template <typename real_t, int VEC>
void launch_op(Plan const &p, int op, cudaStream_t stream) {
    switch (op) {
        case OP_INIT: launch_one<real_t, VEC, OP_INIT>(p, stream); break;
        case OP_K1:   launch_one<real_t, VEC, OP_K1>(p, stream); break;
        case OP_E1:   launch_one<real_t, VEC, OP_E1>(p, stream); break;
        case OP_K2:   launch_one<real_t, VEC, OP_K2>(p, stream); break;
        case OP_K3:   launch_one<real_t, VEC, OP_K3>(p, stream); break;
        case OP_E2:   launch_one<real_t, VEC, OP_E2>(p, stream); break;
        case OP_K4:   launch_one<real_t, VEC, OP_K4>(p, stream); break;
        case OP_N3:   launch_one<real_t, VEC, OP_N3>(p, stream); break;
        default: break;
    }
}

} // namespace

// every shard's decision from all shards' exported parts (kind 0: after K4, kind 1: after N3)
tfqmrgpuStatus_t launch_decide(Plan const &p, int kind, cudaStream_t stream)
{
    if (nullptr == p.exch.slots) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    decide_kernel<<<1, 1, 0, stream>>>(ws<Control>(p, p.off_ctl), exchange_slots(p, kind), p.exch.nshards, p.exch.nrhs_global, kind);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

// K1 / K3 that also emit the tensor-core product's X operand from the v6 they write (plans with use_tc16 only)
tfqmrgpuStatus_t launch_vecop_xop(Plan const &p, int op, cudaStream_t stream)
{
    if (!p.use_tc16 || (OP_K1 != op && OP_K3 != op)) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    VecArgs<float> const a = make_args<float>(p);
#define TFQ_VX(LM, LN) case LM*1000 + LN: \
        if (p.tc_planar) { if (OP_K1 == op) vec_xop_kernel<OP_K1, LM, LN, true><<<p.nTiles, 256, 0, stream>>>(a); \
                           else             vec_xop_kernel<OP_K3, LM, LN, true><<<p.nTiles, 256, 0, stream>>>(a); } \
        else             { if (OP_K1 == op) vec_xop_kernel<OP_K1, LM, LN, false><<<p.nTiles, 256, 0, stream>>>(a); \
                           else             vec_xop_kernel<OP_K3, LM, LN, false><<<p.nTiles, 256, 0, stream>>>(a); } \
        break;
    switch (p.LM*1000 + p.LN) {
        TFQ_VX(16, 16) TFQ_VX(16, 32) TFQ_VX(16, 64) TFQ_VX(32, 32) TFQ_VX(32, 64) TFQ_VX(64, 64)
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
#undef TFQ_VX
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t launch_vecop(Plan const &p, int op, cudaStream_t stream)
{
    int const LN = p.LN;
    if ('z' == p.precision) {
        if (LN % 2 == 0) launch_op<double, 2>(p, op, stream); else launch_op<double, 1>(p, op, stream);
    } else if ('c' == p.precision) {
        if (LN % 4 == 0) launch_op<float, 4>(p, op, stream);
        else if (LN % 2 == 0) launch_op<float, 2>(p, op, stream);
        else launch_op<float, 1>(p, op, stream);
    } else {
        return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, p.precision);
    }
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t launch_unit_rhs(Plan const &p, cudaStream_t stream)
{
    if (p.nnzbB < 1) return TFQMRGPU_STATUS_SUCCESS;
    int const threads = std::min(256, ((2*p.LM*p.LN + 31)/32)*32);
    if ('z' == p.precision) unit_blocks_kernel<double><<<p.nnzbB, threads, 0, stream>>>(ws<double>(p, p.off_B), p.LM, p.LN);
    else                    unit_blocks_kernel<float ><<<p.nnzbB, threads, 0, stream>>>(ws<float >(p, p.off_B), p.LM, p.LN);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t launch_add_rhs(Plan const &p, void *v, double scal, int expect, cudaStream_t stream)
{
    if (p.nnzbB < 1) return TFQMRGPU_STATUS_SUCCESS;
    int const blockElems = 2*p.LM*p.LN;
    int const threads = std::min(256, ((blockElems + 31)/32)*32);
    if ('z' == p.precision)
        add_rhs_kernel<double><<<p.nnzbB, threads, 0, stream>>>(static_cast<double*>(v), ws<double const>(p, p.off_B), scal,
                                                                p.d_bpos, blockElems, ws<Control const>(p, p.off_ctl), expect);
    else
        add_rhs_kernel<float><<<p.nnzbB, threads, 0, stream>>>(static_cast<float*>(v), ws<float const>(p, p.off_B), float(scal),
                                                               p.d_bpos, blockElems, ws<Control const>(p, p.off_ctl), expect);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace tfq
