// Host block layout  <->  internal block layout.
//
// Role of the reference's set_or_getMatrix + transpose_blocks_kernel (tfqmrgpu.cu:467-603,
// tfqmrgpu_linalg.hxx:282-380): a host block holds rows x cols complex numbers as RIRIRIRI
// [rows][cols][2], RRIIRRII [rows][2][cols] or RRRRIIII [2][rows][cols], optionally transposed
// ([cols][rows]) and/or conjugated; internally every block is RRRRIIII real_t[2][rows][cols].
// One CTA stages a whole block in shared memory (<= 64 KiB for 64x64 complex double), so both global
// accesses are linear and the conversion may run in place (A and B).  X-shaped data is additionally
// permuted between the caller's block order and the column-sorted storage order (out of place).
// Deviation from the reference (documented in DESIGN.md): a transposed RECTANGULAR block is read with
// the mathematically correct leading dimension; the reference indexes out of the block there.
#include "tfq_internal.hpp"

namespace tfq {

namespace {

struct HostIndex { // strides of the logical element (i, j, c) inside a host block
    uint32_t Ni, Nj, Nc;
};

HostIndex host_index(int layout, int rows, int cols, bool trans) {
    uint32_t const fast = trans ? rows : cols;  // length of the contiguous host dimension
    uint32_t slow_s = 0, fast_s = 0, cplx_s = 0;
    switch (layout) {
        case 0x0f: cplx_s = uint32_t(rows)*cols; slow_s = fast; fast_s = 1; break; // RRRRIIII
        case 0x33: slow_s = 2*fast; cplx_s = fast; fast_s = 1; break;              // RRIIRRII
        default:   slow_s = 2*fast; fast_s = 2; cplx_s = 1; break;                 // RIRIRIRI
    }
    HostIndex h;
    if (trans) { h.Ni = fast_s; h.Nj = slow_s; } else { h.Ni = slow_s; h.Nj = fast_s; }
    h.Nc = cplx_s;
    return h;
}

// to_internal: dst[map(b)][c][i][j] = scal(c) * src[b][host(i,j,c)]
// otherwise  : dst[b][host(i,j,c)]  = scal(c) * src[map(b)][c][i][j]
// map = perm (caller index -> storage index) or identity when perm == nullptr
template <typename real_t>
__global__ void convert_kernel(real_t *dst, real_t const *src, uint32_t const *__restrict__ perm,
                               int rows, int cols, HostIndex h, real_t scal_imag, bool to_internal)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real_t *const tmp = reinterpret_cast<real_t*>(smem_raw);
    uint32_t const b = blockIdx.x;
    uint32_t const bs = 2u*rows*cols, plane = uint32_t(rows)*cols;
    size_t const mapped = perm ? size_t(perm[b]) : size_t(b);
    real_t const *const s = src + (to_internal ? size_t(b) : mapped)*bs;
    real_t *const d = dst + (to_internal ? mapped : size_t(b))*bs;
    for (uint32_t q = threadIdx.x; q < bs; q += blockDim.x) tmp[q] = s[q]; // linear read
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < bs; q += blockDim.x) {
        // q runs over the DESTINATION linearly; find the matching source element in shared memory
        if (to_internal) {
            uint32_t const c = q / plane, ij = q - c*plane, i = ij / cols, j = ij - i*cols;
            real_t const v = tmp[h.Ni*i + h.Nj*j + h.Nc*c];
            d[q] = c ? scal_imag*v : v;
        } else {
            // decode the host position q -> (i, j, c): invert h by trying the three stride orders
            uint32_t c, i, j;
            if (1 == h.Nc) { c = q & 1u; uint32_t const r = q >> 1; // RIRIRIRI: q = 2*(slow*F + fast) + c
                if (2 == h.Nj) { i = r / cols; j = r - i*cols; } else { j = r / rows; i = r - j*rows; }
            } else if (plane == h.Nc) { c = q / plane; uint32_t const r = q - c*plane; // RRRRIIII
                if (1 == h.Nj) { i = r / cols; j = r - i*cols; } else { j = r / rows; i = r - j*rows; }
            } else { // RRIIRRII: q = (slow*2 + c)*F + fast
                uint32_t const F = h.Nc, sl = q / (2*F), r = q - sl*2*F; c = r / F; uint32_t const f = r - c*F;
                if (1 == h.Nj) { i = sl; j = f; } else { j = sl; i = f; }
            }
            real_t const v = tmp[(c*rows + i)*cols + j];
            d[q] = c ? scal_imag*v : v;
        }
    }
}

__global__ void permute_blocks_f32(float *dst, float const *src, uint32_t const *__restrict__ perm, uint32_t blockElems, bool to_storage) {
    uint32_t const b = blockIdx.x;
    size_t const so = size_t(to_storage ? b : perm[b])*blockElems, dof = size_t(to_storage ? perm[b] : b)*blockElems;
    for (uint32_t q = threadIdx.x; q < blockElems; q += blockDim.x) dst[dof + q] = src[so + q];
}

template <typename real_t>
tfqmrgpuStatus_t run_convert(void *dst, void const *src, uint32_t const *perm, uint32_t nnzb, int rows, int cols,
                             int layout, bool trans, double scal_imag, bool to_internal, cudaStream_t stream)
{
    if (nnzb < 1) return TFQMRGPU_STATUS_SUCCESS;
    size_t const smem = 2*size_t(rows)*cols*sizeof(real_t);
    auto kernel = convert_kernel<real_t>;
    if (smem > 48*1024) TFQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int const threads = int(std::min<size_t>(256, ((2*size_t(rows)*cols + 31)/32)*32));
    kernel<<<nnzb, threads, smem, stream>>>(static_cast<real_t*>(dst), static_cast<real_t const*>(src), perm, rows, cols,
                                           host_index(layout, rows, cols, trans), real_t(scal_imag), to_internal);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace

tfqmrgpuStatus_t convert_inplace(Plan const &, void *blocks, uint32_t nnzb, int rows, int cols, bool is_double,
                                 int layout, bool trans, double scal_imag, cudaStream_t stream)
{
    return is_double ? run_convert<double>(blocks, blocks, nullptr, nnzb, rows, cols, layout, trans, scal_imag, true, stream)
                     : run_convert<float >(blocks, blocks, nullptr, nnzb, rows, cols, layout, trans, scal_imag, true, stream);
}

tfqmrgpuStatus_t convert_permuted(Plan const &p, void *dst, void const *src, uint32_t nnzb, int rows, int cols,
                                  bool is_double, int layout, bool trans, double scal_imag, bool to_internal, cudaStream_t stream)
{
    return is_double ? run_convert<double>(dst, src, p.d_perm, nnzb, rows, cols, layout, trans, scal_imag, to_internal, stream)
                     : run_convert<float >(dst, src, p.d_perm, nnzb, rows, cols, layout, trans, scal_imag, to_internal, stream);
}

tfqmrgpuStatus_t permute_v3(Plan const &p, float *dst, float const *src, bool to_storage, cudaStream_t stream)
{
    uint32_t const blockElems = 2u*p.LM*p.LN;
    permute_blocks_f32<<<p.nnzbX, std::min(256u, ((blockElems + 31)/32)*32), 0, stream>>>(dst, src, p.d_perm, blockElems, to_storage);
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace tfq
