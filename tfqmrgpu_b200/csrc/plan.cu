// Plan analysis on the device.
//
// Replaces the single-threaded host analysis of the reference's createPlan (tfqmrgpu.cu:183-337):
// the multiplication pair list (pairs/starts, one search per (Y block, A block)), the B-subset list,
// the dense renumbering of X's block columns and the structural checks are built by CUDA kernels from
// the caller's BSR index arrays.  The four reference-format lists are kept bit-identical to the
// reference (tests compare them with the oracle and with the reference itself); on top of them this
// file derives the B200-specific structures: the column-sorted storage permutation of X-shaped
// vectors, the vector tiles and the row-grouped SpMM units.
#include "tfq_internal.hpp"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <functional>

namespace tfq {

namespace {

template <typename T> struct DevArray {
    T *ptr = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&ptr, std::max<size_t>(n, 1)*sizeof(T)); }
    ~DevArray() { if (ptr) cudaFree(ptr); }
    T* release() { T *q = ptr; ptr = nullptr; return q; }
};

// first index in [begin, end) with array[index] == value, -1 if absent (bsr.hxx:27-39 semantics)
__device__ __forceinline__ int find_first(int32_t const *__restrict__ array, int begin, int end, int value, bool sorted) {
    if (sorted) { // strictly ascending row: lower bound
        int lo = begin, hi = end;
        while (lo < hi) { int const mid = (lo + hi) >> 1; if (array[mid] < value) lo = mid + 1; else hi = mid; }
        return (lo < end && array[lo] == value) ? lo : -1;
    }
    for (int i = begin; i < end; ++i) if (array[i] == value) return i;
    return -1;
}

__global__ void k_row_of(int32_t const *__restrict__ rp, int mb, int off, uint32_t *__restrict__ rowOf) {
    int const r = blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= mb) return;
    for (int i = rp[r] - off; i < rp[r + 1] - off; ++i) rowOf[i] = r;
}

__global__ void k_rows_sorted(int32_t const *__restrict__ rp, int32_t const *__restrict__ ci, int mb, int off, int *unsorted) {
    int const r = blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= mb) return;
    for (int i = rp[r] - off + 1; i < rp[r + 1] - off; ++i) if (ci[i - 1] >= ci[i]) { *unsorted = 1; return; }
}

// tfqmrgpu.cu:198-219 : one thread per Y block, A's given column order is preserved
template <bool Fill>
__global__ void k_pairs(int nnzbY, uint32_t const *__restrict__ rowOfX, int32_t const *__restrict__ rpA,
                        int32_t const *__restrict__ ciA, int32_t const *__restrict__ rpX,
                        int32_t const *__restrict__ ciX, int off, bool sorted,
                        uint32_t *__restrict__ count, uint32_t const *__restrict__ starts, uint32_t *__restrict__ pairs) {
    int const iY = blockIdx.x*blockDim.x + threadIdx.x;
    if (iY >= nnzbY) return;
    int const r = rowOfX[iY];
    int const jcol = ciX[iY];
    uint32_t n = 0;
    size_t const base = Fill ? starts[iY] : 0;
    for (int inza = rpA[r] - off; inza < rpA[r + 1] - off; ++inza) {
        int const k = ciA[inza] - off;
        int const inzx = find_first(ciX, rpX[k] - off, rpX[k + 1] - off, jcol, sorted);
        if (inzx >= 0) {
            if (Fill) { pairs[2*(base + n)] = uint32_t(inza); pairs[2*(base + n) + 1] = uint32_t(inzx); }
            ++n;
        }
    }
    if (!Fill) count[iY] = n;
}

// tfqmrgpu.cu:236-249
__global__ void k_subset(int nnzbB, uint32_t const *__restrict__ rowOfB, int32_t const *__restrict__ ciB,
                         int32_t const *__restrict__ rpX, int32_t const *__restrict__ ciX, int off, bool sorted,
                         uint32_t *__restrict__ subset, int *__restrict__ err_row) {
    int const ib = blockIdx.x*blockDim.x + threadIdx.x;
    if (ib >= nnzbB) return;
    int const r = rowOfB[ib];
    int const inzx = find_first(ciX, rpX[r] - off, rpX[r + 1] - off, ciB[ib], sorted);
    if (inzx < 0) { atomicMin(err_row, r); subset[ib] = 0; } else subset[ib] = uint32_t(inzx);
}

__global__ void k_mark_used(int n, int32_t const *__restrict__ ciX, int min_c, uint32_t *__restrict__ used) {
    int const i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) used[ciX[i] - min_c] = 1;
}
__global__ void k_colindx(int n, int32_t const *__restrict__ ciX, int min_c, uint32_t const *__restrict__ translate,
                          uint16_t *__restrict__ colindx, uint32_t *__restrict__ iota, uint32_t *__restrict__ xcount) {
    int const i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t const jb = translate[ciX[i] - min_c];
    colindx[i] = uint16_t(jb);
    iota[i] = i;
    atomicAdd(&xcount[jb], 1u);
}
__global__ void k_bcount(int nnzbB, uint32_t const *__restrict__ subset, uint16_t const *__restrict__ colindx,
                         uint32_t *__restrict__ bcount) {
    int const ib = blockIdx.x*blockDim.x + threadIdx.x;
    if (ib < nnzbB) atomicAdd(&bcount[colindx[subset[ib]]], 1u);
}
__global__ void k_count_zero(int nb, uint32_t const *__restrict__ bcount, uint32_t *__restrict__ nzero) {
    int const jb = blockIdx.x*blockDim.x + threadIdx.x;
    if (jb < nb && 0 == bcount[jb]) atomicAdd(nzero, 1u);
}
__global__ void k_invert(int n, uint32_t const *__restrict__ iperm, uint32_t *__restrict__ perm) {
    int const s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s < n) perm[iperm[s]] = s;
}
__global__ void k_bpos(int nnzbB, uint32_t const *__restrict__ subset, uint32_t const *__restrict__ perm, uint32_t *__restrict__ bpos) {
    int const ib = blockIdx.x*blockDim.x + threadIdx.x;
    if (ib < nnzbB) bpos[ib] = perm[subset[ib]];
}

__global__ void k_blockcol(int n, uint32_t const *__restrict__ iperm, uint16_t const *__restrict__ colindx, uint32_t *__restrict__ blockcol) {
    int const s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s < n) blockcol[s] = colindx[iperm[s]];
}

// 64 x 64 blocks on the 32 x 32 tensor-core kernel (spmm_tc16.cu): every unit (one Y block) becomes two units (row halves
// ih) whose two block columns are the column halves jh of that Y block; every entry becomes two (k halves kh).
// Sub-block numbering: A (ia, ih, kh), X (ix, kh, jh), Y (iy, ih).
__global__ void k_expand_v64(uint32_t nUnits, uint32_t const *__restrict__ e0, uint32_t const *__restrict__ unit_y,
                             uint32_t const *__restrict__ unit_row, uint32_t const *__restrict__ ent_a, uint32_t const *__restrict__ ent_x,
                             uint32_t *__restrict__ e0v, uint32_t *__restrict__ unit_yv, uint32_t *__restrict__ unit_rowv,
                             uint32_t *__restrict__ ent_av, uint32_t *__restrict__ ent_xv) {
    uint32_t const u = blockIdx.x*blockDim.x + threadIdx.x;
    if (u > nUnits) return;
    if (u == nUnits) { e0v[2*u] = 4*e0[u]; return; }
    uint32_t const b = e0[u], n = e0[u + 1] - b, iy = unit_y[u];
    for (uint32_t ih = 0; ih < 2; ++ih) {
        uint32_t const uv = 2*u + ih, bv = 4*b + ih*2*n;
        e0v[uv] = bv;
        unit_rowv[uv] = unit_row[u];
        unit_yv[2*uv] = unit_yv[2*uv + 1] = (kNoBlock == iy) ? kNoBlock : (2*iy + ih);
        for (uint32_t t = 0; t < n; ++t) {
            uint32_t const ia = ent_a[b + t], ix = ent_x[b + t];
            for (uint32_t kh = 0; kh < 2; ++kh) {
                uint32_t const ev = bv + 2*t + kh;
                ent_av[ev] = (2*ia + ih)*2 + kh;
                ent_xv[2*ev] = (kNoBlock == ix) ? kNoBlock : ((2*ix + kh)*2 + 0);
                ent_xv[2*ev + 1] = (kNoBlock == ix) ? kNoBlock : ((2*ix + kh)*2 + 1);
            }
        }
    }
}

// SpMM units: merge the (ascending) pair lists of the unit's Y blocks into one entry list
template <bool Fill>
__global__ void k_units(uint32_t nUnits, uint32_t gmax, uint32_t const *__restrict__ unit_first, uint32_t const *__restrict__ unit_ng,
                        uint32_t const *__restrict__ starts, uint32_t const *__restrict__ pairs, uint32_t const *__restrict__ perm,
                        uint32_t *__restrict__ count, uint32_t const *__restrict__ e0, uint32_t *__restrict__ unit_y,
                        uint32_t *__restrict__ ent_a, uint32_t *__restrict__ ent_x) {
    uint32_t const u = blockIdx.x*blockDim.x + threadIdx.x;
    if (u >= nUnits) return;
    uint32_t const y0 = unit_first[u], ng = unit_ng[u];
    uint32_t head[16], stop[16];
    for (uint32_t g = 0; g < 16; ++g) {
        head[g] = (g < ng) ? starts[y0 + g] : 0;
        stop[g] = (g < ng) ? starts[y0 + g + 1] : 0;
    }
    if (Fill) for (uint32_t g = 0; g < gmax; ++g) unit_y[size_t(u)*gmax + g] = (g < ng) ? perm[y0 + g] : kNoBlock;
    size_t e = Fill ? e0[u] : 0;
    uint32_t n = 0;
    for (;;) {
        uint32_t amin = 0xffffffffu;
        for (uint32_t g = 0; g < 16; ++g) if (head[g] < stop[g]) amin = min(amin, pairs[2*size_t(head[g])]);
        if (0xffffffffu == amin) break;
        if (Fill) ent_a[e + n] = amin;
        for (uint32_t g = 0; g < 16; ++g) {
            bool const hit = (head[g] < stop[g]) && (pairs[2*size_t(head[g])] == amin);
            if (Fill && g < gmax) ent_x[(e + n)*gmax + g] = hit ? perm[pairs[2*size_t(head[g]) + 1]] : kNoBlock;
            if (hit) ++head[g];
        }
        ++n;
    }
    if (!Fill) count[u] = n;
}

inline unsigned nblk(size_t n, unsigned t = 256) { return unsigned((n + t - 1)/t); }

template <typename T>
cudaError_t exclusive_scan(T *d_out, T const *d_in, size_t n, cudaStream_t stream) {
    void *tmp = nullptr; size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, d_in, d_out, int(n), stream);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tmp, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, bytes, d_in, d_out, int(n), stream);
    cudaError_t e2 = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    return (e != cudaSuccess) ? e : e2;
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

} // namespace

// ------------------------------------------------------------------------------------------------
tfqmrgpuStatus_t plan_analyse(Plan &p, cudaStream_t stream,
    int32_t const *rpA, int32_t const *ciA, int32_t const *rpX, int32_t const *ciX,
    int32_t const *rpB, int32_t const *ciB, int echo)
{
    int const mb = p.mb, nnzbA = p.nnzbA, nnzbX = p.nnzbX, nnzbB = p.nnzbB, off = p.indexOffset;

    // host-side validation of what the kernels index with (the reference trusts its input here)
    auto rows_ok = [&](int32_t const *rp, int nnz) {
        if (rp[0] != off) return false;
        for (int r = 0; r < mb; ++r) if (rp[r + 1] < rp[r]) return false;
        return rp[mb] - off == nnz;
    };
    if (!rows_ok(rpA, nnzbA) || !rows_ok(rpX, nnzbX) || !rows_ok(rpB, nnzbB)) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    for (int i = 0; i < nnzbA; ++i) if (ciA[i] - off < 0 || ciA[i] - off >= mb) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);

    // block-column range of X (tfqmrgpu.cu:257-264)
    int32_t min_c = 2147483647, max_c = -2147483647;
    for (int i = 0; i < nnzbX; ++i) { min_c = std::min(min_c, ciX[i]); max_c = std::max(max_c, ciX[i]); }
    long long const nc = 1LL + max_c - min_c;
    if (nc < 1) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nc > (1LL << 28)) return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED);
    if (echo > 5) std::printf("# tfqmrgpu_bsrsv_createPlan: column indices of X are in [%d, %d]\n", min_c, max_c);

    p.h_rowptrX.resize(mb + 1);
    p.maxColsPerRow = 1;
    for (int r = 0; r <= mb; ++r) p.h_rowptrX[r] = rpX[r] - off;
    p.h_rpA.resize(mb + 1); p.h_rpB.resize(mb + 1);
    for (int r = 0; r <= mb; ++r) { p.h_rpA[r] = rpA[r] - off; p.h_rpB[r] = rpB[r] - off; }
    p.h_ciA.resize(nnzbA); for (int i = 0; i < nnzbA; ++i) p.h_ciA[i] = ciA[i] - off;
    p.h_ciX.assign(ciX, ciX + nnzbX);                 // X and B column values are labels, compared unshifted (tfqmrgpu.cu:238-241)
    p.h_ciB.assign(ciB, ciB + nnzbB);
    for (int r = 0; r < mb; ++r) p.maxColsPerRow = std::max(p.maxColsPerRow, p.h_rowptrX[r + 1] - p.h_rowptrX[r]);

    // ---- upload the six index arrays --------------------------------------------------------------
    DevArray<int32_t> d_rpA, d_ciA, d_rpX, d_ciX, d_rpB, d_ciB;
    TFQ_CUDA(d_rpA.alloc(mb + 1)); TFQ_CUDA(d_ciA.alloc(nnzbA));
    TFQ_CUDA(d_rpX.alloc(mb + 1)); TFQ_CUDA(d_ciX.alloc(nnzbX));
    TFQ_CUDA(d_rpB.alloc(mb + 1)); TFQ_CUDA(d_ciB.alloc(nnzbB));
    auto up = [&](int32_t *d, int32_t const *h, size_t n) {
        return n ? cudaMemcpyAsync(d, h, n*sizeof(int32_t), cudaMemcpyHostToDevice, stream) : cudaSuccess; };
    TFQ_CUDA(up(d_rpA.ptr, rpA, mb + 1)); TFQ_CUDA(up(d_ciA.ptr, ciA, nnzbA));
    TFQ_CUDA(up(d_rpX.ptr, rpX, mb + 1)); TFQ_CUDA(up(d_ciX.ptr, ciX, nnzbX));
    TFQ_CUDA(up(d_rpB.ptr, rpB, mb + 1)); TFQ_CUDA(up(d_ciB.ptr, ciB, nnzbB));

    DevArray<uint32_t> d_rowOfX, d_rowOfB, d_cnt, d_flags;
    TFQ_CUDA(d_rowOfX.alloc(nnzbX)); TFQ_CUDA(d_rowOfB.alloc(nnzbB));
    TFQ_CUDA(d_cnt.alloc(size_t(nnzbX) + 1));
    TFQ_CUDA(d_flags.alloc(4)); // [0] unsorted, [1] err_row, [2] nzero
    int const h_flags_init[4] = {0, 2147483647, 0, 0};
    TFQ_CUDA(cudaMemcpyAsync(d_flags.ptr, h_flags_init, sizeof(h_flags_init), cudaMemcpyHostToDevice, stream));

    k_row_of<<<nblk(mb), 256, 0, stream>>>(d_rpX.ptr, mb, off, d_rowOfX.ptr);
    if (nnzbB > 0) k_row_of<<<nblk(mb), 256, 0, stream>>>(d_rpB.ptr, mb, off, d_rowOfB.ptr);
    k_rows_sorted<<<nblk(mb), 256, 0, stream>>>(d_rpX.ptr, d_ciX.ptr, mb, off, (int*)d_flags.ptr);
    int h_flags[4];
    TFQ_CUDA(cudaMemcpyAsync(h_flags, d_flags.ptr, sizeof(h_flags), cudaMemcpyDeviceToHost, stream));
    TFQ_CUDA(cudaStreamSynchronize(stream));
    bool const sorted = (0 == h_flags[0]);

    // ---- pairs / starts ---------------------------------------------------------------------------
    TFQ_CUDA(cudaMalloc((void**)&p.d_starts, (size_t(nnzbX) + 1)*sizeof(uint32_t)));
    TFQ_CUDA(cudaMemsetAsync(d_cnt.ptr, 0, (size_t(nnzbX) + 1)*sizeof(uint32_t), stream));
    k_pairs<false><<<nblk(nnzbX, 128), 128, 0, stream>>>(nnzbX, d_rowOfX.ptr, d_rpA.ptr, d_ciA.ptr, d_rpX.ptr, d_ciX.ptr,
                                                         off, sorted, d_cnt.ptr, nullptr, nullptr);
    TFQ_CUDA(exclusive_scan(p.d_starts, d_cnt.ptr, size_t(nnzbX) + 1, stream));
    uint32_t npairs = 0;
    TFQ_CUDA(cudaMemcpy(&npairs, p.d_starts + nnzbX, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    p.nPairs = npairs;
    TFQ_CUDA(cudaMalloc((void**)&p.d_pairs, std::max<size_t>(2*size_t(npairs), 1)*sizeof(uint32_t)));
    k_pairs<true><<<nblk(nnzbX, 128), 128, 0, stream>>>(nnzbX, d_rowOfX.ptr, d_rpA.ptr, d_ciA.ptr, d_rpX.ptr, d_ciX.ptr,
                                                        off, sorted, nullptr, p.d_starts, p.d_pairs);
    if (echo > 6) std::printf("# tfqmrgpu_bsrsv_createPlan: found %u pairs in A*X multiplication\n", npairs);

    // ---- subset (B must be a subset of X) ----------------------------------------------------------
    TFQ_CUDA(cudaMalloc((void**)&p.d_subset, std::max(nnzbB, 1)*sizeof(uint32_t)));
    if (nnzbB > 0) k_subset<<<nblk(nnzbB), 256, 0, stream>>>(nnzbB, d_rowOfB.ptr, d_ciB.ptr, d_rpX.ptr, d_ciX.ptr, off, sorted,
                                                              p.d_subset, (int*)d_flags.ptr + 1);
    TFQ_CUDA(cudaMemcpyAsync(h_flags, d_flags.ptr, sizeof(h_flags), cudaMemcpyDeviceToHost, stream));
    TFQ_CUDA(cudaStreamSynchronize(stream));
    if (h_flags[1] != 2147483647) {
        if (echo > 0) std::printf("# tfqmrgpu_bsrsv_createPlan: in row #%d B has a block that X does not have!\n", h_flags[1] + off);
        return TFQMRGPU_B_IS_NOT_SUBSET_OF_X + TFQMRGPU_CODE_LINE*h_flags[1]; // tfqmrgpu.cu:245
    }

    // ---- dense renumbering of the used block columns (tfqmrgpu.cu:268-314) --------------------------
    DevArray<uint32_t> d_used, d_translate, d_iota, d_xcount;
    TFQ_CUDA(d_used.alloc(size_t(nc) + 1)); TFQ_CUDA(d_translate.alloc(size_t(nc) + 1));
    TFQ_CUDA(cudaMemsetAsync(d_used.ptr, 0, (size_t(nc) + 1)*sizeof(uint32_t), stream));
    k_mark_used<<<nblk(nnzbX), 256, 0, stream>>>(nnzbX, d_ciX.ptr, min_c, d_used.ptr);
    TFQ_CUDA(exclusive_scan(d_translate.ptr, d_used.ptr, size_t(nc) + 1, stream));
    uint32_t nb = 0;
    TFQ_CUDA(cudaMemcpy(&nb, d_translate.ptr + nc, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (nb < 1) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nb > 65536u) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR); // colIndex_t is uint16 (core.hxx:81)
    p.nCols = nb;
    if (echo > 5) std::printf("# tfqmrgpu_bsrsv_createPlan: found %lld empty columns and %u columns with entries\n", nc - nb, nb);

    TFQ_CUDA(cudaMalloc((void**)&p.d_colindx, size_t(nnzbX)*sizeof(uint16_t)));
    TFQ_CUDA(d_iota.alloc(nnzbX)); TFQ_CUDA(d_xcount.alloc(size_t(nb) + 1));
    TFQ_CUDA(cudaMemsetAsync(d_xcount.ptr, 0, (size_t(nb) + 1)*sizeof(uint32_t), stream));
    k_colindx<<<nblk(nnzbX), 256, 0, stream>>>(nnzbX, d_ciX.ptr, min_c, d_translate.ptr, p.d_colindx, d_iota.ptr, d_xcount.ptr);

    // ---- every block column of X needs a B block (tfqmrgpu.cu:316-337) ------------------------------
    {
        DevArray<uint32_t> d_bcount;
        TFQ_CUDA(d_bcount.alloc(nb));
        TFQ_CUDA(cudaMemsetAsync(d_bcount.ptr, 0, size_t(nb)*sizeof(uint32_t), stream));
        if (nnzbB > 0) k_bcount<<<nblk(nnzbB), 256, 0, stream>>>(nnzbB, p.d_subset, p.d_colindx, d_bcount.ptr);
        k_count_zero<<<nblk(nb), 256, 0, stream>>>(int(nb), d_bcount.ptr, d_flags.ptr + 2);
        TFQ_CUDA(cudaMemcpyAsync(h_flags, d_flags.ptr, sizeof(h_flags), cudaMemcpyDeviceToHost, stream));
        TFQ_CUDA(cudaStreamSynchronize(stream));
        if (h_flags[2] > 0) {
            if (echo > 0) std::printf("# tfqmrgpu_bsrsv_createPlan: found %d zero columns in B!\n", h_flags[2]);
            return TFQMRGPU_B_HAS_A_ZERO_COLUMN + TFQMRGPU_CODE_LINE*h_flags[2];
        }
    }

    // ---- column-sorted storage order: stable sort of the blocks by block column ----------------------
    TFQ_CUDA(cudaMalloc((void**)&p.d_perm,  size_t(nnzbX)*sizeof(uint32_t)));
    TFQ_CUDA(cudaMalloc((void**)&p.d_iperm, size_t(nnzbX)*sizeof(uint32_t)));
    {
        DevArray<uint16_t> d_keys_out; TFQ_CUDA(d_keys_out.alloc(nnzbX));
        void *tmp = nullptr; size_t bytes = 0;
        TFQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, p.d_colindx, d_keys_out.ptr, d_iota.ptr, p.d_iperm, nnzbX, 0, 16, stream));
        TFQ_CUDA(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
        cudaError_t const e = cub::DeviceRadixSort::SortPairs(tmp, bytes, p.d_colindx, d_keys_out.ptr, d_iota.ptr, p.d_iperm, nnzbX, 0, 16, stream);
        cudaError_t const e2 = cudaStreamSynchronize(stream);
        cudaFree(tmp);
        TFQ_CUDA(e); TFQ_CUDA(e2);
    }
    k_invert<<<nblk(nnzbX), 256, 0, stream>>>(nnzbX, p.d_iperm, p.d_perm);
    {
        DevArray<uint32_t> d_colstart; TFQ_CUDA(d_colstart.alloc(size_t(nb) + 1));
        TFQ_CUDA(exclusive_scan(d_colstart.ptr, d_xcount.ptr, size_t(nb) + 1, stream));
        p.h_colstart.resize(size_t(nb) + 1);
        TFQ_CUDA(cudaMemcpy(p.h_colstart.data(), d_colstart.ptr, (size_t(nb) + 1)*sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    TFQ_CUDA(cudaMalloc((void**)&p.d_bpos, std::max(nnzbB, 1)*sizeof(uint32_t)));
    if (nnzbB > 0) k_bpos<<<nblk(nnzbB), 256, 0, stream>>>(nnzbB, p.d_subset, p.d_perm, p.d_bpos);
    TFQ_CUDA(cudaMalloc((void**)&p.d_blockcol, size_t(nnzbX)*sizeof(uint32_t)));
    k_blockcol<<<nblk(nnzbX), 256, 0, stream>>>(nnzbX, p.d_iperm, p.d_colindx, p.d_blockcol);
    {
        std::vector<int32_t> rp0(size_t(mb) + 1);
        for (int r = 0; r <= mb; ++r) rp0[r] = rpA[r] - off;
        TFQ_CUDA(cudaMalloc((void**)&p.d_rowptrA, rp0.size()*sizeof(int32_t)));
        TFQ_CUDA(cudaMemcpyAsync(p.d_rowptrA, rp0.data(), rp0.size()*sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        TFQ_CUDA(cudaStreamSynchronize(stream));     // rp0 is a local
    }
    TFQ_CUDA(cudaStreamSynchronize(stream));
    TFQ_CUDA(cudaGetLastError());
    return TFQMRGPU_STATUS_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
void plan_drop_graph(Plan &p) {
    if (p.body_exec) { cudaGraphExecDestroy(p.body_exec); p.body_exec = nullptr; }
}

static void free_configured(Plan &p) {
    plan_drop_graph(p);
    cudaFree(p.d_guess_scratch); p.d_guess_scratch = nullptr;
    cudaFree(p.d_precond_tmp); p.d_precond_tmp = nullptr;
    cudaFree(p.d_tiles); p.d_tiles = nullptr;
    cudaFree(p.d_coltile); p.d_coltile = nullptr;
    cudaFree(p.d_unit_e0); p.d_unit_e0 = nullptr;
    cudaFree(p.d_unit_y); p.d_unit_y = nullptr;
    cudaFree(p.d_ent_a); p.d_ent_a = nullptr;
    cudaFree(p.d_ent_x); p.d_ent_x = nullptr;
    cudaFree(p.d_cta_u0); p.d_cta_u0 = nullptr;
    cudaFree(p.d_unit_row); p.d_unit_row = nullptr;
    cudaFree(p.d_unit_of_block); p.d_unit_of_block = nullptr;
    p.resident_fits = -1;
    cudaFree(p.d_res_tiles); p.d_res_tiles = nullptr;
    cudaFree(p.d_res_coltile); p.d_res_coltile = nullptr;
    cudaFree(p.d_res_part); p.d_res_part = nullptr;
}

void plan_release(Plan &p) {
    multi_destroy(p);
    mixed_destroy(p);
    free_configured(p);
    cudaFree(p.d_starts); cudaFree(p.d_pairs); cudaFree(p.d_subset); cudaFree(p.d_colindx);
    cudaFree(p.d_perm); cudaFree(p.d_iperm); cudaFree(p.d_bpos); cudaFree(p.d_blockcol); cudaFree(p.d_rowptrA);
    if (p.d_guess_scratch) cudaFree(p.d_guess_scratch);
    if (p.d_resident_bar) cudaFree(p.d_resident_bar);
    if (p.d_resident_trace) cudaFree(p.d_resident_trace);
    if (p.h_ctl) cudaFreeHost(p.h_ctl);
    for (auto &e : p.ev) if (e) cudaEventDestroy(e);
    for (auto &e : p.prof_ev) if (e) cudaEventDestroy(e);
    if (p.capture_stream) cudaStreamDestroy(p.capture_stream);
    if (p.copy_stream) cudaStreamDestroy(p.copy_stream);
    for (auto &e : p.chunk_ev) if (e) cudaEventDestroy(e);
}

// X blocks per vector tile of a block column with nColBlocks blocks.  The rule looks at the COLUMN only (its length, the block
// size, the SM count), never at how many columns the plan has: a column is then cut into the same tiles - and its sums are added in
// the same order - whether it is solved alone, with all other columns, or in a shard of them on another GPU.  (A first version
// sized the tiles by the whole plan and shards had to be told the unsharded plan's tile size: an 8-GPU shard then ran its
// vector kernels on 1/8 of the CTAs, 75 instead of 50 ms per solve at config 3.)
// At least 16 KiB of a vector per tile, 4 KiB when the column is so short that this would leave most SMs without a tile
// (FD_problem.xml, 25 KB per column: 2.98 -> 2.64 ms per solve); long columns: 2 tiles per SM and column.
size_t plan_tile_blocks(size_t nColBlocks, size_t blockBytes, int nsm)
{
    size_t const target = size_t(nsm)*2;
    size_t const tb = std::max<size_t>(1, (nColBlocks + target - 1)/target);
    char const *env_tile = std::getenv("TFQMRGPU_TILE_KB");      // dev switch: minimum bytes of one vector per tile
    size_t const col_bytes = nColBlocks*blockBytes;
    size_t const tile_kb = env_tile ? size_t(std::max(1, std::atoi(env_tile))) : ((col_bytes/(16*1024) < size_t(nsm)) ? 4 : 16);
    size_t const tb_min = std::max<size_t>(1, (tile_kb*1024 + blockBytes - 1)/blockBytes);
    return std::max(tb, tb_min);
}

tfqmrgpuStatus_t plan_configure(Plan &p, cudaStream_t stream, int LM, int LN, char precision)
{
    free_configured(p);
    p.configured = false;
    p.LM = LM; p.LN = LN; p.precision = precision;
    bool const is_double = ('z' == precision);
    size_t const s = is_double ? 8 : 4;
    size_t const blockBytes = 2*size_t(LM)*LN*s;
    uint32_t const nb = p.nCols;

    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);

    // ---- vector tiles: every tile is a contiguous range of blocks of ONE block column ---------------
    {
        p.tile_blocks = 0;
        std::vector<Tile> tiles;
        std::vector<uint32_t> coltile(size_t(nb) + 1, 0);
        for (uint32_t c = 0; c < nb; ++c) {
            uint32_t const b0 = p.h_colstart[c], b1 = p.h_colstart[c + 1];
            uint32_t const n = b1 - b0;
            size_t const tb = (p.tile_blocks_hint > 0) ? p.tile_blocks_hint : plan_tile_blocks(n, blockBytes, nsm);
            p.tile_blocks = std::max(p.tile_blocks, tb);
            uint32_t const nt = std::max<uint32_t>(1, uint32_t((n + tb - 1)/tb));
            coltile[c] = uint32_t(tiles.size());
            for (uint32_t t = 0; t < nt; ++t) {
                Tile tile; tile.col = c; tile.pad = 0;
                tile.b0 = b0 + uint32_t((uint64_t(n)*t)/nt);
                tile.b1 = b0 + uint32_t((uint64_t(n)*(t + 1))/nt);
                tiles.push_back(tile);
            }
        }
        coltile[nb] = uint32_t(tiles.size());
        p.nTiles = uint32_t(tiles.size());
        TFQ_CUDA(cudaMalloc((void**)&p.d_tiles, tiles.size()*sizeof(Tile)));
        TFQ_CUDA(cudaMalloc((void**)&p.d_coltile, coltile.size()*sizeof(uint32_t)));
        TFQ_CUDA(cudaMemcpy(p.d_tiles, tiles.data(), tiles.size()*sizeof(Tile), cudaMemcpyHostToDevice));
        TFQ_CUDA(cudaMemcpy(p.d_coltile, coltile.data(), coltile.size()*sizeof(uint32_t), cudaMemcpyHostToDevice));
    }

    // ---- SpMM units: one block row times up to gmax of its block columns ----------------------------
    {
        int const TI = spmm_ti(is_double, LM, LN), TJ = spmm_tj(is_double, LN);
        // aim at ~128 threads per CTA: threads = (LM/TI) * (g*LN/TJ)
        // LM <= 8, measured on the block-size sweep (1728 block rows, 64 RHS columns) and the FD example: the small-block
        // kernel (4 entry groups of <= 32 threads, spmm.cu) wins in fp32 (4x4: 55 vs 97 us, 8x8: 95 vs 141 us; not at LN = 9)
        // and for tiny fp64 units (FD_problem.xml, one 8x8 block column: 3.0 vs 3.5 ms per solve); wide fp64 units stay on the ring
        {
            char const *env_small = std::getenv("TFQMRGPU_SMALL");
            bool const allowed = env_small ? (0 != std::atoi(env_small)) : true;
            int const cols_per_row = (p.max_cols_hint > 0) ? p.max_cols_hint : p.maxColsPerRow;   // a shard decides like the unsharded plan
            p.use_small = allowed && (LM <= 8) && (is_double ? (cols_per_row*LN <= 16) : (9 != LN));
        }
        int g = std::max(1, ((p.use_small ? 32 : 128)*TI*TJ)/(LM*LN));
        g = std::min(g, 16);
        g = std::min(g, p.maxColsPerRow);
        // keep one pipeline stage (A block + g X blocks, possibly k-chunked) reasonable: <= 48 KiB at 4 k-rows
        while (g > 1 && 2*4*(size_t(LM) + size_t(g)*LN)*s > 48*1024) --g;
        // TFQMRGPU_TENSOR: 0 SIMT kernels only, 1 (default) fp16-pair tcgen05 product (complex fp32) / DMMA (complex fp64)
        char const *env = std::getenv("TFQMRGPU_TENSOR");
        int const level = env ? std::atoi(env) : 1;
        p.use_tc16 = spmm_tc16_supported(LM, LN, precision, level);
        if (p.use_tc16) g = spmm_tc16_columns_per_unit(LM, LN);
        p.use_dmma = spmm_dmma_supported(LM, LN, precision) && (level >= 1);
        if (p.use_dmma) g = std::min(spmm_dmma_columns_per_unit(LM, LN), std::max(1, p.maxColsPerRow));
        p.gmax = uint32_t(g);
        std::vector<uint32_t> first, ng, urow;
        for (int r = 0; r < p.mb; ++r) {
            uint32_t const y0 = p.h_rowptrX[r], n = p.h_rowptrX[r + 1] - p.h_rowptrX[r];
            if (0 == n) continue;
            uint32_t const nu = (n + g - 1)/g;
            for (uint32_t u = 0; u < nu; ++u) {
                uint32_t const a = uint32_t((uint64_t(n)*u)/nu), b = uint32_t((uint64_t(n)*(u + 1))/nu);
                first.push_back(y0 + a); ng.push_back(b - a); urow.push_back(uint32_t(r));
            }
        }
        p.nUnits = uint32_t(first.size());
        DevArray<uint32_t> d_first, d_ng, d_cnt;
        TFQ_CUDA(d_first.alloc(first.size())); TFQ_CUDA(d_ng.alloc(ng.size())); TFQ_CUDA(d_cnt.alloc(first.size() + 1));
        TFQ_CUDA(cudaMemcpyAsync(d_first.ptr, first.data(), first.size()*4, cudaMemcpyHostToDevice, stream));
        TFQ_CUDA(cudaMemcpyAsync(d_ng.ptr, ng.data(), ng.size()*4, cudaMemcpyHostToDevice, stream));
        TFQ_CUDA(cudaMemsetAsync(d_cnt.ptr, 0, (first.size() + 1)*4, stream));
        TFQ_CUDA(cudaMalloc((void**)&p.d_unit_e0, (first.size() + 1)*4));
        TFQ_CUDA(cudaMalloc((void**)&p.d_unit_y, std::max<size_t>(first.size()*g, 1)*4));
        k_units<false><<<nblk(p.nUnits, 128), 128, 0, stream>>>(p.nUnits, g, d_first.ptr, d_ng.ptr, p.d_starts, p.d_pairs, p.d_perm,
                                                                 d_cnt.ptr, nullptr, nullptr, nullptr, nullptr);
        if (p.use_tc16) {
            // The persistent kernel runs one CTA per SM.  Deal the units, in block-row order, to the CTA with the fewest entries
            // so far (equal rows: plain round robin, so the CTAs work on neighbouring rows at any time and the X blocks they share
            // stay in L2; ragged rows: balanced entry counts), and store every CTA's units - and thereby its entries -
            // contiguously: the copy warp and the converter warps then stream over ONE flat entry range.
            std::vector<uint32_t> cnt(first.size() + 1, 0);
            TFQ_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt.ptr, first.size()*4, cudaMemcpyDeviceToHost, stream));
            TFQ_CUDA(cudaStreamSynchronize(stream));
            uint32_t const grid = std::max<uint32_t>(1, std::min<uint32_t>(uint32_t(nsm), p.nUnits));
            std::vector<std::vector<uint32_t>> mine(grid);
            {
                using Load = std::pair<uint64_t, uint32_t>;     // (entries so far, CTA)
                std::priority_queue<Load, std::vector<Load>, std::greater<Load>> q;
                for (uint32_t c = 0; c < grid; ++c) q.push({0, c});
                for (uint32_t u = 0; u < p.nUnits; ++u) {
                    Load l = q.top(); q.pop();
                    mine[l.second].push_back(u);
                    l.first += cnt[u] + 1;                      // (+1: a unit without entries still costs an epilogue)
                    q.push(l);
                }
            }
            std::vector<uint32_t> first2, ng2, urow2, e02(1, 0), cta_u0(1, 0);
            first2.reserve(first.size()); ng2.reserve(first.size()); urow2.reserve(first.size());
            for (uint32_t c = 0; c < grid; ++c) {
                for (uint32_t u : mine[c]) {
                    first2.push_back(first[u]); ng2.push_back(ng[u]); urow2.push_back(urow[u]);
                    e02.push_back(e02.back() + cnt[u]);
                }
                cta_u0.push_back(uint32_t(first2.size()));
            }
            first.swap(first2); ng.swap(ng2); urow.swap(urow2);
            p.tc_grid = grid;
            TFQ_CUDA(cudaMemcpyAsync(d_first.ptr, first.data(), first.size()*4, cudaMemcpyHostToDevice, stream));
            TFQ_CUDA(cudaMemcpyAsync(d_ng.ptr, ng.data(), ng.size()*4, cudaMemcpyHostToDevice, stream));
            TFQ_CUDA(cudaMemcpyAsync(p.d_unit_e0, e02.data(), e02.size()*4, cudaMemcpyHostToDevice, stream));
            TFQ_CUDA(cudaMalloc((void**)&p.d_cta_u0, cta_u0.size()*4));
            TFQ_CUDA(cudaMalloc((void**)&p.d_unit_row, std::max<size_t>(urow.size(), 1)*4));
            TFQ_CUDA(cudaMemcpyAsync(p.d_cta_u0, cta_u0.data(), cta_u0.size()*4, cudaMemcpyHostToDevice, stream));
            TFQ_CUDA(cudaMemcpyAsync(p.d_unit_row, urow.data(), urow.size()*4, cudaMemcpyHostToDevice, stream));
            TFQ_CUDA(cudaStreamSynchronize(stream));             // the host vectors are locals
            // Which form of the product: the planar one (four real products, spmm_tc16p.cu) is the fast one, but its error on sums that
            // cancel between the real products grows with the row length (measured ~8e-8 * entries * LM on the reference harness's cos/sin
            // fill, whose pass bar is 1e-4 absolute): plans whose longest row has entries * LM <= 864 (the 27-point stencil of 32 x 32
            // blocks: 7e-5) use it, longer rows the direct form (spmm_tc16.cu, ~1e-5 on the same operands).
            uint32_t max_entries = 0;
            for (uint32_t u = 0; u < p.nUnits; ++u) max_entries = std::max(max_entries, cnt[u]);
            char const *env_form = std::getenv("TFQMRGPU_TC_FORM");
            p.tc_planar = env_form ? ('p' == env_form[0]) : (uint64_t(max_entries)*uint64_t(LM) <= 864);
            char const *env_seg = std::getenv("TFQMRGPU_TC_CHAIN");   // entries per accumulation segment
            int const seg_env = env_seg ? std::atoi(env_seg) : 0;
            p.tc_seg = (seg_env > 0) ? seg_env : (p.tc_planar ? spmm_tc16p_default_segment(std::min(LM, 32)) : spmm_tc16_default_segment(std::min(LM, 32)));
        } else {
            TFQ_CUDA(exclusive_scan(p.d_unit_e0, d_cnt.ptr, first.size() + 1, stream));
        }
        uint32_t ne = 0;
        TFQ_CUDA(cudaMemcpy(&ne, p.d_unit_e0 + p.nUnits, 4, cudaMemcpyDeviceToHost));
        p.nEntries = ne;
        TFQ_CUDA(cudaMalloc((void**)&p.d_ent_a, std::max<size_t>(ne, 1)*4));
        TFQ_CUDA(cudaMalloc((void**)&p.d_ent_x, std::max<size_t>(size_t(ne)*g, 1)*4));
        k_units<true><<<nblk(p.nUnits, 128), 128, 0, stream>>>(p.nUnits, g, d_first.ptr, d_ng.ptr, p.d_starts, p.d_pairs, p.d_perm,
                                                                nullptr, p.d_unit_e0, p.d_unit_y, p.d_ent_a, p.d_ent_x);
        TFQ_CUDA(cudaStreamSynchronize(stream));
        TFQ_CUDA(cudaGetLastError());
        if (p.use_tc16 && 64 == LM) {
            // 64 x 64 blocks: virtual units of 32 x 32 sub-blocks (two per unit, two block columns each, four entries per entry)
            if (uint64_t(ne)*4 > 0xffffffffull) return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED);
            uint32_t *e0v = nullptr, *yv = nullptr, *rowv = nullptr, *av = nullptr, *xv = nullptr;
            size_t const nu = p.nUnits;
            TFQ_CUDA(cudaMalloc((void**)&e0v, (2*nu + 1)*4)); TFQ_CUDA(cudaMalloc((void**)&yv, std::max<size_t>(4*nu, 1)*4));
            TFQ_CUDA(cudaMalloc((void**)&rowv, std::max<size_t>(2*nu, 1)*4));
            TFQ_CUDA(cudaMalloc((void**)&av, std::max<size_t>(4*size_t(ne), 1)*4)); TFQ_CUDA(cudaMalloc((void**)&xv, std::max<size_t>(8*size_t(ne), 1)*4));
            k_expand_v64<<<nblk(nu + 1, 128), 128, 0, stream>>>(p.nUnits, p.d_unit_e0, p.d_unit_y, p.d_unit_row, p.d_ent_a, p.d_ent_x,
                                                                e0v, yv, rowv, av, xv);
            std::vector<uint32_t> cta_u0(size_t(p.tc_grid) + 1);
            TFQ_CUDA(cudaMemcpyAsync(cta_u0.data(), p.d_cta_u0, cta_u0.size()*4, cudaMemcpyDeviceToHost, stream));
            TFQ_CUDA(cudaStreamSynchronize(stream));
            for (auto &v : cta_u0) v *= 2;
            TFQ_CUDA(cudaMemcpy(p.d_cta_u0, cta_u0.data(), cta_u0.size()*4, cudaMemcpyHostToDevice));
            TFQ_CUDA(cudaGetLastError());
            cudaFree(p.d_unit_e0); cudaFree(p.d_unit_y); cudaFree(p.d_unit_row); cudaFree(p.d_ent_a); cudaFree(p.d_ent_x);
            p.d_unit_e0 = e0v; p.d_unit_y = yv; p.d_unit_row = rowv; p.d_ent_a = av; p.d_ent_x = xv;
            p.nUnits *= 2; p.nEntries = uint64_t(ne)*4; p.gmax = 2;
        }
    }

    // ---- workspace layout (all pieces 256-byte aligned like the reference's bump allocator) ----------
    {
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t const at = off; off = align256(off + bytes); return at; };
        p.vecBytes = size_t(p.nnzbX)*blockBytes;
        // v1 (X) first, then v4..v9 contiguous so that one memset clears them (core.hxx:114-125)
        p.off_v[1] = take(p.vecBytes);
        if (p.lean_vectors) {      // the fp64 side of a mixed-precision plan (mixed.cu) never iterates: X, the product and one scratch vector
            p.off_v[8] = take(p.vecBytes); p.off_v[9] = take(p.vecBytes);
            for (int v = 4; v <= 7; ++v) p.off_v[v] = p.off_v[9];
        } else {
            for (int v = 4; v <= 9; ++v) p.off_v[v] = take(p.vecBytes);
        }
        if (p.use_tc16) p.off_xop = take(p.vecBytes);                // fp16 pairs: the same 4 bytes per element
        p.off_v[3]  = take(size_t(p.nnzbX)*2*LM*LN*sizeof(float));   // v3 is always float (core.hxx:60)
        p.off_B     = take(size_t(p.nnzbB)*blockBytes);
        p.off_A     = take(size_t(p.nnzbA)*2*LM*LM*s + (p.use_tc16 ? size_t(p.mb)*4 : 0));
        p.off_ainv  = p.off_A + size_t(p.nnzbA)*2*LM*LM*s;            // 1/scale of the block rows: part of the 'A' window
        p.off_zero  = take(blockBytes);
        size_t const sc = size_t(nb)*2*LN*s;
        p.off_rho = take(sc); p.off_alfa = take(sc); p.off_beta = take(sc); p.off_c67 = take(sc); p.off_eta = take(sc);
        p.off_tau = take(size_t(nb)*LN*8); p.off_var = take(size_t(nb)*LN*8); p.off_invBn2 = take(size_t(nb)*LN*8);
        p.off_status = take(size_t(nb)*LN); p.off_snap = take(size_t(nb)*LN);
        p.off_part   = take(size_t(p.nTiles)*kPartD*LN*8);
        p.off_colmon = take(size_t(nb)*4*8);
        p.off_ticket = take((size_t(nb) + 8)*4);
        p.off_ctl    = take(sizeof(Control));
        if (p.use_tc16) {
            p.off_xs = take(size_t(nb)*LN*4); p.off_xsinv = take(size_t(nb)*LN*4);
            p.off_ablkmax = take(size_t(p.nnzbA)*4); p.off_arowscale = take(size_t(p.mb)*4);
            p.off_xpart = take(size_t(p.nTiles)*64*4);
            p.off_mx = take(3*size_t(nb)*LN*4);                       // column maxima of |v4|, |v5|, |v6| (vecops.cu)
        }
        p.bufferBytes = off + 256;
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace tfq
