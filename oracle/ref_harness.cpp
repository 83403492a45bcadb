// TEST INFRASTRUCTURE (oracle). Not part of the product.
//
// Small C-callable harness that is compiled TOGETHER with the unmodified reference
// (oracle/build_ref.sh) and peeks into the reference's plan object so that tests can
//   * dump the reference's plan lists (pairs/starts/subset/colindx)  -> bit-exact plan parity,
//   * read / overwrite the reference's random shadow vector v3        -> same-v3 solver parity,
//   * allocate a 256-byte aligned workspace for the CPU build (the reference's bump allocator
//     assumes that alignment, plain malloc() breaks it; SURVEY.md §8c caveat 2).
// It includes the reference's headers from /root/reference at build time; no reference code is copied.
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef HAS_NO_CUDA
  #include <cuda_runtime.h>   // before the reference headers: tfqmrgpu.hxx #defines devPtr
#endif
#include "tfqmrgpu.hxx"       // reference: cuda.h or cudaStubs, tfqmrgpu.h, colIndex_t
#include "tfqmrgpu_plan.hxx"  // reference: bsrsv_plan_t (tfqmrgpu_plan.hxx:9-55)

namespace {
    inline bsrsv_plan_t* P(void* plan) { return reinterpret_cast<bsrsv_plan_t*>(plan); }

    // windows are buffer-relative after bufferSize() and absolute after the first solve()
    inline char* window_ptr(bsrsv_plan_t const* p, memWindow_t const& w) {
        size_t const base = size_t(p->pBuffer);
        if (w.offset >= base && base != 0) return reinterpret_cast<char*>(w.offset);
        return p->pBuffer + w.offset;
    }

    inline void copy_out(void* host, void const* dev, size_t n) {
#ifdef HAS_NO_CUDA
        std::memcpy(host, dev, n);
#else
        cudaMemcpy(host, dev, n, cudaMemcpyDeviceToHost);
#endif
    }
    inline void copy_in(void* dev, void const* host, size_t n) {
#ifdef HAS_NO_CUDA
        std::memcpy(dev, host, n);
#else
        cudaMemcpy(dev, host, n, cudaMemcpyHostToDevice);
#endif
    }
} // namespace

extern "C" {

    // out[0..7] = nnzbX, nnzbB, nCols, nPairs, LM, LN, nRows, nnzbA
    void refh_plan_counts(void* plan, uint64_t out[8]) {
        auto const p = P(plan);
        out[0] = p->colindx.size();
        out[1] = p->subset.size();
        out[2] = p->nCols;
        out[3] = p->pairs.size()/2;
        out[4] = p->LM;
        out[5] = p->LN;
        out[6] = p->nRows;
        out[7] = p->nnzbA;
    }

    // kind: 0 starts u32[nnzbX+1], 1 pairs u32[2*nPairs], 2 subset u32[nnzbB], 3 colindx u16[nnzbX]
    int refh_plan_copy(void* plan, int kind, void* out) {
        auto const p = P(plan);
        switch (kind) {
            case 0: std::memcpy(out, p->starts.data(),  p->starts.size()*sizeof(uint32_t)); return 0;
            case 1: std::memcpy(out, p->pairs.data(),   p->pairs.size()*sizeof(uint32_t)); return 0;
            case 2: std::memcpy(out, p->subset.data(),  p->subset.size()*sizeof(uint32_t)); return 0;
            case 3: std::memcpy(out, p->colindx.data(), p->colindx.size()*sizeof(uint16_t)); return 0;
        }
        return -1;
    }

    // number of floats in the reference's v3 window
    uint64_t refh_v3_count(void* plan) { return P(plan)->vec3win.length/sizeof(float); }

    // read (to_host=1) or overwrite (to_host=0) v3, layout float[nnzbX][2][LM][LN] in X's BSR order
    int refh_v3_copy(void* plan, float* host, int to_host) {
        auto const p = P(plan);
        if (nullptr == p->pBuffer) return -1;
        char* const v3 = window_ptr(p, p->vec3win);
        if (to_host) copy_out(host, v3, p->vec3win.length);
        else         copy_in (v3, host, p->vec3win.length);
        return 0;
    }

    // raw read of the X window in the reference's internal layout real_t[nnzbX][2][LM][LN]
    int refh_x_copy(void* plan, void* host) {
        auto const p = P(plan);
        if (nullptr == p->pBuffer) return -1;
        copy_out(host, window_ptr(p, p->matXwin), p->matXwin.length);
        return 0;
    }

    void* refh_aligned_alloc(size_t bytes) {
        void* ptr = nullptr;
        if (0 != posix_memalign(&ptr, 256, ((bytes + 255)/256)*256)) return nullptr;
        return ptr;
    }
    void refh_aligned_free(void* ptr) { std::free(ptr); }

    int refh_is_cpu_build(void) {
#ifdef HAS_NO_CUDA
        return 1;
#else
        return 0;
#endif
    }

} // extern "C"
