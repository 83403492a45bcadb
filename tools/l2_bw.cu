// Measures what an SM-side reader can pull out of L2 and out of HBM on this GPU, by two paths: LDG.128 and bulk copies
// (cp.async.bulk global -> shared).  The working set is either L2-resident (32 MB) or far larger than L2 (2 GB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_bw tools/l2_bw.cu && tools/l2_bw
// Used for DESIGN.md 4.1: the tcgen05 block product re-reads X blocks through L2 (27x for the stencil), so its bound
// is the L2 -> SM rate, not HBM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(void const *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(256) ldg_kernel(float4 const *src, size_t n_vec, int reps, float *sink)
{
    float4 acc = make_float4(0, 0, 0, 0);
    size_t const stride = size_t(gridDim.x)*blockDim.x;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n_vec; i += 4*stride) {
            float4 v[4];
            #pragma unroll
            for (int k = 0; k < 4; ++k) { size_t j = i + k*stride; v[k] = (j < n_vec) ? __ldcg(src + j) : make_float4(0, 0, 0, 0); }
            #pragma unroll
            for (int k = 0; k < 4; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}

// each CTA: ring of kSlots x kChunk bytes, one elected thread issues the copies and waits; nobody reads the data
template <int kChunk, int kSlots>
__global__ void __launch_bounds__(32) bulk_kernel(unsigned char const *src, size_t n_chunks, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t*>(smem);
    unsigned char *ring = smem + 1024;
    if (0 == threadIdx.x) {
        for (int s = 0; s < kSlots; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        size_t const per = (n_chunks + gridDim.x - 1)/gridDim.x;
        size_t const c0 = size_t(blockIdx.x)*per, c1 = (c0 + per < n_chunks) ? c0 + per : n_chunks;
        size_t issued = 0, waited = 0, total = (c1 > c0 ? c1 - c0 : 0)*size_t(reps);
        auto issue = [&](size_t k) {
            int const s = int(k % kSlots);
            size_t const c = c0 + k % (c1 - c0);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[s])), "r"(kChunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(ring + size_t(s)*kChunk)), "l"(src + c*kChunk), "r"(kChunk), "r"(smem_u32(&bar[s])) : "memory");
        };
        for (; issued < kSlots && issued < total; ++issued) issue(issued);
        for (; waited < total; ++waited) {
            int const s = int(waited % kSlots); unsigned const parity = unsigned((waited / kSlots) & 1);
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
            if (issued < total) { issue(issued); ++issued; }
        }
    }
}

template <typename F> float time_ms(F &&f, int n = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < n; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms/n;
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    size_t const big = size_t(2) << 30;
    unsigned char *buf = nullptr; float *sink = nullptr;
    if (cudaSuccess != cudaMalloc(&buf, big)) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 4); cudaMemset(buf, 1, big);
    constexpr int kChunk = 8192, kSlots = 12;
    size_t const smem = 1024 + size_t(kChunk)*kSlots;
    cudaFuncSetAttribute(bulk_kernel<kChunk, kSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    for (int pass = 0; pass < 2; ++pass) {
        size_t const bytes = pass ? big : (size_t(32) << 20);
        int const reps = pass ? 1 : 64;
        double const total = double(bytes)*reps;
        float ms = time_ms([&] { ldg_kernel<<<sms*8, 256>>>(reinterpret_cast<float4 const*>(buf), bytes/16, reps, sink); });
        printf("%-12s LDG.128      %8.1f GB/s\n", pass ? "HBM (2 GB)" : "L2 (32 MB)", total/ms*1e-6);
        for (int ctas = 1; ctas <= 2; ++ctas) {
            ms = time_ms([&] { bulk_kernel<kChunk, kSlots><<<sms*ctas, 32, smem>>>(buf, bytes/kChunk, reps); });
            printf("%-12s bulk 8 KB x%d  %8.1f GB/s   (%d CTAs/SM, %d copies in flight each)\n", pass ? "HBM (2 GB)" : "L2 (32 MB)", ctas, total/ms*1e-6, ctas, kSlots);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (cudaSuccess != e) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
