// Microbenchmark (dev tool): the MMA sequence of spmm_tc_kernel<32,32> per A block ("entry"), without any data movement:
//   4 k-steps x [ N'=128 MMA (Xhi * [Ahi;Alo]) + N'=64 MMA (Xlo * Ahi) ], TF32, M=128, A operand in TMEM, B in shared memory,
// cycling through 2 TMEM stages and 6 ring slots, commit per entry.  1 or 2 CTAs per SM.  Reports cycles per entry.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_entry_rate tools/mma_entry_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(void const *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
    return pred;
}
__device__ __forceinline__ uint64_t desc_noswizzle(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((saddr >> 4) & 0x3fff) | (uint64_t((lbo >> 4) & 0x3fff) << 16) | (uint64_t((sbo >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(288, 2) k(int entries, int same_operands, int split_acc, int serial, int bg, volatile int *stop, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t*>(smem + 64);
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(smem + 72) = 0;
    float *ops = reinterpret_cast<float*>(smem + 1024);
    for (int i = threadIdx.x; i < 96*1024/4; i += blockDim.x) ops[i] = 1.0f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t const tm = *slot;
    if (threadIdx.x < 32) {
        uint32_t const leader = elect_one_sync();
        uint32_t const base = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(128 >> 4) << 24);
        uint32_t const id128 = base | (uint32_t(128 >> 3) << 17), id64 = base | (uint32_t(64 >> 3) << 17);
        long long t0 = clock64();
        if (leader) {
            for (int e = 0; e < entries; ++e) {
                int const s = same_operands ? 0 : (e & 1), r = same_operands ? 0 : (e % 6);
                uint32_t const xa = tm + 128 + uint32_t(s)*64, sa = smem_u32(ops) + uint32_t(r)*16384;
                #pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint64_t const b = desc_noswizzle(sa + ks*4096, 2048, 128);
                    mma_ts(tm, xa + 8*ks, b, id128, (e > 0 || ks > 0) ? 1u : 0u);
                    mma_ts(tm + (split_acc ? 64 : 0), xa + 32 + 8*ks, b, id64, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
                // (no wait: the commits just accumulate arrivals on phases; we only wait for the last one below)
                if (serial && e + 1 < entries) { uint32_t ok = 0; while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(uint32_t(e & 1)) : "memory"); }
            }
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(uint32_t((entries - 1) & 1)) : "memory");
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(smem + 72) = 1;   // stop the background warps
    } else if (bg) {
        // background traffic while the MMAs run: bg&1 = tcgen05.st into the operand stages (like the converters),
        // bg&2 = shared-memory load/store of 16-byte chunks (like the lo computation)
        int const w = threadIdx.x >> 5, q = w & 3;
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = 0x3f800000u;
        float4 *sm4 = reinterpret_cast<float4*>(ops) + 2048;
        while (0 == *reinterpret_cast<volatile int*>(smem + 72)) {
            if (bg & 1) {
                uint32_t const t0 = tm + (uint32_t(32*q) << 16) + 128 + 16*((w >> 2) & 1);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                    :: "r"(t0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            if (bg & 2) { float4 x = sm4[threadIdx.x]; x.x += 1.f; sm4[threadIdx.x + 512] = x; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(256u) : "memory");
}
int main() {
    long long *d_out; cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100*1024);
    int const entries = 2048;
    for (int ctas = 1; ctas <= 2; ++ctas) for (int serial = 0; serial <= 1; ++serial) for (int bg = 0; bg <= 3; ++bg) {
        long long h = 0;
        for (int it = 0; it < 2; ++it) { k<<<148*ctas, 288, 100*1024>>>(entries, 0, 1, serial, bg, nullptr, d_out); cudaDeviceSynchronize(); }
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        std::printf("CTAs/SM=%d wait-per-entry=%d background(1=tcgen05.st,2=lds/sts)=%d : %7.1f cycles per entry per CTA  [%s]\n",
                    ctas, serial, bg, double(h)/entries, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
