// Microbenchmark (dev tool, not part of the library): issue rate of tcgen05.mma on B200 for the shapes of the round-1 3xTF32 product (kept as the measurement behind DESIGN.md: one MMA per ~54 cycles for N <= 64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu && tools/mma_rate
// One CTA per SM, one elected lane issues `reps` MMAs back to back on static operands and commits to an mbarrier;
// clock64 around issue+completion gives cycles per MMA.  kind: 0 = tf32 (K=8), 1 = bf16 (K=16).  mode: 0 = SS, 1 = TS.
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
template <int KIND, int MODE>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0 && MODE == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" :: "r"(d), "r"(a_t), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 0 && MODE == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a_d), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 1 && MODE == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" :: "r"(d), "r"(a_t), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 1 && MODE == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a_d), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int ndst, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t*>(smem + 64);
    float *ops = reinterpret_cast<float*>(smem + 1024);
    for (int i = threadIdx.x; i < 48*1024/4; i += blockDim.x) ops[i] = 1.0f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t const tm = *slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        uint32_t const leader = elect_one_sync();
        uint32_t const fmt = (KIND == 0) ? 2u : 1u;   // tf32 : bf16
        uint32_t const idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
        uint64_t const bd = desc_noswizzle(smem_u32(ops), 2048, 128);
        uint64_t const ad = desc_noswizzle(smem_u32(ops) + 16384, 2048, 128);
        t0 = clock64();
        if (leader) {
            for (int r = 0; r < reps; ++r) {
                uint32_t const d = tm + uint32_t(r % ndst)*uint32_t(N);       // ndst independent accumulators (1 = fully dependent chain)
                mma<KIND, MODE>(d, tm + 448, ad, bd, idesc, r >= ndst ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
        t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
}

template <int KIND, int MODE> void run(char const *name, int N, int ndst, long long *d_out) {
    int const reps = 4096;
    auto k = rate_kernel<KIND, MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64*1024);
    long long h = 0;
    for (int it = 0; it < 2; ++it) { k<<<148, 128, 64*1024>>>(N, reps, ndst, d_out); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    double const cyc = double(h)/reps;
    int const K = KIND == 0 ? 8 : 16;
    std::printf("%-10s N=%3d accumulators=%d : %7.1f cycles/MMA  -> %.0f flop/clk/SM (%s)\n", name, N, ndst, cyc, 2.0*128*N*K/cyc,
                cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long *d_out; cudaMalloc(&d_out, 8);
    for (int N : {64, 128, 256}) for (int nd : {1, 2}) {
        if (N*nd > 448) continue;
        run<0, 1>("tf32 TS", N, nd, d_out);
        run<0, 0>("tf32 SS", N, nd, d_out);
        run<1, 1>("bf16 TS", N, nd, d_out);
        run<1, 0>("bf16 SS", N, nd, d_out);
    }
    return 0;
}
