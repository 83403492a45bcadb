// FP64 rate of this GPU, measured: the denominator of the complex-fp64 product's roofline (MEASURED_PEAKS.json holds no fp64 figure).
//   dfma_tflops : independent DFMA chains, 2 flop per instruction and lane
//   dmma_tflops : mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4), 512 flop per warp instruction
// Prints "dfma_tflops=<x> dmma_tflops=<y>" (best of 5 launches each, CUDA events).   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kChains = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b) {
    double acc[kChains];
    #pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < kIters; ++i) {
        #pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
    #pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    if (s == 123.456) out[0] = s;          // keep the chains alive
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, double a, double b) {
    double c0[kChains], c1[kChains];
    #pragma unroll
    for (int c = 0; c < kChains; ++c) { c0[c] = threadIdx.x; c1[c] = c; }
    for (int i = 0; i < kIters; ++i) {
        #pragma unroll
        for (int c = 0; c < kChains; ++c)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
    }
    double s = 0;
    #pragma unroll
    for (int c = 0; c < kChains; ++c) s += c0[c] + c1[c];
    if (s == 123.456) out[0] = s;
}

template <typename K> static double best_ms(K kernel, int grid, double *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        kernel<<<grid, 256>>>(d, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int dev = 0, sms = 0;
    if (cudaSuccess != cudaGetDevice(&dev)) { std::printf("no device\n"); return 1; }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d = nullptr; cudaMalloc(&d, 64);
    int const grid = sms*8;
    double const ms_f = best_ms(dfma_kernel, grid, d), ms_m = best_ms(dmma_kernel, grid, d);
    double const threads = double(grid)*256;
    double const dfma = threads*kIters*kChains*2.0/(ms_f*1e-3)*1e-12;
    double const dmma = (threads/32)*kIters*kChains*512.0/(ms_m*1e-3)*1e-12;
    if (cudaSuccess != cudaGetLastError()) { std::printf("cuda error\n"); return 1; }
    std::printf("dfma_tflops=%.3f dmma_tflops=%.3f\n", dfma, dmma);
    cudaFree(d);
    return 0;
}
