// FP32 FMA throughput microbenchmark (roofline denominator for the RANSAC hypothesis kernel):
// every thread runs 8 independent FFMA chains; reports TFLOP/s (FMA = 2 FLOP).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    float* out;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(s);
        fma_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms = 0;
        cudaEventElapsedTime(&ms, s, e);
        const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        printf("rep %d: %.3f ms  %.2f TFLOP/s\n", rep, ms, tf);
    }
    printf("{\"fp32_fma_tflops\": %.2f, \"sms\": %d}\n", best, sms);
    return 0;
}
