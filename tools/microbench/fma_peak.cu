// FP32 throughput microbenchmark (roofline denominator for the RANSAC hypothesis kernel).
// Two kernels: scalar FFMA (8 independent chains per thread) and Blackwell's packed FFMA2 (8 independent float2
// chains per thread = 16 FMAs per instruction group).  Reports TFLOP/s (FMA = 2 FLOP) for both, best of 5 after a
// warm-up, as one JSON line; `python tools/microbench/run_fma_peak.py` stores it in profiles/fp32_peak.json.
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

__global__ void __launch_bounds__(256) fma2_kernel(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = make_float2(threadIdx.x + k, threadIdx.x + k + 0.5f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __ffma2_rn(x[k], a2, b2);
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k].x + x[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class Kern>
static double run(Kern kern, float* out, int blocks, int threads, int iters, double fma_per_thread_iter, const char* name) {
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(s);
        kern<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms = 0;
        cudaEventElapsedTime(&ms, s, e);
        const double flop = 2.0 * fma_per_thread_iter * iters * (double)blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        fprintf(stderr, "%s rep %d: %.3f ms  %.2f TFLOP/s\n", name, rep, ms, tf);
    }
    return best;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    float* out;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    const double ffma = run(fma_kernel, out, blocks, threads, iters, 64.0, "FFMA ");
    const double ffma2 = run(fma2_kernel, out, blocks, threads, iters, 128.0, "FFMA2");
    printf("{\"fp32_fma_tflops\": %.2f, \"fp32_fma2_tflops\": %.2f, \"sms\": %d, \"max_sm_mhz\": %d, "
           "\"nominal_tflops_at_max_clock\": %.2f}\n",
           ffma, ffma2, sms, khz / 1000, 2.0 * 128 * sms * (khz * 1e3) / 1e12);
    return 0;
}
