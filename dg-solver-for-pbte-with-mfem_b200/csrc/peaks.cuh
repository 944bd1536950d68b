// Micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry:
// FP64 FMA pipe (DFMA), FP64 tensor pipe (DMMA m8n8k4) and a plain device copy (cross-check of hbm_gbs).
#pragma once

#include <cuda_runtime.h>

namespace pbte_dev
{
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i)
    {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double seed)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = seed;
    const double a = 1e-3 * (threadIdx.x & 7), b = 1e-3;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

inline cudaError_t measure_peaks(double *dfma_tflops, double *dmma_tflops, double *copy_gbs)
{
    cudaError_t e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int blocks = sms * 8, threads = 256;
    double *out = nullptr;
    if ((e = cudaMalloc(&out, sizeof(double) * blocks * threads)) != cudaSuccess) return e;
    float ms = 0.f;
    auto best_of = [&](auto launch, int reps) {
        float best = 1e30f;
        for (int r = 0; r < reps; ++r)
        {
            cudaEventRecord(t0);
            launch();
            cudaEventRecord(t1);
            cudaEventSynchronize(t1);
            cudaEventElapsedTime(&ms, t0, t1);
            if (r > 0 && ms < best) best = ms;
        }
        return best;
    };
    {
        const int iters = 20000;
        const float t = best_of([&] { dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0); }, 4);
        if (dfma_tflops) *dfma_tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (t * 1e-3) / 1e12;
    }
    {
        const int iters = 4000;
        const float t = best_of([&] { dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0); }, 4);
        // one m8n8k4 = 8*8*4 FMA per warp
        if (dmma_tflops) *dmma_tflops = 2.0 * 256.0 * 8.0 * iters * (double)blocks * (threads / 32) / (t * 1e-3) / 1e12;
    }
    cudaFree(out);
    {
        const size_t bytes = (size_t)1 << 30;
        double2 *a = nullptr, *b = nullptr;
        if ((e = cudaMalloc(&a, bytes)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&b, bytes)) != cudaSuccess) { cudaFree(a); return e; }
        cudaMemset(a, 0, bytes);
        const float t = best_of([&] { copy_kernel<<<sms * 16, 512>>>(a, b, bytes / sizeof(double2)); }, 6);
        if (copy_gbs) *copy_gbs = 2.0 * (double)bytes / (t * 1e-3) / 1e9;
        cudaFree(a);
        cudaFree(b);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return cudaGetLastError();
}
}  // namespace pbte_dev
