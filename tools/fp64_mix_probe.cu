// Micro-benchmark (development tool, not part of the library): do mma.sync.m8n8k4.f64 (DMMA) and DFMA share one FP64 pipe
// on sm_100a, or can they run side by side?  Three kernels with the same structure: MMAs only, FMAs only, and both
// interleaved (NF warp-wide FMAs per MMA).  If the pipes are shared the mixed time is the sum of the two, if not the maximum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix_probe fp64_mix_probe.cu && ./fp64_mix_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// per loop trip: NM independent MMAs (4 chains) and NF independent FMAs (8 chains)
template <int NM, int NF>
__global__ void __launch_bounds__(512) probe(double *out, int iters)
{
    double acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 1.0 + i + threadIdx.x * 1e-9;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9, m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(acc[i & 3][0], acc[i & 3][1], a, b);
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i & 7] = fma(f[i & 7], m, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
float run(int warps_per_sm, double *out)
{
    const int iters = 4000;
    probe<NM, NF><<<148, warps_per_sm * 32>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<NM, NF><<<148, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double nm = (double)iters * 8 * NM, nf = (double)iters * 8 * NF;  // per warp
    const double tf = (nm * 512.0 + nf * 64.0) * warps_per_sm * 148 / (ms * 1e-3) / 1e12;
    printf("warps/SM %2d  MMA %d + FMA %2d per trip: %8.3f ms  %6.2f TFLOP/s (both kinds counted)\n", warps_per_sm, NM, NF, ms, tf);
    return ms;
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    for (int w : {4, 8, 16})
    {
        const float a = run<4, 0>(w, out);
        const float b8 = run<0, 8>(w, out);
        const float b32 = run<0, 32>(w, out);
        const float c8 = run<4, 8>(w, out);
        const float c32 = run<4, 32>(w, out);
        printf("  -> 4 MMA + 8 FMA : mixed %.3f  sum %.3f  max %.3f\n", c8, a + b8, a > b8 ? a : b8);
        printf("  -> 4 MMA + 32 FMA: mixed %.3f  sum %.3f  max %.3f\n", c32, a + b32, a > b32 ? a : b32);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
