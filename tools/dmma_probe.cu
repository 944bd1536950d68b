// Micro-benchmark (development tool, not part of the library): latency / issue rate of mma.sync.m8n8k4.f64 on sm_100a
// as a function of independent accumulator chains per warp (ILP) and warps per SM, with and without an LDS per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP, bool LDS>
__global__ void probe(double *out, long long *cycles, int iters)
{
    __shared__ double sm[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = 1.0 / (1 + i);
    __syncthreads();
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9;
    double b = 1.0 - threadIdx.x * 1e-9;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i)
            {
                if (LDS) b = sm[((r * ILP + i) & 63) * 32 + lane];
                dmma(acc[i][0], acc[i][1], a, b);
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP, bool LDS>
void run(int warps_per_sm, double *out, long long *cyc)
{
    const int iters = 2000;
    probe<ILP, LDS><<<148, warps_per_sm * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<ILP, LDS><<<148, warps_per_sm * 32>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * ILP;  // MMAs per warp
    const double tf = n * warps_per_sm * 148 * 512.0 / (ms * 1e-3) / 1e12;
    printf("ILP %d LDS %d warps/SM %2d : %7.2f cycles per MMA per warp, %6.2f TFLOP/s\n", ILP, (int)LDS, warps_per_sm, c / n, tf);
}

int main()
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    for (int w : {1, 4, 8, 16, 32})
    {
        run<1, false>(w, out, cyc);
        run<2, false>(w, out, cyc);
        run<4, false>(w, out, cyc);
        run<8, false>(w, out, cyc);
        run<4, true>(w, out, cyc);
        run<8, true>(w, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
