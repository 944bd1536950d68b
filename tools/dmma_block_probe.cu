// Development micro-benchmark: the consumer's matrix block (28 MMAs m8n8k4 FP64, B fragments by LDS.128 from shared memory,
// A fragment by LDS.64) as a lone warp per SM sub-partition, in three code shapes:
//   0 interleaved (compiler schedules loads among the MMAs)      1 all fragments loaded, compiler barrier, then the MMAs
//   2 as 1 with two tiles per pass (B fragments reused, 8 accumulators)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int KS = 7, MP = 2, MT = 4, FR = MP * KS * 64, NS = 28;
template <int V>
__global__ void probe(double *out, long long *cycles, int iters)
{
    extern __shared__ __align__(16) double sm[];
    double *ops = sm;                 // 4 operators
    double *vec = sm + 4 * FR;        // 2 tiles x 4 vectors x 8 x NS
    for (int i = threadIdx.x; i < 4 * FR + 2 * 4 * 8 * NS; i += blockDim.x) sm[i] = 1.0 / (1 + (i % 97));
    __syncthreads();
    const int lane = threadIdx.x & 31, q = lane >> 2, r = lane & 3;
    double acc[2][MT][2];
    for (int t = 0; t < 2; ++t)
        for (int I = 0; I < MT; ++I) acc[t][I][0] = acc[t][I][1] = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll 1
        for (int m = 0; m < 4; ++m)
        {
            const double2 *bp = reinterpret_cast<const double2 *>(ops + m * FR) + lane;
            const double *va = vec + (size_t)m * 8 * NS + q * NS + r;
            if (V == 0)
            {
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                {
                    const double x = va[4 * kk];
#pragma unroll
                    for (int I = 0; I < MT; ++I)
                    {
                        const double2 b = bp[((I >> 1) * KS + kk) * 32];
                        dmma(acc[0][I][0], acc[0][I][1], x, (I & 1) ? b.y : b.x);
                    }
                }
            }
            else
            {
                double2 bq[KS][MP];
                double xa[KS], xb[KS];
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                {
#pragma unroll
                    for (int Ip = 0; Ip < MP; ++Ip) bq[kk][Ip] = bp[(Ip * KS + kk) * 32];
                    xa[kk] = va[4 * kk];
                    if (V == 2) xb[kk] = va[4 * 8 * NS + 4 * kk];
                }
                asm volatile("" ::: "memory");
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
#pragma unroll
                    for (int I = 0; I < MT; ++I)
                    {
                        dmma(acc[0][I][0], acc[0][I][1], xa[kk], (I & 1) ? bq[kk][I >> 1].y : bq[kk][I >> 1].x);
                        if (V == 2) dmma(acc[1][I][0], acc[1][I][1], xb[kk], (I & 1) ? bq[kk][I >> 1].y : bq[kk][I >> 1].x);
                    }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int t = 0; t < 2; ++t)
        for (int I = 0; I < MT; ++I) s += acc[t][I][0] + acc[t][I][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int V>
void run(int warps, double *out, long long *cyc)
{
    const int iters = 500;
    const size_t smem = (4 * FR + 2 * 4 * 8 * NS) * sizeof(double);
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<V><<<148, warps * 32, smem>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double mmas = (double)iters * 4 * KS * MT * (V == 2 ? 2 : 1);
    printf("shape %d warps/SM %2d: %6.2f cycles per MMA per warp  (%s)\n", V, warps, c / mmas, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&cyc, 8);
    for (int w : {4, 8})
    {
        run<0>(w, out, cyc);
        run<1>(w, out, cyc);
        run<2>(w, out, cyc);
    }
    return 0;
}
