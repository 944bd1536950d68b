// Development micro-benchmark: how fast can one warp gather 40 rows of 224 B (one marching tile: 8 elements x 5 vectors of
// 28 doubles) from L2-resident global memory into shared memory on sm_100a?  Variants:
//   1 cp.async 16 B, one row per lane (lanes own rows; 14 instructions, 16 B per lane per instruction, scattered)
//   2 cp.async 16 B, chunks in memory order over the lanes (coalesced 512 B per instruction)
//   3 cp.async.bulk (TMA) one 224 B row per lane, completion on an mbarrier
//   4 ld.global.v2.f64 + st.shared.v2.f64 in memory order (register staged)
//   5 cp.async 16 B, a lane owns one chunk column, two whole rows per instruction (the pattern of sweep_regop.cuh), zero-fill operand
//   6 as 5, but issue only: no wait at all inside the loop (cycles a warp needs to get a tile's copies out)
// Each CTA = W producer warps, each warp gathers its own sequence of tiles (pseudo-random rows of a 0.9 MB region).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int ROWB = 224, NROW = 40, TILEB = ROWB * NROW, CH = ROWB / 16;

template <int V>
__global__ void probe(const unsigned char *g, int nrows_g, int tiles, long long *cyc, double *sink, int depth)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bars[8 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *my = sm + (size_t)warp * depth * TILEB;
    uint64_t *bar = bars + warp * 8;
    if (lane == 0)
        for (int d = 0; d < depth; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + d)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned char *base = g + (size_t)blockIdx.x * nrows_g * ROWB;  // per-CTA region
    uint32_t rng = 1234567u + blockIdx.x * 97u + warp * 13u;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t)
    {
        unsigned char *dst = my + (size_t)(t % depth) * TILEB;
        // rows of this tile: lane l (< 32) and l+32 (< 40)
        rng = rng * 1664525u + 1013904223u;
        const int r0 = (int)((rng >> 8) % (unsigned)nrows_g);
        const int r1 = (int)((rng >> 12) % (unsigned)nrows_g);
        if (V == 1)
        {
            const unsigned char *s0 = base + (size_t)r0 * ROWB, *s1 = base + (size_t)r1 * ROWB;
#pragma unroll
            for (int c = 0; c < CH; ++c)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + lane * ROWB + c * 16)), "l"(s0 + c * 16) : "memory");
            if (lane < 8)
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + (32 + lane) * ROWB + c * 16)), "l"(s1 + c * 16) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else if (V == 2)
        {
#pragma unroll
            for (int i0 = 0; i0 < NROW * CH; i0 += 32)
            {
                const int idx = i0 + lane, row = idx / CH, c = idx - row * CH;
                const int rr = __shfl_sync(0xffffffffu, row < 32 ? r0 : r1, row & 31);
                if (idx < NROW * CH)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + idx * 16)), "l"(base + (size_t)rr * ROWB + c * 16) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else if (V == 5 || V == 6)
        {
            const int r_sub = lane / CH, c = lane - r_sub * CH;
            const unsigned char *cb = base + c * 16;
            const uint32_t d0 = s32(dst) + r_sub * ROWB + c * 16;
#pragma unroll
            for (int k = 0; k < NROW / 2; ++k)
            {
                const int row = 2 * k + r_sub;
                const int rr = __shfl_sync(0xffffffffu, row < 32 ? r0 : r1, row & 31);
                if (r_sub < 2)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + k * 2 * ROWB), "l"(cb + (size_t)rr * ROWB), "r"(rr >= 0 ? 16u : 0u) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (V == 5)
            {
                if (depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
                else asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
        }
        else if (V == 3)
        {
            uint64_t *b = bar + (t % depth);
            if (t >= depth)
            {
                // wait for the previous use of this buffer
                uint32_t ok = 0;
                const uint32_t par = ((t / depth) - 1) & 1;
                while (!ok)
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
            }
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(TILEB) : "memory");
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst + lane * ROWB)),
                         "l"(base + (size_t)r0 * ROWB), "r"(ROWB), "r"(s32(b))
                         : "memory");
            if (lane < 8)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst + (32 + lane) * ROWB)),
                             "l"(base + (size_t)r1 * ROWB), "r"(ROWB), "r"(s32(b))
                             : "memory");
        }
        else
        {
            double2 v[9];
#pragma unroll
            for (int k = 0; k < 9; ++k)
            {
                const int idx = k * 32 + lane, row = idx / CH, c = idx - row * CH;
                const int rr = __shfl_sync(0xffffffffu, row < 32 ? r0 : r1, row & 31);
                v[k] = idx < NROW * CH ? __ldcg(reinterpret_cast<const double2 *>(base + (size_t)rr * ROWB + c * 16)) : make_double2(0, 0);
            }
#pragma unroll
            for (int k = 0; k < 9; ++k)
            {
                const int idx = k * 32 + lane;
                if (idx < NROW * CH) *reinterpret_cast<double2 *>(dst + idx * 16) = v[k];
            }
        }
        if (V != 3 && V != 6) acc += reinterpret_cast<const double *>(dst)[lane];
    }
    if (V == 3)
        for (int d = 0; d < depth && d < tiles; ++d)
        {
            const int t = tiles - 1 - d;
            uint64_t *b = bar + (t % depth);
            uint32_t ok = 0;
            const uint32_t par = (t / depth) & 1;
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
            acc += reinterpret_cast<const double *>(my + (size_t)(t % depth) * TILEB)[lane];
        }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
void run(const unsigned char *g, int nrows_g, int warps, int depth, int ctas, long long *cyc, double *sink)
{
    const int tiles = 2000;
    const size_t smem = (size_t)warps * depth * TILEB;
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<V><<<ctas, warps * 32, smem>>>(g, nrows_g, tiles, cyc, sink, depth);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("variant %d warps/CTA %d depth %d CTAs %3d: %8.1f cycles per tile per warp, %6.2f B/clk/SM  %s\n", V, warps, depth, ctas, (double)c / tiles,
           (double)TILEB * tiles * warps * (ctas > 148 ? 2 : 1) / c, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const int nrows_g = 4096;  // 0.9 MB per CTA region
    unsigned char *g;
    cudaMalloc(&g, (size_t)296 * nrows_g * ROWB);
    cudaMemset(g, 0, (size_t)296 * nrows_g * ROWB);
    long long *cyc;
    double *sink;
    cudaMalloc(&cyc, 8);
    cudaMalloc(&sink, 296 * 256 * 8);
    for (int ctas : {148, 296})
        for (int warps : {1, 2, 4, 8})
            for (int depth : {1, 2})
            {
                run<1>(g, nrows_g, warps, depth, ctas, cyc, sink);
                run<2>(g, nrows_g, warps, depth, ctas, cyc, sink);
                run<3>(g, nrows_g, warps, depth, ctas, cyc, sink);
                if (depth == 1) run<4>(g, nrows_g, warps, depth, ctas, cyc, sink);
                run<5>(g, nrows_g, warps, depth, ctas, cyc, sink);
                if (depth == 2) run<6>(g, nrows_g, warps, depth, ctas, cyc, sink);
            }
    return 0;
}
