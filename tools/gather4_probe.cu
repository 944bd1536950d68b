// Development micro-benchmark (sm_100a): the upwind-neighbour rows of a marching tile (3 faces x 8 elements = 24 rows of 28 doubles,
// scattered over an L2-resident flux block) gathered into shared memory
//   variant 0: cp.async 16 B, a lane owns one chunk column, two whole rows per instruction (what the neighbour producer does today:
//              12 LDGSTS warp instructions per tile, completion through cp.async.mbarrier.arrive.noinc)
//   variant 1: tensor-TMA  cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4  - four rows per instruction, six instructions per
//              tile issued by six lanes at once, completion by complete_tx on the tile's mbarrier; rows out of range are zero-filled
// W producer warps per CTA work on their own tile sequences with `depth` tiles in flight; C CTAs per SM.  Prints cycles per tile per
// warp and checks what arrived against the source rows.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/gather4_probe tools/gather4_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                                   \
    do                                                                                          \
    {                                                                                           \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess)                                                                  \
        {                                                                                       \
            std::printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
            std::exit(1);                                                                       \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int NP = 28, ROWB = NP * 8, NROW = 24, TILEB = ROWB * NROW, CH = ROWB / 16, RPI = 32 / CH;
constexpr int kMaxDepth = 4, kMaxWarps = 4;

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(s32(bar)), "r"(parity)
            : "memory");
}

template <int V>
__global__ void probe(const double *g, const CUtensorMap *tmap, int nrows_g, int tiles, int depth, long long *cyc, int *rows_out, double *tile_out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bars[kMaxWarps * kMaxDepth];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *my = sm + (size_t)warp * depth * TILEB;
    uint64_t *bar = bars + warp * kMaxDepth;
    if (lane == 0)
        for (int d = 0; d < depth; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar + d)), "r"(V == 0 ? 32 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t rng = 1234567u + blockIdx.x * 97u + warp * 13u;
    const int r_sub = lane / CH, lch = lane - r_sub * CH;
    const long long t0 = clock64();
    int row_last = 0;
    for (int t = 0; t < tiles; ++t)
    {
        const int d = t % depth;
        if (t >= depth) mbar_wait(bar + d, ((t / depth) - 1) & 1);  // the tile that used this buffer `depth` tiles ago has landed
        unsigned char *dst = my + (size_t)d * TILEB;
        // lane l < 24 holds the source row of tile row l (pseudo-random; every 16th is "no neighbour")
        rng = rng * 1664525u + 1013904223u + lane * 2654435761u;
        int row = (int)((rng >> 8) % (unsigned)nrows_g);
        if (((rng >> 4) & 15u) == 0u) row = -1;
        row_last = row;
        if (V == 0)
        {
            int cr[NROW / RPI];
#pragma unroll
            for (int k = 0; k < NROW / RPI; ++k) cr[k] = __shfl_sync(0xffffffffu, row, (k * RPI + r_sub) & 31);
#pragma unroll
            for (int k = 0; k < NROW / RPI; ++k)
                if (r_sub < RPI)
                {
                    const unsigned char *src = reinterpret_cast<const unsigned char *>(g) + (size_t)(cr[k] < 0 ? 0 : cr[k]) * ROWB + lch * 16;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s32(dst + (k * RPI + r_sub) * ROWB + lch * 16)), "l"(src),
                                 "r"(cr[k] >= 0 ? 16u : 0u)
                                 : "memory");
                }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(bar + d)) : "memory");
        }
        else
        {
            // lane j < 6 issues rows 4j .. 4j+3
            const int r0 = __shfl_sync(0xffffffffu, row, (4 * lane) & 31), r1 = __shfl_sync(0xffffffffu, row, (4 * lane + 1) & 31);
            const int r2 = __shfl_sync(0xffffffffu, row, (4 * lane + 2) & 31), r3 = __shfl_sync(0xffffffffu, row, (4 * lane + 3) & 31);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + d)), "r"((uint32_t)TILEB) : "memory");
            __syncwarp();
            if (lane < NROW / 4)
                asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                                 s32(dst + lane * 4 * ROWB)),
                             "l"(tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(bar + d))
                             : "memory");
        }
    }
    // drain
    for (int t = tiles; t < tiles + depth; ++t)
        if (t >= depth) mbar_wait(bar + t % depth, ((t / depth) - 1) & 1);
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * kMaxWarps + warp] = t1 - t0;
    if (blockIdx.x == 0 && warp == 0)
    {
        // the last tile of warp 0: rows and content for the host check
        if (lane < NROW) rows_out[lane] = row_last;
        const double *tl = reinterpret_cast<const double *>(my + (size_t)((tiles - 1) % depth) * TILEB);
        for (int i = lane; i < NROW * NP; i += 32) tile_out[i] = tl[i];
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int nrows = 4096 * 64;  // 58 MB: inside L2
    const int tiles = 2000;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double *g = nullptr;
    CK(cudaMalloc(&g, (size_t)nrows * ROWB));
    std::vector<double> h((size_t)nrows * NP);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)(i % 1000003) * 0.5 + 1.0;
    CK(cudaMemcpy(g, h.data(), h.size() * 8, cudaMemcpyHostToDevice));

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess)
    {
        std::printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    const int box_rows = argc > 1 ? std::atoi(argv[1]) : 1;  // rows of the box in the tensor map (the gather4 form wants to be told)
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)NP, (cuuint64_t)nrows}, strides[1] = {(cuuint64_t)ROWB};
    const cuuint32_t box[2] = {(cuuint32_t)NP, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    std::printf("cuTensorMapEncodeTiled (box %d x %d): %d\n", NP, box_rows, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    CUtensorMap *d_tm = nullptr;
    CK(cudaMalloc(&d_tm, sizeof(CUtensorMap)));
    CK(cudaMemcpy(d_tm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice));

    long long *d_cyc = nullptr;
    int *d_rows = nullptr;
    double *d_tile = nullptr;
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 2 * kMaxWarps));
    CK(cudaMalloc(&d_rows, sizeof(int) * NROW));
    CK(cudaMalloc(&d_tile, sizeof(double) * NROW * NP));
    for (int V = 0; V < 2; ++V)
        for (int ctas = 1; ctas <= 2; ++ctas)
            for (int warps = 1; warps <= 4; warps *= 2)
                for (int depth = 1; depth <= 4; depth *= 2)
                {
                    const size_t smem = (size_t)warps * depth * TILEB;
                    if (smem * ctas > 200 * 1024) continue;
                    auto fn = V == 0 ? probe<0> : probe<1>;
                    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    CK(cudaMemset(d_tile, 0, sizeof(double) * NROW * NP));
                    fn<<<sms * ctas, warps * 32, smem>>>(g, d_tm, nrows, tiles, depth, d_cyc, d_rows, d_tile);
                    CK(cudaGetLastError());
                    CK(cudaDeviceSynchronize());
                    std::vector<long long> c(sms * ctas * kMaxWarps);
                    CK(cudaMemcpy(c.data(), d_cyc, sizeof(long long) * c.size(), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (int b = 0; b < sms * ctas; ++b)
                        for (int w = 0; w < warps; ++w) mx = c[b * kMaxWarps + w] > mx ? c[b * kMaxWarps + w] : mx;
                    std::vector<int> rows(NROW);
                    std::vector<double> tl(NROW * NP);
                    CK(cudaMemcpy(rows.data(), d_rows, sizeof(int) * NROW, cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(tl.data(), d_tile, sizeof(double) * NROW * NP, cudaMemcpyDeviceToHost));
                    int bad = 0;
                    for (int i = 0; i < NROW; ++i)
                        for (int k = 0; k < NP; ++k)
                        {
                            const double want = rows[i] < 0 ? 0.0 : h[(size_t)rows[i] * NP + k];
                            bad += tl[i * NP + k] != want;
                        }
                    std::printf("variant %d CTAs/SM %d warps/CTA %d depth %d: %8.1f cycles per tile per warp = %6.1f per tile per SM, %s\n", V, ctas, warps, depth,
                                (double)mx / tiles, (double)mx / tiles / (warps * ctas), bad ? "CONTENT MISMATCH" : "content ok");
                }
    return 0;
}
