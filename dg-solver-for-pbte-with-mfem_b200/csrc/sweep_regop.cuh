// Register-operator marching kernel: the shared-operator sweep (sweep_shared.cuh) for meshes with ONE operator class
// (uniformly refined Cartesian quads / hexes: BASELINE configs C2, C4, C5).
//
// With one operator class a slot (direction, band) has 1 + nin operators in all (P and the Q_f of its nin <= 3 inflow faces),
// the same for every tile.  They leave shared memory: a consumer warp keeps the B fragments of ONE 8-column block of all
// operators in registers for the whole sweep ((1 + nin) * ceil(n/4) doubles per lane), and the MT = ceil(n/8) warps of a
// GROUP evaluate one tile together, one column block each, on the four FP64 tensor pipes of the SM at once:
//     U_tile(8 x n)[:, 8w .. 8w+7] = X(8 x n) P^T[:, 8w..] + sum_m Unbr_m(8 x n) Q_m^T[:, 8w..]        (warp w of the group)
// Against one warp per tile this divides the latency of a tile by MT, and the latency of a tile is what a wavefront sweep
// pays once per level (46 levels x ~6000 cycles were a quarter of the C2 sweep, profiles/README.md); shared memory only
// carries the element data (A fragments), and nothing is staged per CTA any more, so the ring is deeper.
//
// CTA = one slot.  Warps 0-3 (one warpgroup, registers released with setmaxnreg): 2 source producers + 2 upwind-neighbour
// producers with the roles of sweep_march_kernel.  Then G groups of MT consumer warps; group g takes tiles g, g + G, ...
// Finished tiles are published per consumer warp as a monotone counter (no atomics); the neighbour producers wait until every
// warp of every group has passed the tile's last prerequisite (descriptor header word 3).
#pragma once

#include "sweep_shared.cuh"

namespace pbte_dev
{
constexpr int kRegMaxIn = 3;  // inflow faces of a Cartesian element: dim
constexpr int kRegProducerWarps = 4;  // one warpgroup: 2 source + 2 upwind-neighbour loaders (tile p, p + 2, ... each); 4 + 4 measured no faster
constexpr int kRegRoleWarps = kRegProducerWarps / 2;

// groups per CTA: eight consumer warps where the column blocks allow it
__host__ __device__ constexpr int regop_groups(int n)
{
    const int mt = (n + 7) / 8;
    return mt >= 5 ? 1 : 8 / mt;
}
__host__ __device__ constexpr int regop_consumer_warps(int n) { return ((regop_groups(n) * ((n + 7) / 8) + 3) / 4) * 4; }  // whole warpgroups
__host__ __device__ constexpr int regop_threads(int n) { return (kRegProducerWarps + regop_consumer_warps(n)) * 32; }
// CTAs per SM the register file allows: (1 + nin) * KS operator doubles per consumer lane
__host__ __device__ constexpr int regop_ctas(int n) { return n <= 32 ? 2 : 1; }
__host__ __device__ constexpr int regop_consumer_regs(int n) { return n <= 32 ? 104 : 232; }
__host__ __device__ constexpr int regop_producer_regs(int n) { return n <= 32 ? 32 : 40; }
// register pool of a CTA: (launch registers) * threads >= producers * producer_regs + consumers * consumer_regs
//   n <= 32: 384 threads, 2 CTAs/SM -> 80 at launch: 128 * 32 + 256 * 104 = 30720;  n > 32: 384 threads, 1 CTA/SM -> 168: 128 * 40 + 256 * 232 = 64512
__host__ __device__ constexpr bool regop_supported(int n) { return n >= 4 && n <= 64; }

__host__ __device__ inline size_t regop_meta_bytes(int n, int ring, size_t h_doubles)
{
    return 64 * sizeof(int32_t) + sizeof(uint64_t) * (2 * (size_t)ring) + sizeof(double) * h_doubles;
}

template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}

// progress counters: release store by a consumer (after its tile is in global memory), acquire load by the neighbour producers
__device__ __forceinline__ void st_release_shared(int32_t *p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_shared(const int32_t *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// grid = nslots CTAs of regop_threads(N) threads.  Dynamic shared memory = [progress counters | barriers | wall vectors | tile ring].
template <int N>
__global__ void __launch_bounds__(regop_threads(N), regop_ctas(N)) sweep_march_regop_kernel(DevProblem P, DevMarch S)
{
    using FG = Frag<N>;
    constexpr int MT = FG::MT, KS = FG::KS, FR = FG::FR;
    constexpr int G = regop_groups(N);
    constexpr int CW = G * MT;                      // working consumer warps
    constexpr int NP = (N + 1) & ~1;                // global row stride (doubles): 16-byte aligned rows
    constexpr int NS = stage_stride(N);             // shared-memory row stride
    constexpr int CH = NP / 2;                      // 16-byte chunks per row
    static_assert(CW <= 32, "one lane of a neighbour producer polls one consumer warp");
    extern __shared__ __align__(128) unsigned char sm_raw[];
    if (*P.stop) return;  // the iteration before this one met the stop test (csrc/tail.cuh)
    const int s = blockIdx.x;
    const int c = S.slot_class[s];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = S.F, rows = 2 + F, R = S.ring, nv = S.nv;
    const int lb = S.cls_level_begin[c], nlev = S.cls_level_begin[c + 1] - lb - 1;
    const int tbase = S.level_tile_begin[lb], NT = S.level_tile_begin[lb + nlev] - tbase;
    const unsigned inm = S.cls_inmask[c];  // one operator class: [nclass][1]
    const int nin = __popc(inm);

    // ---- shared memory carve-up ---------------------------------------------------------------------------------
    int32_t *prog = reinterpret_cast<int32_t *>(sm_raw);           // [CW] tiles finished by consumer warp x: all of its tiles < prog[x]
    uint64_t *full = reinterpret_cast<uint64_t *>(sm_raw + 64 * sizeof(int32_t));  // [R]
    uint64_t *empty = full + R;                                                     // [R]
    const int Fp = (F + 1) & ~1;
    const size_t hdr_bytes = (size_t)rows * 32 + (size_t)kTile * Fp * sizeof(double);  // descriptor + wall temperatures
    const size_t slot_bytes = hdr_bytes + (size_t)nv * kTile * NS * sizeof(double);
    double *s_h = reinterpret_cast<double *>(empty + R);  // [F][N] wall vectors of this slot (when S.h_smem)
    const size_t h_one = (size_t)F * N;
    const size_t h_doubles = S.h_smem ? h_one : 0;
    unsigned char *ring = reinterpret_cast<unsigned char *>(s_h + h_doubles);
    // 16-byte alignment as an OFFSET from the shared-memory base (an integer round trip of the pointer would make every ring access a
    // generic load instead of LDS)
    ring = sm_raw + ((static_cast<size_t>(ring - sm_raw) + 15) & ~static_cast<size_t>(15));
    for (size_t i = threadIdx.x; i < h_doubles; i += blockDim.x) s_h[i] = S.oph[(size_t)s * h_one + i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) prog[i] = 0;
    // the ring starts zeroed: copies write NP doubles per row, the k padding up to NS stays zero for the whole kernel
    for (size_t i = threadIdx.x; i < (size_t)R * slot_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0, 0, 0, 0);
    // full: one source and one neighbour warp, every lane; empty: lane 0 of the MT warps of the group
    for (int i = threadIdx.x; i < 2 * R; i += blockDim.x) mbar_init(full + i, i < R ? 2 * 32 : MT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    double *u = P.flux + (size_t)s * P.ne * NP;
    const int32_t *desc0 = S.tile_desc + (size_t)tbase * rows * kTile;

    // Copy pattern of the producers: a lane owns one 16-byte chunk column, an instruction covers RPI = 32 / CH whole rows (rows are
    // contiguous, so the instruction still touches few cache lines: the memory-order pattern of sweep_march_kernel, minus its
    // per-chunk index arithmetic).  Per instruction a lane needs one shuffle (the row's element) and one multiply-add (the
    // address); the shared-memory offset is a per-lane constant plus an immediate, and the instructions of a tile are independent
    // of each other.  (A pattern that had to be decoded per chunk made a tile cost ~1900 cycles of a producer warp: profiles/.)
    constexpr int RPI = 32 / CH;
    static_assert(RPI >= 1, "a row is at most 32 chunks of 16 bytes");
    const int r_sub = lane / CH, ch = lane - r_sub * CH;
    const bool lane_copies = r_sub < RPI;
    const uint32_t lane_off = (uint32_t)(r_sub * NS * 8 + ch * 16);  // offset of the lane's chunk inside the staged vectors

    if (warp < kRegProducerWarps)
    {
        setmaxnreg_dec<regop_producer_regs(N)>();
        if (warp < kRegRoleWarps)
        {
            // ---- source producer: descriptor, wall temperatures, u_old / T_old rows; runs up to R tiles ahead ------------
            constexpr int ITS = (16 + RPI - 1) / RPI;
            const int q = lane & 7;
            const unsigned char *ub = reinterpret_cast<const unsigned char *>(u) + ch * 16, *tb = reinterpret_cast<const unsigned char *>(P.tc_prev) + ch * 16;
            int e_next = warp < NT ? __ldg(desc0 + (size_t)warp * rows * kTile + q) : -1;
            for (int i = warp; i < NT; i += kRegRoleWarps)
            {
                const int slot = i % R, it = i / R;
                const int e = e_next;
                const int32_t *desc_g = desc0 + (size_t)i * rows * kTile;
                if (i + kRegRoleWarps < NT) e_next = __ldg(desc_g + (size_t)kRegRoleWarps * rows * kTile + q);
                if (it > 0) mbar_wait(empty + slot, (it - 1) & 1);
                PBTE_STAMP(i, 7);
                unsigned char *sl = ring + (size_t)slot * slot_bytes;
                const uint32_t vb = smem_u32(sl + hdr_bytes) + lane_off;
#pragma unroll
                for (int k = 0; k < ITS; ++k)
                {
                    const int row = k * RPI + r_sub;  // rows 0-7 u_old, 8-15 T_old of the tile's elements
                    const int er = __shfl_sync(0xffffffffu, e, row & 7);
                    if (lane_copies && row < 16)
                    {
                        const unsigned char *src = (row < 8 ? ub : tb) + (size_t)(er < 0 ? 0 : er) * (NP * 8);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vb + (uint32_t)(k * RPI * NS * 8)), "l"(src), "r"(er >= 0 ? 16u : 0u) : "memory");
                    }
                }
                if (lane < rows * 2) cp_async16(sl + lane * 16, reinterpret_cast<const unsigned char *>(desc_g) + lane * 16, 16u);
                {
                    const int hc = Fp >> 1, row = lane / hc, ch = lane - row * hc;
                    const int er = __shfl_sync(0xffffffffu, e, row & 7);
                    if (lane < kTile * hc) cp_async16(sl + rows * 32 + lane * 16, S.ctbc + (size_t)(er < 0 ? 0 : er) * Fp + 2 * ch, er >= 0 ? 16u : 0u);
                }
                cp_async_arrive(full + slot);
                PBTE_STAMP(i, 8);
            }
        }
        else
        {
            // ---- upwind-neighbour producer: follows the wavefront -----------------------------------------------------
            constexpr int ITN = (kRegMaxIn * kTile + RPI - 1) / RPI;
            const int first = warp - kRegRoleWarps;
            const unsigned char *ub = reinterpret_cast<const unsigned char *>(u) + ch * 16;
            const int q = lane & 7, mq = lane >> 3;
            const int gx = lane / MT;  // group of the consumer warp this lane polls
            int L_next = first < NT ? __ldg(desc0 + ((size_t)first * rows + 1 + F) * kTile + 3) : -1;
            int code_next = first < NT ? __ldg(desc0 + ((size_t)first * rows + 1 + mq) * kTile + q) : -1;  // rows 1..4 always exist (F >= 3)
            for (int i = first; i < NT; i += kRegRoleWarps)
            {
                const int slot = i % R, it = i / R;
                const int dep = L_next, code = code_next;
                if (i + kRegRoleWarps < NT)
                {
                    const int32_t *dn = desc0 + (size_t)(i + kRegRoleWarps) * rows * kTile;
                    L_next = __ldg(dn + (1 + F) * kTile + 3);
                    code_next = __ldg(dn + (1 + mq) * kTile + q);
                }
                if (it > 0) mbar_wait(empty + slot, (it - 1) & 1);
                PBTE_STAMP(i, 4);
                // every tile up to the last prerequisite must be stored: group gx has passed it once its counter exceeds the last
                // tile of its own sequence (gx, gx + G, ...) that is <= dep
                {
                    const int need = (lane < CW && dep >= gx) ? dep - ((dep - gx) % G) + 1 : 0;
                    const int32_t *pl = prog + (lane < CW ? lane : 0);
                    while (!__all_sync(0xffffffffu, ld_acquire_shared(pl) >= need)) __nanosleep(20);  // acquire: the stores of those tiles
                }                                                                                     // are visible to the copies below
                PBTE_STAMP(i, 5);
                const uint32_t vb = smem_u32(ring + (size_t)slot * slot_bytes + hdr_bytes) + (uint32_t)(2 * kTile * NS * 8) + lane_off;
                const int nrow = nin * kTile;
#pragma unroll
                for (int k = 0; k < ITN; ++k)
                {
                    const int row = k * RPI + r_sub;  // lane `row` holds the source of row `row` (face row / 8, element row % 8)
                    const int cr = __shfl_sync(0xffffffffu, code, row & 31);
                    if (lane_copies && row < nrow)
                    {
                        const unsigned char *src = ub + (size_t)(cr < 0 ? 0 : cr) * (NP * 8);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vb + (uint32_t)(k * RPI * NS * 8)), "l"(src), "r"(cr >= 0 ? 16u : 0u) : "memory");
                    }
                }
                cp_async_arrive(full + slot);
                PBTE_STAMP(i, 6);
            }
        }
    }
    else
    {
        // ---- consumers ----------------------------------------------------------------------------------------------
        setmaxnreg_inc<regop_consumer_regs(N)>();
        const int cw = warp - kRegProducerWarps;
        if (cw >= CW) return;  // padding warps of the last warpgroup
        const int g = cw / MT, w = cw - g * MT;
        const int q = lane >> 2, r = lane & 3;
        const double c1 = P.slot_c1[s], c2 = P.slot_c2[s];
        // B fragments of column block w of P and of the Q_m: lane holds R[8w + q][4kk + r] (fragment order of store_fragment_order)
        double breg[1 + kRegMaxIn][KS];
        {
            const double *ops = S.ops + S.slot_op_base[s] + (size_t)(((w >> 1) * KS) * 32 + lane) * 2 + (w & 1);
#pragma unroll
            for (int b = 0; b <= kRegMaxIn; ++b)
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) breg[b][kk] = b <= nin ? __ldg(ops + (size_t)b * FR + (size_t)kk * 64) : 0.0;
        }
        // canonical face of the m-th inflow face and this lane's entries of its wall vector
        const double *oph = S.h_smem ? s_h : S.oph + (size_t)s * h_one;
        const int d0 = 8 * w + 2 * r;
        int jm[kRegMaxIn];
        {
            unsigned rest = inm;
#pragma unroll
            for (int m = 0; m < kRegMaxIn; ++m)
            {
                jm[m] = rest ? __ffs(rest) - 1 : 0;
                rest &= rest - 1;
            }
        }
        for (int i = g; i < NT; i += G)
        {
            const int slot = i % R, it = i / R;
            const unsigned char *sl = ring + (size_t)slot * slot_bytes;
            const int32_t *sd = reinterpret_cast<const int32_t *>(sl);
            const double *vec = reinterpret_cast<const double *>(sl + hdr_bytes) + (size_t)q * NS + r;
            if (w == 0) PBTE_STAMP(i, 0);
            mbar_wait(full + slot, it & 1);
            if (w == 0) PBTE_STAMP(i, 1);
            const int e = sd[q];
            int code[kRegMaxIn];
#pragma unroll
            for (int m = 0; m < kRegMaxIn; ++m) code[m] = m < nin ? sd[(1 + m) * kTile + q] : kSrcSkipCode;  // rows 1..3 always exist
            double a0x = 0.0, a0y = 0.0, a1x = 0.0, a1y = 0.0;  // two accumulators: consecutive MMAs never wait for each other
            {
                // source / pseudo-time term: X = c1 T_old + c2 u_old
                double xa[KS];
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) xa[kk] = fma(c2, vec[4 * kk], c1 * vec[(size_t)kTile * NS + 4 * kk]);
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                {
                    if (kk & 1) dmma_m8n8k4(a1x, a1y, xa[kk], breg[0][kk]);
                    else dmma_m8n8k4(a0x, a0y, xa[kk], breg[0][kk]);
                }
            }
#pragma unroll
            for (int m = 0; m < kRegMaxIn; ++m)
            {
                if (__ballot_sync(0xffffffffu, code[m] >= 0) == 0u) continue;  // no element of the tile has an interior neighbour there
                double xa[KS];
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) xa[kk] = vec[(size_t)(2 + m) * kTile * NS + 4 * kk];
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                {
                    if ((kk + (m + 1) * KS) & 1) dmma_m8n8k4(a1x, a1y, xa[kk], breg[1 + m][kk]);
                    else dmma_m8n8k4(a0x, a0y, xa[kk], breg[1 + m][kk]);
                }
            }
            double ux = a0x + a1x, uy = a0y + a1y;
#ifdef PBTE_TRACE
            if (ux == 1.2345e300) PBTE_STAMP(i, 15);  // the stamp below waits for the accumulators
#endif
            if (w == 0) PBTE_STAMP(i, 2);
            // isothermal walls on inflow faces: + T_bc h_j   (this lane owns element q, dofs d0, d0 + 1)
            {
                const double *tbq = reinterpret_cast<const double *>(sl + rows * 32) + (size_t)q * Fp;
#pragma unroll
                for (int m = 0; m < kRegMaxIn; ++m)
                    if (code[m] == kSrcWallCode)
                    {
                        const double t = tbq[jm[m]];
                        const double *hj = oph + (size_t)jm[m] * N;
                        if (d0 < N) ux = fma(t, hj[d0], ux);
                        if (d0 + 1 < N) uy = fma(t, hj[d0 + 1], uy);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot);  // this warp has consumed every staged value of the tile
            if (e >= 0 && d0 < N)
            {
                double *dst = u + (size_t)e * NP + d0;
                if (d0 + 1 < N) *reinterpret_cast<double2 *>(dst) = make_double2(ux, uy);
                else *dst = ux;
            }
            __syncwarp();  // orders the stores of all lanes before lane 0's release
            if (lane == 0) st_release_shared(prog + cw, i + 1);
            if (w == 0) PBTE_STAMP(i, 3);
        }
    }
}

}  // namespace pbte_dev
