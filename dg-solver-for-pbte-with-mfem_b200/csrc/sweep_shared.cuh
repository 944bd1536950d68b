// Shared-operator sweep ("marching" kernel): the fast path for meshes whose elements fall into few operator classes
// (csrc/shared_plan.hpp).  For an operator class o, direction k and band b the local system matrix A is the same for every
// element of the class, so the sweep update is applied through precomputed solution operators
//     P   = A^-1 M^T                         u_e  = P (c1 T_old + c2 u_old)
//     Q_f = -c_f A^-1 K_f                         + sum_{inflow f} Q_f u_nbr(f)
//     h_f = -c_f (C/omega) A^-1 g_f               + sum_{inflow wall f} T_bc(f) h_f
// built once per (class, slot) by build_operators_kernel (pivoted LU, same assembly as the per-element kernels).
//
// sweep_march_kernel: one CTA per slot (direction, band).  The CTA stages the slot's operators in shared memory with
// one TMA bulk copy (or, for meshes with many operator classes, reads them compactly stored through L1), then marches through
// the tile sequence of its upwind DAG: producer warps gather a tile's element data into a shared-memory ring with cp.async,
// a consumer warp takes one tile of 8 elements of one operator class and evaluates
//     U_tile(8 x n) = X(8 x n) P^T + sum_f Unbr_f(8 x n) Q_f^T
// as FP64 tensor-core MMAs (mma.sync m8n8k4: A fragment = element data from the ring, B fragment = operator, accumulators in
// registers), adds the wall terms, writes the tile back in place and publishes it.  A tile starts once every tile up to its last
// prerequisite is published (no level barrier); no inter-CTA dependency exists, one launch per sweep.
// sweep_regop.cuh holds the variant that keeps the operators in registers (one operator class, large elements).
// iteration_tail_kernel (tail.cuh) then forms [Tc | Qc] by one streaming pass over the flux and finishes the iteration.
#pragma once

#include "common.cuh"

#include <type_traits>

namespace pbte_dev
{
constexpr int kTile = 8;             // elements per tile (must equal pbte_host::kTileElems)
constexpr int kSrcSkipCode = -1;     // pbte_host::kSrcSkip
constexpr int kSrcWallCode = -2;     // pbte_host::kSrcWall

struct DevMarch
{
    int noclass, F;
    int h_smem;    // 1: the slot's wall vectors h_f are staged in shared memory (they fit), 0: read through L1
    int ring, nv;  // tile-ring slots per CTA; staged vectors per slot = 2 (u_old, T_old) + max inflow faces of any class
    const int32_t *slot_class;        // [nslots] upwind-DAG class of the slot
    const int32_t *cls_level_begin;   // [nclass+1] into level_tile_begin
    const int32_t *level_tile_begin;  // per class nlevels+1 tile offsets
    const int32_t *tile_oclass;       // [ntiles]
    const int32_t *tile_desc;         // [ntiles][1+F][8]
    const uint8_t *cls_inmask;        // [nclass][noclass] canonical inflow faces
    int op_stride;                    // doubles per operator in ops (by layout)
    int op_layout;                    // 0: MMA fragment pairs (frag_doubles: register-operator kernel), 1: compact row-major n x n
                                      // (operators that are never staged: many operator classes), 2: split order (split_doubles: staged)
    int max_op_cnt, max_class_tiles;  // most operators per slot / most tiles of any upwind-DAG class (shared-memory carve-up)
    int ncta;                         // persistent CTAs of sweep_march_kernel
    const int32_t *cta_piece_begin;   // [ncta+1] into pieces
    const int32_t *pieces;            // [npieces][4] slot, first tile, end tile, tiles of the slot (sweep_march_kernel: work list)
    int32_t *progress;                // [nslots] tiles of the slot stored so far by the CTA that holds its first part (0 between sweeps)
    const int32_t *cls_op_off;        // [nclass][noclass] first operator of the class inside a slot's operator block (units of op_stride)
    const int32_t *cls_op_cnt;        // [nclass] operators per slot
    const unsigned long long *slot_op_base;  // [nslots] offset (doubles) of the slot's operator block in ops
    double *ops;                      // operators in MMA fragment order
    double *oph;                      // [nslots][noclass][F][n] wall vectors h_f
    const double *ctbc;               // [ne][Fp] wall temperature per canonical face, Fp = F rounded up to even (16-byte rows)
    // operator construction: the local matrices of the class representatives only (the marching path never needs the
    // per-element matrix image; with pbte_options_t-independent PBTE_IMAGE=reps it is not even uploaded)
    const double *rep_mass, *rep_massT;  // [noclass][ms]
    const double *rep_stiff;             // [noclass][dim][ms]
    const double *rep_fmass;             // [noclass][F][ms]   by LOCAL face of the representative
    const double *rep_fcoup;             // [noclass][F][ms]   by CANONICAL face: K of an element of the class that has an interior face there
    const double *rep_fint;              // [noclass][F][n]    by canonical face: g of an element of the class that has a wall there
    const double *rep_normal;            // [noclass][F][3]    by local face
    const uint8_t *rep_has_k, *rep_has_g;  // [noclass][F] the class has an interior / a wall face at canonical face j
    const uint8_t *rep_cface;            // [noclass][F] canonical -> local face of the representative
    const uint8_t *rep_present;          // [noclass][F] canonical face j exists
    long long *trace;                 // development builds (-DPBTE_TRACE) only: per-tile clock stamps of CTA 0, else null
};

// fragment geometry for ndof = N
template <int N>
struct Frag
{
    static constexpr int MT = (N + 7) / 8;   // 8-row blocks of the operator
    static constexpr int MP = (MT + 1) / 2;  // pairs of row blocks: one 16-byte shared-memory load feeds two MMAs
    static constexpr int KS = (N + 3) / 4;   // k steps of 4
    static constexpr int FR = MP * KS * 64;  // doubles per operator in fragment order
};
__host__ __device__ inline int frag_doubles(int n) { return ((((n + 7) / 8) + 1) / 2) * ((n + 3) / 4) * 64; }
// Compact storage, row-major n x n rounded to an even count: used when the operators are read from global memory by every tile
// (meshes with many operator classes).  Fragment order pads ndof 10 to 16 x 12 (192 doubles against 100); there the operator
// fragments are most of what the kernel reads from DRAM (ncu on C3: 777 GB per sweep against 80 GB of flux), so the padding
// goes.
__host__ __device__ constexpr int compact_doubles(int n) { return (n * n + 1) & ~1; }

// Largest remainder that goes to DFMAs.  Measured (profiles/r02_kernel_probes.txt): one remainder column (ndof 25) gains 4 % on the C5
// mesh; three (ndof 27: two more shared-memory loads and three DFMAs per k step against one MMA saved) lose 2.5 % on C2, although the
// FP64 pipe has 16 % less to do: the consumers are not bound by that pipe alone.
#ifndef PBTE_REMSPLIT_MAX
#define PBTE_REMSPLIT_MAX 2
#endif
constexpr int kRemSplitMax = PBTE_REMSPLIT_MAX;

// Column split of an operator application U(8 x n) = X(8 x n) R^T: the MMA shape m8n8k4 works on 8 result columns at a time.  When
// n is not a multiple of 8 the last column block is mostly padding (ndof 27: 3 of 8 columns, ndof 25: 1 of 8), and on sm_100 an MMA
// block costs the FP64 pipe exactly what 8 x 8 x 4 scalar multiply-adds cost (DMMA and DFMA share one pipe at one rate:
// profiles/r02_fp64_mix_probe.txt).  So up to four remainder columns are evaluated with plain DFMAs instead: the lane that holds the
// A-fragment value X[q][4kk + r] multiplies it with R[8 MTM + c][4kk + r], c < REM, and the four lanes of a row add their partial
// sums with two shuffles at the end of the tile.  ndof 27: 81 MMAs + 81 DFMAs per full tile instead of 108 MMAs (-16 % pipe time).
template <int N, bool STAGE = true>
struct Split
{
    static constexpr int REM0 = N % 8;
    // staged operators only: read from global memory (compact layout, many operator classes) the two extra loads per k step cost
    // more than the MMA they save (C3 slice: 3.6 -> 4.8 ms)
    static constexpr bool kSplit = STAGE && N >= 8 && REM0 >= 1 && REM0 <= kRemSplitMax;
    static constexpr int MTM = kSplit ? N / 8 : (N + 7) / 8;  // column blocks evaluated by MMAs
    static constexpr int REM = kSplit ? REM0 : 0;             // remainder columns evaluated by DFMAs
    static constexpr int REMP = (REM + 1) & ~1;               // ... padded to whole 16-byte words
    static constexpr int P2 = MTM / 2;                        // pairs of column blocks (one 16-byte fragment load feeds two MMAs)
    static constexpr int KS = (N + 3) / 4;                    // k steps of 4
    static constexpr int MAIN = KS * 32 * MTM;                // doubles of the MMA part of a staged operator
    static constexpr int OPS = MAIN + KS * 4 * REMP;          // doubles per staged operator (even)
};
// Staged-operator layout ("split order") of R (n x n):
//   MMA part:  for k step kk: for pair p < P2: lane `lane` finds R[16p + lane/4][4kk + lane%4] and R[16p + 8 + lane/4][..] in one 16-byte
//              word at (kk*MTM*32 + p*64 + lane*2); an odd last block 2*P2 sits at (kk*MTM*32 + P2*64 + lane)
//   DFMA part: at MAIN + (kk*4 + r)*REMP + c: R[8 MTM + c][4kk + r]            (zero padded in k and c)
__host__ __device__ inline int split_doubles(int n)
{
    const int rem0 = n % 8;
    const bool split = n >= 8 && rem0 >= 1 && rem0 <= kRemSplitMax;
    const int mtm = split ? n / 8 : (n + 7) / 8, remp = split ? ((rem0 + 1) & ~1) : 0, ks = (n + 3) / 4;
    return ks * 32 * mtm + ks * 4 * remp;
}
__device__ __forceinline__ void store_split_order(double *out, const double *R, int n)
{
    const int rem0 = n % 8;
    const bool split = n >= 8 && rem0 >= 1 && rem0 <= kRemSplitMax;
    const int MTM = split ? n / 8 : (n + 7) / 8, REM = split ? rem0 : 0, REMP = (REM + 1) & ~1, P2 = MTM / 2, KS = (n + 3) / 4;
    const int MAIN = KS * 32 * MTM, OPS = MAIN + KS * 4 * REMP;
    for (int idx = threadIdx.x; idx < OPS; idx += blockDim.x)
    {
        int row, col;
        if (idx < MAIN)
        {
            const int kk = idx / (32 * MTM), rest = idx - kk * 32 * MTM;
            int I, lane;
            if (rest < P2 * 64)
            {
                lane = (rest >> 1) & 31;
                I = 2 * (rest >> 6) + (rest & 1);
            }
            else
            {
                lane = rest - P2 * 64;
                I = 2 * P2;
            }
            row = 8 * I + (lane >> 2);
            col = 4 * kk + (lane & 3);
        }
        else
        {
            const int remp = REMP > 0 ? REMP : 1, rest = idx - MAIN, kr = rest / remp, c = rest - kr * remp;
            row = c < REM ? 8 * MTM + c : n;
            col = kr;  // = 4 kk + r
        }
        out[idx] = (row < n && col < n) ? R[row + (size_t)col * n] : 0.0;
    }
}

// thread j solves column j of R in place with the row-pivoted LU factors in A (swaps, unit-lower forward, upper backward)
__device__ __forceinline__ void lu_solve_columns(const double *A, double *R, const int *piv, int n, int ncol)
{
    for (int j = threadIdx.x; j < ncol; j += blockDim.x)
    {
        double *x = R + (size_t)j * n;
        for (int k = 0; k < n; ++k)
        {
            const int pv = piv[k];
            if (pv != k)
            {
                const double tmp = x[k];
                x[k] = x[pv];
                x[pv] = tmp;
            }
        }
        for (int cc = 0; cc < n; ++cc)
        {
            const double xc = x[cc];
            for (int i = cc + 1; i < n; ++i) x[i] -= A[i + (size_t)cc * n] * xc;
        }
        for (int cc = n - 1; cc >= 0; --cc)
        {
            const double xc = (x[cc] /= A[cc + (size_t)cc * n]);
            for (int i = 0; i < cc; ++i) x[i] -= A[i + (size_t)cc * n] * xc;
        }
    }
}

// Fragment order of an operator R (n x n, column-major): the B fragment of mma m8n8k4 for row block I and k block kk holds
// R[8I + lane/4][4kk + lane%4] in lane `lane`; row blocks are stored in pairs so that lane `lane` finds the fragments of
// blocks 2Ip and 2Ip+1 in one 16-byte word:  out[((Ip*KS + kk)*32 + lane)*2 + (I & 1)], zero padded.
__device__ __forceinline__ void store_fragment_order(double *out, const double *R, int n)
{
    const int MT = (n + 7) / 8, KS = (n + 3) / 4, FR = frag_doubles(n);
    for (int idx = threadIdx.x; idx < FR; idx += blockDim.x)
    {
        const int half = idx & 1, lane = (idx >> 1) & 31, blk = idx >> 6, kk = blk % KS, I = 2 * (blk / KS) + half;
        const int row = 8 * I + (lane >> 2), col = 4 * kk + (lane & 3);
        out[idx] = (I < MT && row < n && col < n) ? R[row + (size_t)col * n] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One CTA per (operator class, slot): assemble A of the class representative for the slot's direction and band, factor it
// with row pivoting, solve for M^T, the K_f of the inflow faces and their g_f, store the results in fragment order.
// grid (nslots, noclass); dynamic shared memory 2*n*n doubles.
__global__ void __launch_bounds__(128) build_operators_kernel(DevProblem P, DevMarch S)
{
    extern __shared__ double sm_ops[];
    const int n = P.n, F = P.F, dim = P.dim, T = blockDim.x, t = threadIdx.x;
    double *A = sm_ops;                 // [n*n] LU factors
    double *R = A + (size_t)n * n;      // [n*n] right-hand sides / solutions
    __shared__ double s_af[kMaxFaces];
    __shared__ int s_piv[128];
    __shared__ int s_fail;
    const int s = blockIdx.x, o = blockIdx.y;
    const int c = S.slot_class[s];
    const unsigned inm = S.cls_inmask[(size_t)c * S.noclass + o];
    const size_t ms = P.ms;
    const double om[3] = {P.slot_dir[3 * s], P.slot_dir[3 * s + 1], P.slot_dir[3 * s + 2]};
    const double v = P.slot_v[s], cb = P.slot_cb[s];
    const int FR = frag_doubles(n);
    if (t < kMaxFaces)
    {
        // a_j of canonical face j (local face cface[e][j])
        double a = 0.0;
        if (t < F)
        {
            const int f = S.rep_cface[(size_t)o * F + t];
            if (S.rep_present[(size_t)o * F + t]) a = dot_seq(S.rep_normal + ((size_t)o * F + f) * 3, om, dim);
        }
        s_af[t] = a;
    }
    if (t == 0) s_fail = 0;
    __syncthreads();

    const double *Me = S.rep_mass + (size_t)o * ms, *Ge = S.rep_stiff + (size_t)o * dim * ms, *Fe = S.rep_fmass + (size_t)o * F * ms;
    for (int idx = t; idx < n * n; idx += T)
    {
        double b = 0.0;
        for (int d = 0; d < dim; ++d) b = fma(-om[d], Ge[(size_t)d * ms + idx], b);
        for (int j = 0; j < F; ++j)
            if (s_af[j] > 0.0) b = fma(s_af[j], Fe[(size_t)S.rep_cface[(size_t)o * F + j] * ms + idx], b);
        A[idx] = fma(v, b, P.nu_max * Me[idx]);
    }
    __syncthreads();
    // LU with row pivoting (first maximum, as mfem::LUFactors); rows are swapped in place, swaps recorded
    for (int k = 0; k < n; ++k)
    {
        if (t == 0)
        {
            int piv = k;
            double best = fabs(A[k + (size_t)k * n]);
            for (int i = k + 1; i < n; ++i)
            {
                const double x = fabs(A[i + (size_t)k * n]);
                if (x > best)
                {
                    best = x;
                    piv = i;
                }
            }
            s_piv[k] = piv;
            if (best == 0.0) s_fail = 1;
        }
        __syncthreads();
        const int piv = s_piv[k];
        if (piv != k)
            for (int j = t; j < n; j += T)
            {
                const double tmp = A[k + (size_t)j * n];
                A[k + (size_t)j * n] = A[piv + (size_t)j * n];
                A[piv + (size_t)j * n] = tmp;
            }
        __syncthreads();
        const double inv = 1.0 / A[k + (size_t)k * n];
        __syncthreads();
        for (int i = k + 1 + t; i < n; i += T) A[i + (size_t)k * n] *= inv;
        __syncthreads();
        const int w = n - k - 1;
        for (int idx = t; idx < w * w; idx += T)
        {
            const int i = k + 1 + idx % w, j = k + 1 + idx / w;
            A[i + (size_t)j * n] -= A[i + (size_t)k * n] * A[k + (size_t)j * n];
        }
        __syncthreads();
    }

    const int OPS = S.op_stride;
    double *ops = S.ops + S.slot_op_base[s] + (size_t)S.cls_op_off[(size_t)c * S.noclass + o] * OPS;
    auto store_operator = [&](double *out, const double *Rm) {
        if (S.op_layout == 1)
            for (int idx = t; idx < OPS; idx += T) out[idx] = idx < n * n ? Rm[idx / n + (size_t)(idx % n) * n] : 0.0;  // row-major
        else if (S.op_layout == 2) store_split_order(out, Rm, n);
        else store_fragment_order(out, Rm, n);
    };
    // P = A^-1 M^T
    {
        const double *MTe = S.rep_massT + (size_t)o * ms;
        for (int idx = t; idx < n * n; idx += T) R[idx] = MTe[idx];
        __syncthreads();
        lu_solve_columns(A, R, s_piv, n, n);
        __syncthreads();
        store_operator(ops, R);
        __syncthreads();
    }
    int m = 1;
    for (int j = 0; j < F; ++j)
    {
        if (!(inm >> j & 1u)) continue;
        const double cin = 0.5 * v * (s_af[j] - fabs(s_af[j]));  // c_f (src/PBTESolver.cpp:282)
        const int ks = S.rep_has_k[(size_t)o * F + j] ? o * F + j : -1, gs = S.rep_has_g[(size_t)o * F + j] ? o * F + j : -1;
        double *out = ops + (size_t)m * OPS;
        if (ks >= 0)
        {
            // Q_j = -c_j A^-1 K_j
            const double *K = S.rep_fcoup + (size_t)ks * ms;
            for (int idx = t; idx < n * n; idx += T) R[idx] = -cin * K[idx];
            __syncthreads();
            lu_solve_columns(A, R, s_piv, n, n);
            __syncthreads();
            store_operator(out, R);
        }
        else
            for (int idx = t; idx < OPS; idx += T) out[idx] = 0.0;  // no element of the class has an interior face j
        __syncthreads();
        if (gs >= 0)
        {
            // h_j = -c_j (C/omega) A^-1 g_j
            for (int i = t; i < n; i += T) R[i] = -cin * cb * S.rep_fint[(size_t)gs * n + i];
            __syncthreads();
            lu_solve_columns(A, R, s_piv, n, 1);
            __syncthreads();
            double *oh = S.oph + (((size_t)s * S.noclass + o) * F + j) * n;
            for (int i = t; i < n; i += T) oh[i] = R[i];
            __syncthreads();
        }
        ++m;
    }
    if (t == 0 && s_fail) atomicCAS(P.error_flag, 0, PBTE_ERR_SINGULAR);
}

// ---------------------------------------------------------------------------------------------------------------
// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// d0,d1 = D[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------------------
// mbarrier / asynchronous-copy primitives of the marching kernel
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for a barrier phase.  A failed try_wait comes back after a short, implementation-defined time; polling it in a tight loop
// made waiting warps the main source of issued instructions (ncu source page of the C2 kernel: 29 M polls for 136 k tiles).
// PBTE_WAIT_HINT_NS > 0 passes a suspend-time hint to try_wait, PBTE_WAIT_SLEEP_NS > 0 sleeps between polls.
#ifndef PBTE_WAIT_HINT_NS
#define PBTE_WAIT_HINT_NS 0
#endif
#ifndef PBTE_WAIT_SLEEP_NS
#define PBTE_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
    {
#if PBTE_WAIT_HINT_NS > 0
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)PBTE_WAIT_HINT_NS)
            : "memory");
#else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
#endif
#if PBTE_WAIT_SLEEP_NS > 0
        if (!ok) __nanosleep(PBTE_WAIT_SLEEP_NS);
#endif
    }
}
// one non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// 16-byte asynchronous copy global -> shared through L2 (src_bytes = 0 writes zeros)
__device__ __forceinline__ void cp_async16(void *dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// L2 eviction hints (experiment, PBTE_L2_HINTS): the u_old rows of a sweep are read once (evict first), the rows a sweep stores are
// read again by the downwind tiles, by the iteration tail and by the next sweep (evict last)
#ifndef PBTE_L2_HINTS
#define PBTE_L2_HINTS 0  // bit 0: evict-first loads of u_old, bit 1: evict-last stores of the new flux
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_global_f64x2_hint(double *dst, double2 v, uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(dst), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
// the barrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Development aid (-DPBTE_TRACE, never in the shipped library): CTA 0 records clock64 stamps per tile, 16 columns per tile:
// consumer 0 wait start, 1 full, 9 descriptor read, 2 MMAs done, 10 stored, 3 published; neighbour producer 4 slot free, 5 prerequisites
// stored, 6 copies issued; source producer 7 slot free, 8 copies issued.
#ifdef PBTE_TRACE
#define PBTE_STAMP(tile, col)                                                                              \
    do                                                                                                     \
    {                                                                                                      \
        if (S.trace && blockIdx.x == 0 && lane == 0) S.trace[(size_t)(tile) * 16 + (col)] = clock64();     \
    } while (0)
#else
#define PBTE_STAMP(tile, col) \
    do                        \
    {                         \
    } while (0)
#endif

constexpr int kMaxIn = 4;          // inflow faces whose wall temperature a consumer lane prefetches (tri 2, quad/hex dim, tet 3)

// MMA (consumer) warps of a marching CTA: two per SM sub-partition, so that one covers the other's descriptor / wall / store
// phases; one per sub-partition for large elements (registers, and the ring must hold at least one tile per consumer)
#ifndef PBTE_SMALL_CONSUMERS
#define PBTE_SMALL_CONSUMERS 8
#endif
#ifndef PBTE_SMALL_CTAS
#define PBTE_SMALL_CTAS 2
#endif
#ifndef PBTE_MID_CONSUMERS
#define PBTE_MID_CONSUMERS 8
#endif
#ifndef PBTE_MID_CTAS
#define PBTE_MID_CTAS 2
#endif
#ifndef PBTE_MID_PRODUCERS
#define PBTE_MID_PRODUCERS 2
#endif
__host__ __device__ constexpr int march_consumers(int n) { return n <= 16 ? PBTE_SMALL_CONSUMERS : (n <= 36 ? PBTE_MID_CONSUMERS : 4); }
__host__ __device__ constexpr int march_ctas_per_sm(int n) { return n <= 16 ? PBTE_SMALL_CTAS : (n <= 36 ? PBTE_MID_CTAS : 1); }
// Consumer warps that share one tile (compile-time experiment, off): with three or four MMA column blocks (ndof 25..36) two warps
// split the columns — warp 0 the blocks 0 and 1, warp 1 the rest and the DFMA remainder — so a tile leaves its ring slot twice as
// fast and the eight slots of a CTA hold four tiles in the MMAs and four on their way in.  Measured on C2 (profiles/r02_kernel_probes.txt):
// 0.447 - 0.469 ms against 0.40 ms with one warp per tile: both warps read the descriptor and every A fragment, and that costs more
// than the deeper prefetch returns.
#ifndef PBTE_WPT
#define PBTE_WPT 1
#endif
__host__ __device__ constexpr int march_wpt(int n)
{
    const int rem0 = n % 8;
    const int mtm = (n >= 8 && rem0 >= 1 && rem0 <= kRemSplitMax) ? n / 8 : (n + 7) / 8;
    return (PBTE_WPT == 2 && (mtm == 3 || mtm == 4) && n <= 36) ? 2 : 1;
}
// ring slots a CTA needs at least: one per tile a consumer warp can hold (an mbarrier parity wait is only meaningful within one
// phase of its barrier: a consumer must find the slot of its next tile at most one phase behind)
__host__ __device__ constexpr int march_min_ring(int n) { return march_consumers(n) / march_wpt(n); }
// Producer warps per role (source loaders run ahead of the wavefront, upwind-neighbour loaders follow it); the warp that takes tile
// g (counted over the CTA's whole work list) is g mod warps.  A producer warp needs ~1000 cycles per tile whatever the tile's size
// (the copies of a tile are issued and tracked at memory latency, profiles/README.md), a consumer warp needs ~50 cycles per MMA:
// small elements (tets order 2: 24 MMAs per tile) are bound by the producers and get four warps per role, large ones two.
#ifndef PBTE_SMALL_PRODUCERS
#define PBTE_SMALL_PRODUCERS 4
#endif
__host__ __device__ constexpr int march_role_warps(int n) { return n <= 16 ? PBTE_SMALL_PRODUCERS : (n <= 36 ? PBTE_MID_PRODUCERS : 2); }
// source / neighbour loader warps (the same number of each unless an experiment build says otherwise)
#ifndef PBTE_MID_NBR
#define PBTE_MID_NBR PBTE_MID_PRODUCERS
#endif
__host__ __device__ constexpr int march_src_warps(int n) { return march_role_warps(n); }
__host__ __device__ constexpr int march_nbr_warps(int n) { return (n > 16 && n <= 36) ? PBTE_MID_NBR : march_role_warps(n); }
__host__ __device__ constexpr int march_threads(int n) { return (march_consumers(n) + march_src_warps(n) + march_nbr_warps(n)) * 32; }

// shared-memory row stride (doubles) of a staged element vector: 4*KS rounded so that the four element rows a half warp reads
// with one LDS.64 fall into disjoint bank groups (stride mod 16 in {4, 12})
__host__ __device__ constexpr int stage_stride(int n)
{
    const int ns = ((n + 3) / 4) * 4;
    return (ns % 16 == 0 || ns % 16 == 8) ? ns + 4 : ns;
}
// bytes of one ring slot: tile descriptor (rows x 8 int32) + wall temperatures (8 x Fp) + nv staged vectors of 8 elements
__host__ __device__ inline size_t march_slot_bytes(int n, int rows, int nv, int F)
{
    return (size_t)rows * 32 + (size_t)kTile * ((F + 1) & ~1) * sizeof(double) + (size_t)nv * kTile * stage_stride(n) * sizeof(double);
}
// ---------------------------------------------------------------------------------------------------------------
// REGISTER-STAGED PRODUCERS (round 2, late).  With a ring slot per consumer warp a freed slot is the start of a latency chain:
// cp.async issued after the slot frees lands 700 - 1500 cycles later, and the ncu source page of the C2 kernel showed the consumers
// waiting for a full slot 36 % of their time.  Rows of at most 16 chunks (ndof <= 32) are therefore loaded into REGISTERS while the
// producer still waits for the slot (the loads need no slot, only their stores do); when the slot frees, the warp stores them to
// shared memory and arrives: the chain shrinks to a few store instructions.  The source producer forms X = c1 T_old + c2 u_old on
// the way (same expression, same rounding as the consumers used to evaluate per A fragment), so T_old needs no staged vector (one
// vector less per slot: the C2 ring grows from 8 to 10 tiles) and the first matrix block of a tile is a plain block like the others.
#ifndef PBTE_XSRC
#define PBTE_XSRC 0  // measured (profiles/r02_kernel_probes.txt, runs q and r): a source warp then has ONE tile in flight instead of a ring of
                     // them and pays the HBM latency per tile (C2 0.406 against 0.367 ms, C5 slice 2.78 against 1.98 ms); with the neighbour
                     // rows in registers as well ptxas funnels the twelve 16-byte loads of a hex tile through one register quad (0.75 - 0.92 ms)
#endif
__host__ __device__ constexpr bool march_xsrc(int n) { return PBTE_XSRC != 0 && ((n + 1) & ~1) <= 32; }
#ifndef PBTE_XNBR
#define PBTE_XNBR 1  // 0: only the source producer goes through registers, the neighbour rows stay cp.async (issued when the slot frees)
#endif
__host__ __device__ constexpr bool march_xnbr(int n) { return march_xsrc(n) && PBTE_XNBR != 0; }
constexpr int kRegFaces = 3;  // inflow faces whose rows one register-staged pass of the neighbour producer holds (tri/quad 2, tet/hex 3)

// 16-byte load through L2 (rows written by other warps / CTAs during this sweep must not come from L1); predicated INSIDE the asm
// so that every load keeps its own destination registers (a select around the asm made ptxas funnel all loads of a tile through one
// register pair, one L2 round trip after the other); `ok` false leaves zeros
__device__ __forceinline__ double2 ld_cg_f64x2_if(const void *p, bool ok)
{
    double2 v = make_double2(0.0, 0.0);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %3, 0;\n"
        "@p ld.global.cg.v2.f64 {%0, %1}, [%2];\n"
        "}\n"
        : "+d"(v.x), "+d"(v.y)
        : "l"(p), "r"((int)ok)
        : "memory");
    return v;
}
// the same with two 8-byte loads: 12 in-flight 16-byte loads per lane (the neighbour rows of a hex tile) need 12 aligned register
// quads, which ptxas only finds by funnelling every load through one temporary quad - pairs it does find
__device__ __forceinline__ double2 ld_cg_f64_f64_if(const void *p, bool ok)
{
    double2 v = make_double2(0.0, 0.0);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %3, 0;\n"
        "@p ld.global.cg.f64 %0, [%2];\n"
        "@p ld.global.cg.f64 %1, [%2+8];\n"
        "}\n"
        : "+d"(v.x), "+d"(v.y)
        : "l"(p), "r"((int)ok)
        : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_f64x2(uint32_t addr, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_shared_u32x4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Source producer, one tile: lane q = lane & 7 holds element q of the tile in `e` (-1: padding).  u/tc: slot's flux and T_old, row
// stride NP.  sl: the ring slot; the slot must be free before the stores (wait_empty: parity wait on empty_bar).
template <int N>
__device__ __forceinline__ void src_tile_regs(const double *u, const double *tc, int e, const int32_t *desc_g, const double *ctbc, int rows, int Fp,
                                              double c1, double c2, unsigned char *sl, size_t hdr_bytes, uint64_t *empty_bar, bool wait_empty,
                                              uint32_t parity, uint64_t *full_bar)
{
    constexpr int NP = (N + 1) & ~1, NS = stage_stride(N), CH = NP / 2, RPI = 32 / CH, IT = (kTile + RPI - 1) / RPI;
    const int lane = threadIdx.x & 31, r_sub = lane / CH, lch = lane - r_sub * CH;
    const bool lane_copies = r_sub < RPI;
    double2 ru[IT], rt[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k)
    {
        const int row = k * RPI + r_sub;
        const int er = __shfl_sync(0xffffffffu, e, row & 7);
        const bool ok = lane_copies && row < kTile && er >= 0;
        const size_t off = (size_t)(er < 0 ? 0 : er) * NP + 2 * lch;
        ru[k] = ld_cg_f64x2_if(u + off, ok);
        rt[k] = ok ? __ldg(reinterpret_cast<const double2 *>(tc + off)) : make_double2(0.0, 0.0);
    }
    uint4 dsc = make_uint4(0, 0, 0, 0);
    if (lane < rows * 2) dsc = __ldg(reinterpret_cast<const uint4 *>(desc_g) + lane);
    // wall temperatures of the 8 elements: Fp/2 chunks per row (at most 32 chunks in all)
    const int hc = Fp >> 1, wrow = lane / hc, wch = lane - wrow * hc;
    const int ew = __shfl_sync(0xffffffffu, e, wrow & 7);
    double2 wt = make_double2(0.0, 0.0);
    if (lane < kTile * hc && ew >= 0) wt = __ldg(reinterpret_cast<const double2 *>(ctbc + (size_t)ew * Fp + 2 * wch));
    if (wait_empty) mbar_wait(empty_bar, parity);
    const uint32_t sb = smem_u32(sl);
    const uint32_t vb = sb + (uint32_t)hdr_bytes + (uint32_t)(r_sub * NS * 8 + lch * 16);
#pragma unroll
    for (int k = 0; k < IT; ++k)
    {
        const int row = k * RPI + r_sub;
        if (lane_copies && row < kTile)
            st_shared_f64x2(vb + (uint32_t)(k * RPI * NS * 8), make_double2(fma(c2, ru[k].x, c1 * rt[k].x), fma(c2, ru[k].y, c1 * rt[k].y)));
    }
    if (lane < rows * 2) st_shared_u32x4(sb + lane * 16, dsc);
    if (lane < kTile * hc) st_shared_f64x2(sb + rows * 32 + lane * 16, wt);
    mbar_arrive(full_bar);  // release: this lane's stores are visible to whoever acquires the completed phase
}

// Neighbour producer, one tile whose prerequisites are stored: lane (mq, q) = (lane >> 3, lane & 7) holds in `code0` the source row of
// face mq, element q (negative: wall / no neighbour -> zeros).  first_vec: index of face 0's staged vector in the slot.
template <int N>
__device__ __forceinline__ void nbr_tile_regs(const double *u, int code0, int nin, const int32_t *desc_g, int first_vec, unsigned char *sl,
                                              size_t hdr_bytes, uint64_t *empty_bar, bool wait_empty, uint32_t parity, uint64_t *full_bar)
{
    constexpr int NP = (N + 1) & ~1, NS = stage_stride(N), CH = NP / 2, RPI = 32 / CH, IT = (kRegFaces * kTile + RPI - 1) / RPI;
    const int lane = threadIdx.x & 31, r_sub = lane / CH, lch = lane - r_sub * CH, q = lane & 7, mq = lane >> 3;
    const bool lane_copies = r_sub < RPI;
    const uint32_t vb0 = smem_u32(sl) + (uint32_t)hdr_bytes + (uint32_t)(r_sub * NS * 8 + lch * 16);
    for (int m0 = 0; m0 < nin; m0 += kRegFaces)
    {
        const int code = m0 == 0 ? code0 : ((mq < kRegFaces && m0 + mq < nin) ? __ldg(desc_g + (1 + m0 + mq) * kTile + q) : kSrcSkipCode);
        const int nrow = min(nin - m0, kRegFaces) * kTile;
        double2 rv[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k)
        {
            const int row = k * RPI + r_sub;  // lane `row` holds the source of row `row` (face m0 + row / 8, element row % 8)
            const int cr = __shfl_sync(0xffffffffu, code, row & 31);
            if constexpr (IT > 8) rv[k] = ld_cg_f64_f64_if(u + (size_t)(cr < 0 ? 0 : cr) * NP + 2 * lch, lane_copies && row < nrow && cr >= 0);
            else rv[k] = ld_cg_f64x2_if(u + (size_t)(cr < 0 ? 0 : cr) * NP + 2 * lch, lane_copies && row < nrow && cr >= 0);
        }
        if (wait_empty && m0 == 0) mbar_wait(empty_bar, parity);
        const uint32_t vb = vb0 + (uint32_t)((first_vec + m0) * kTile * NS * 8);
#pragma unroll
        for (int k = 0; k < IT; ++k)
        {
            const int row = k * RPI + r_sub;
            if (lane_copies && row < nrow) st_shared_f64x2(vb + (uint32_t)(k * RPI * NS * 8), rv[k]);
        }
    }
    if (wait_empty && nin == 0) mbar_wait(empty_bar, parity);
    mbar_arrive(full_bar);
}

// bytes of metadata and barriers: inflow masks, operator offsets; full/empty per ring slot; the bitmap of finished tiles; h_doubles
// wall-vector entries staged when they fit
__host__ __device__ inline size_t march_meta_bytes(int noclass, int ring, int max_class_tiles, size_t h_doubles, int wpt)
{
    size_t ints = 2 * (size_t)noclass + (size_t)wpt * ((size_t)(max_class_tiles + 31) / 32 + 1);
    ints = (ints + 3) & ~(size_t)3;
    return ints * sizeof(int32_t) + sizeof(uint64_t) * (2 * (size_t)ring) + sizeof(double) * h_doubles;
}

__device__ __forceinline__ int ld_acquire_gpu(const int32_t *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid = S.ncta persistent CTAs of march_threads(N) threads.  Dynamic shared memory = [operators (STAGE) | metadata | barriers | ring].
//
// WORK LIST.  The host cuts the tile sequences of all slots, laid end to end, into S.ncta equal shares (pbte_cuda.cu: plan_pieces):
// a CTA owns a list of PIECES (slot, first tile, end tile), at most one slot tail, whole slots, at most one slot head, so every SM
// carries the same load whatever the ratio of slots to SMs is (256 slots on 148 SMs left 40 SMs half empty before; 320 slots per GPU
// were 1.08 waves).  A CTA works through head, whole slots, tail in that order: the tail of a slot continues where ANOTHER CTA, which
// started with that slot's head, stopped — it waits for `progress[slot]` (release/acquire at GPU scope) and finds the head's flux rows
// in L2.  A CTA never waits for a CTA with a higher index, and what it waits for was started first there: no cycle, and the grid is
// no larger than what is resident at once.
//
// PIPELINE inside a piece, tiles i = t0 .. t1-1 of the slot's sequence (level-major), counted g = gbase + (i - t0) over the CTA's
// whole list; ring slot = g mod R:
//   source producer   waits empty[slot]; cp.async of the tile descriptor and of the u_old / T_old rows of its 8 elements.
//                     Independent of the current sweep: runs up to R tiles ahead and hides the HBM latency.
//   neighbour producer waits empty[slot] and until every tile up to the tile's last prerequisite (descriptor header) is stored:
//                     consumers set a bit per finished tile, the producer advances a watermark over the bitmap (no level-wide
//                     barrier; tiles before t0 were finished by the CTA this piece waited for); cp.async of the upwind neighbour rows
//                     (L2 hits).  Both producers complete full[slot].
//   consumer          waits full[slot]; U = X P^T + sum_m Unbr_m Q_m^T: column blocks by mma.sync m8n8k4 FP64 with A fragments from
//                     the ring and B fragments from the staged operators, remainder columns by DFMAs (Split<N>); releases
//                     empty[slot]; adds wall terms, stores the tile with 16-byte stores, sets its bit in the bitmap (release).
template <int N, bool STAGE>
__global__ void __launch_bounds__(march_threads(N), march_ctas_per_sm(N)) sweep_march_kernel(DevProblem P, DevMarch S)
{
    constexpr int kConsumerWarps = march_consumers(N);
    constexpr int kSrcWarps = march_src_warps(N), kNbrWarps = march_nbr_warps(N);
    using G = Split<N, STAGE>;
    constexpr int MTM = G::MTM, REM = G::REM, REMP = G::REMP, P2 = G::P2, KS = G::KS, MAIN = G::MAIN;
    constexpr int OPS = STAGE ? G::OPS : compact_doubles(N);  // doubles per operator: staged operators are in split order, others compact
    constexpr int NP = (N + 1) & ~1;        // global row stride (doubles): 16-byte aligned rows
    constexpr int NS = stage_stride(N);     // shared-memory row stride
    constexpr int CH = NP / 2;              // 16-byte chunks per row
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    if (*P.stop) return;  // the iteration before this one met the stop test (csrc/tail.cuh)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = S.F, rows = 2 + F, noc = S.noclass, R = S.ring, nv = S.nv;

    // ---- shared memory carve-up (sizes are those of the largest class: the same for every piece) -------------------------------
    const size_t op_bytes = STAGE ? (size_t)S.max_op_cnt * (OPS * sizeof(double)) : 0;
    int32_t *s_inm = reinterpret_cast<int32_t *>(sm_raw + op_bytes);  // [noclass] inflow masks of the piece's upwind-DAG class
    int32_t *s_opoff = s_inm + noc;                                    // [noclass] operator offsets (units of OPS)
    constexpr int WPT = march_wpt(N);  // consumer warps per tile
    constexpr int kTileGroups = kConsumerWarps / WPT;
    const int nwords = (S.max_class_tiles + 31) / 32 + 1;
    unsigned *s_done = reinterpret_cast<unsigned *>(s_opoff + noc);    // [WPT][nwords] bit i: warp h of its group has stored its columns of tile i
    const size_t meta_ints = (2 * (size_t)noc + (size_t)WPT * nwords + 3) & ~(size_t)3;
    uint64_t *full = reinterpret_cast<uint64_t *>(sm_raw + op_bytes + meta_ints * sizeof(int32_t));  // [R]
    uint64_t *empty = full + R;                                                                     // [R]
    const int Fp = (F + 1) & ~1;
    const size_t hdr_bytes = (size_t)rows * 32 + (size_t)kTile * Fp * sizeof(double);  // descriptor + wall temperatures
    const size_t slot_bytes = hdr_bytes + (size_t)nv * kTile * NS * sizeof(double);
    double *s_h = reinterpret_cast<double *>(empty + R);  // [noclass][F][N] wall vectors of the piece's slot (when S.h_smem)
    const size_t h_doubles = S.h_smem ? (size_t)noc * F * N : 0;
    unsigned char *ring = reinterpret_cast<unsigned char *>(s_h + h_doubles);
    // 16-byte alignment as an OFFSET from the shared-memory base (an integer round trip of the pointer would make every ring access a
    // generic load instead of LDS)
    ring = sm_raw + ((static_cast<size_t>(ring - sm_raw) + 15) & ~static_cast<size_t>(15));

    // the ring starts zeroed: copies write NP doubles per row, the k padding up to NS stays zero for the whole kernel
    for (size_t i = threadIdx.x; i < (size_t)R * slot_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 2 * R; i += blockDim.x) mbar_init(full + i, i < R ? 2 * 32 : WPT);  // full: one source and one neighbour warp,
                                                                                                      // every lane; empty: lane 0 of the tile's consumers
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // Copy pattern of the producers.  A tile needs 16 source rows (u_old and T_old of its 8 elements) and 8 rows per inflow
    // face, NP doubles each.  A lane owns one 16-byte chunk column and an instruction covers RPI = 32 / CH whole rows: rows are
    // contiguous, so one cp.async instruction still touches full 32-byte sectors and few cache lines (a row per lane would be
    // bound by the one-line-per-cycle tag stage of L1TEX: 16 B/clk/SM measured, tools/copy_probe.cu), and per instruction a
    // lane needs just one shuffle (the row's element) and one multiply-add (the address): the shared-memory offset is a
    // per-lane constant plus an immediate.
    constexpr int RPI = 32 / CH;
    static_assert(RPI >= 1, "a row is at most 32 chunks of 16 bytes");
    const int r_sub = lane / CH, lch = lane - r_sub * CH;
    const bool lane_copies = r_sub < RPI;
    const uint32_t lane_off = (uint32_t)(r_sub * NS * 8 + lch * 16);  // offset of the lane's chunk inside the staged vectors

    int gbase = 0;  // tiles of the pieces before the current one (ring slots and mbarrier phases run on over the whole list)
    const int pc_begin = S.cta_piece_begin[blockIdx.x], pc_end = S.cta_piece_begin[blockIdx.x + 1];
    for (int pc = pc_begin; pc < pc_end; ++pc)
    {
        const int4 piece = reinterpret_cast<const int4 *>(S.pieces)[pc];  // slot, first tile, end tile, tiles of the slot
        const int s = piece.x, t0 = piece.y, t1 = piece.z, NTs = piece.w, ntile = t1 - t0;
        const int c = S.slot_class[s];
        const int tbase = S.level_tile_begin[S.cls_level_begin[c]];
        const double *ops = S.ops + S.slot_op_base[s];
        // ---- piece set-up (every warp is past the previous piece: barrier at the end of the loop body) ------------------------------
        if (t0 > 0 && threadIdx.x == 0)
            while (ld_acquire_gpu(S.progress + s) < t0) __nanosleep(64);  // the CTA that holds the part before t0 has stored it
        for (int i = threadIdx.x; i < WPT * nwords; i += blockDim.x) s_done[i] = 0u;
        for (int i = threadIdx.x; i < noc; i += blockDim.x)
        {
            s_inm[i] = S.cls_inmask[(size_t)c * noc + i];
            s_opoff[i] = S.cls_op_off[(size_t)c * noc + i];
        }
        for (size_t i = threadIdx.x; i < h_doubles; i += blockDim.x) s_h[i] = S.oph[(size_t)s * noc * F * N + i];
        if (STAGE)
        {
            // one TMA bulk copy stages the slot's operators; phase parity of s_bar = piece index
            const uint32_t bytes = (uint32_t)((size_t)S.cls_op_cnt[c] * (OPS * sizeof(double)));
            if (threadIdx.x == 0)
            {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bytes) : "memory");
                for (uint32_t done = 0; done < bytes;)
                {
                    const uint32_t chunk = min(bytes - done, 65536u);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm_raw + done)),
                                 "l"(reinterpret_cast<const char *>(ops) + done), "r"(chunk), "r"(smem_u32(&s_bar))
                                 : "memory");
                    done += chunk;
                }
            }
            mbar_wait(&s_bar, (uint32_t)(pc - pc_begin) & 1u);
            ops = reinterpret_cast<const double *>(sm_raw);
        }
        __syncthreads();

        double *u = P.flux + (size_t)s * P.ne * NP;
        const int32_t *desc0 = S.tile_desc + (size_t)(tbase + t0) * rows * kTile;  // descriptor of the piece's first tile
        if (warp >= kConsumerWarps && warp < kConsumerWarps + kSrcWarps)
        {
            // ---- source producer ------------------------------------------------------------------------------------
            constexpr int ITS = (16 + RPI - 1) / RPI;
            const unsigned char *ub = reinterpret_cast<const unsigned char *>(u) + lch * 16, *tb = reinterpret_cast<const unsigned char *>(P.tc_prev) + lch * 16;
            const int q = lane & 7;  // lanes 0-7 hold the elements of the u_old rows, lanes 8-15 the same elements for the T_old rows
            const int first = (warp - kConsumerWarps - gbase % kSrcWarps + kSrcWarps) % kSrcWarps;
            int e_next = first < ntile ? __ldg(desc0 + (size_t)first * rows * kTile + q) : -1;
            // Experiment (off): the ring cannot run far ahead of the consumers (a slot per consumer warp), so the u_old rows of a tile come
            // from HBM at the moment its slot frees; pull the rows of the tile kPrefetchAhead tiles further on into L2 now (one 128-byte
            // line per lane: 8 rows x 2 lines).
#ifndef PBTE_L2_PREFETCH
#define PBTE_L2_PREFETCH 0  // measured on C2 / C5 / C3 slices: no effect at 8, 16 or 32 tiles ahead (profiles/r02_kernel_probes.txt); kept as a switch
#endif
            constexpr int kPrefetchAhead = PBTE_L2_PREFETCH;
            [[maybe_unused]] int e_pf = (kPrefetchAhead > 0 && first + kPrefetchAhead < ntile) ? __ldg(desc0 + (size_t)(first + kPrefetchAhead) * rows * kTile + q) : -1;
            [[maybe_unused]] const double xc1 = P.slot_c1[s], xc2 = P.slot_c2[s];
            for (int j = first; j < ntile; j += kSrcWarps)
            {
                const int g = gbase + j, slot = g % R, it = g / R;
                const int e = e_next;
                const int32_t *desc_g = desc0 + (size_t)j * rows * kTile;
                if (j + kSrcWarps < ntile) e_next = __ldg(desc_g + (size_t)kSrcWarps * rows * kTile + q);
                if constexpr (march_xsrc(N))
                {
                    src_tile_regs<N>(u, P.tc_prev, e, desc_g, S.ctbc, rows, Fp, xc1, xc2, ring + (size_t)slot * slot_bytes, hdr_bytes, empty + slot,
                                     it > 0, (uint32_t)(it - 1) & 1u, full + slot);
                    PBTE_STAMP(t0 + j, 8);
                }
                else
                {
                if constexpr (kPrefetchAhead > 0)
                {
                    const int ep = __shfl_sync(0xffffffffu, e_pf, lane >> 1);  // lanes 2q, 2q+1: the two lines of row q
                    if (lane < 16 && ep >= 0)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const unsigned char *>(u) + (size_t)ep * (NP * 8) + (lane & 1) * 128));
                    e_pf = (j + kSrcWarps + kPrefetchAhead < ntile) ? __ldg(desc_g + (size_t)(kSrcWarps + kPrefetchAhead) * rows * kTile + q) : -1;
                }
                if (it > 0) mbar_wait(empty + slot, (it - 1) & 1);
                unsigned char *sl = ring + (size_t)slot * slot_bytes;
                PBTE_STAMP(t0 + j, 7);
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
                    // wall temperatures of the 8 elements: Fp/2 chunks per row (at most 32 chunks in all)
                    const int hc = Fp >> 1, row = lane / hc, ch = lane - row * hc;
                    const int er = __shfl_sync(0xffffffffu, e, row & 7);
                    if (lane < kTile * hc)
                        cp_async16(sl + rows * 32 + lane * 16, S.ctbc + (size_t)(er < 0 ? 0 : er) * Fp + 2 * ch, er >= 0 ? 16u : 0u);
                }
                cp_async_arrive(full + slot);
                PBTE_STAMP(t0 + j, 8);
                }
            }
        }
        else if (warp >= kConsumerWarps + kSrcWarps)
        {
            // ---- upwind-neighbour producer --------------------------------------------------------------------------
            constexpr int ITN = (32 + RPI - 1) / RPI;  // up to 4 inflow faces (32 rows) per pass
            const unsigned char *ub = reinterpret_cast<const unsigned char *>(u) + lch * 16;
            const int q = lane & 7, mq = lane >> 3;
            int wm = t0;  // watermark: tiles [0, wm) of this slot are known to be stored
            const int first = (warp - kConsumerWarps - kSrcWarps - gbase % kNbrWarps + kNbrWarps) % kNbrWarps;
            const int32_t *d0 = desc0 + (size_t)first * rows * kTile;
            int o_next = first < ntile ? __ldg(d0 + (1 + F) * kTile) : 0, L_next = first < ntile ? __ldg(d0 + (1 + F) * kTile + 3) : -1;
            int code_next = first < ntile ? __ldg(d0 + (1 + mq) * kTile + q) : -1;  // rows 1..4 always exist (rows = 2 + F, F >= 3)
            for (int j = first; j < ntile; j += kNbrWarps)
            {
                const int g = gbase + j, slot = g % R, it = g / R;
                const int o = o_next, dep = L_next, code0 = code_next;
                const int32_t *desc_g = desc0 + (size_t)j * rows * kTile;
                if (j + kNbrWarps < ntile)
                {
                    const int32_t *dn = desc_g + (size_t)kNbrWarps * rows * kTile;
                    o_next = __ldg(dn + (1 + F) * kTile);
                    L_next = __ldg(dn + (1 + F) * kTile + 3);
                    code_next = __ldg(dn + (1 + mq) * kTile + q);
                }
                if constexpr (!march_xnbr(N))
                    if (it > 0) mbar_wait(empty + slot, (it - 1) & 1);
                PBTE_STAMP(t0 + j, 4);
                // every tile up to the last prerequisite must be stored: advance the watermark over the bitmap of finished tiles
#ifdef PBTE_NODEP  // timing experiment only (wrong results): how fast is the pipeline when no tile waits for another?
                wm = dep + 1;
#endif
                while (wm <= dep)
                {
                    unsigned w = *reinterpret_cast<volatile unsigned *>(s_done + (wm >> 5));
                    if constexpr (WPT == 2) w &= *reinterpret_cast<volatile unsigned *>(s_done + nwords + (wm >> 5));
                    w >>= (wm & 31);
                    const int ones = (~w == 0u) ? 32 : __ffs(~w) - 1;  // consecutive finished tiles starting at wm (zeros are shifted in)
                    if (ones == 0) __nanosleep(32);
                    else wm += min(ones, 32 - (wm & 31));
                }
                __threadfence_block();  // acquire: the stores of those tiles are visible to the copies issued below
                PBTE_STAMP(t0 + j, 5);
                const int nin = __popc((unsigned)s_inm[o]);
                unsigned char *sl = ring + (size_t)slot * slot_bytes;
                if constexpr (march_xnbr(N))
                {
                    // rows into registers now, into the slot once it is free
                    nbr_tile_regs<N>(u, code0, nin, desc_g, 1, sl, hdr_bytes, empty + slot, it > 0, (uint32_t)(it - 1) & 1u, full + slot);
                    PBTE_STAMP(t0 + j, 6);
                }
                else
                {
                for (int m0 = 0; m0 < nin; m0 += 4)
                {
                    // lane (mq, q) holds the source code of row mq*8 + q of this pass
                    const int code = m0 == 0 ? code0 : ((m0 + mq < nin) ? __ldg(desc_g + (1 + m0 + mq) * kTile + q) : kSrcSkipCode);
                    const int nrow = min(nin - m0, 4) * kTile;
                    const uint32_t vb = smem_u32(sl + hdr_bytes) + (uint32_t)(((march_xsrc(N) ? 1 : 2) + m0) * kTile * NS * 8) + lane_off;
#pragma unroll
                    for (int k = 0; k < ITN; ++k)
                    {
                        if (k * RPI >= nrow) break;
                        const int row = k * RPI + r_sub;  // lane `row` holds the source of row `row` (face m0 + row / 8, element row % 8)
                        const int cr = __shfl_sync(0xffffffffu, code, row & 31);
                        if (lane_copies && row < nrow)
                        {
                            const unsigned char *src = ub + (size_t)(cr < 0 ? 0 : cr) * (NP * 8);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vb + (uint32_t)(k * RPI * NS * 8)), "l"(src), "r"(cr >= 0 ? 16u : 0u) : "memory");
                        }
                    }
                }
                cp_async_arrive(full + slot);
                PBTE_STAMP(t0 + j, 6);
                }
            }
        }
        else
        {
            // ---- consumers ----------------------------------------------------------------------------------------------
            const int q = lane >> 2, r = lane & 3;
            const double c1 = P.slot_c1[s], c2 = P.slot_c2[s];
            const double *oph = S.h_smem ? s_h : S.oph + (size_t)s * noc * F * N;
            // One pass = one tile of one operator class: U = X P^T + sum_m Unbr_m Q_m^T.  One matrix block = KS k steps; per k step
            // the B fragments of the warp's column blocks (two per 16-byte load from the staged operator) feed its MMAs and the REM
            // remainder entries feed REM DFMAs, the A fragment comes from the staged vectors.  The fragments of KC k steps are
            // loaded together into their own registers and only then multiplied: no register is rewritten while MMAs that read it
            // are in flight and every MMA finds its operands ready (interleaved load/MMA code ran at the MMA latency).
            // HALF: the warp's share of the columns: -1 all (one warp per tile); 0 = column blocks 0 and 1; 1 = the other blocks and
            // the DFMA remainder (two warps per tile, march_wpt).
            auto tile_pass = [&](auto half_c, int g, int i) {
                constexpr int HALF = decltype(half_c)::value;
                constexpr int PB = HALF == 1 ? 1 : 0, PE = HALF == 0 ? 1 : P2;                 // pairs of column blocks [PB, PE)
                constexpr bool SINGLE = (MTM & 1) && HALF != 0;                                // the odd last block
                constexpr int MYREM = HALF == 0 ? 0 : REM, MYREMP = HALF == 0 ? 0 : REMP;      // DFMA remainder columns
                constexpr int IB = 2 * PB, IE = SINGLE ? MTM : 2 * PE;                         // column blocks [IB, IE)
#ifdef PBTE_KC
                constexpr int KC0 = PBTE_KC;
#else
                constexpr int KC0 = (MYREM > 0 ? 28 : 16) / ((IE - IB) + MYREMP);
#endif
                constexpr int KC = KC0 < 1 ? 1 : (KC0 > KS ? KS : KC0);
                const unsigned char *sl = ring + (size_t)(g % R) * slot_bytes;
                const int32_t *sd = reinterpret_cast<const int32_t *>(sl);
                const double *vec = reinterpret_cast<const double *>(sl + hdr_bytes) + (size_t)q * NS + r;
                const int e = sd[q];
                const int o = sd[(1 + F) * kTile];
                const unsigned inm = (unsigned)s_inm[o];
                const int nin = __popc(inm);
                const double *op = ops + (size_t)s_opoff[o] * OPS;
                if (!STAGE)
                {
                    // operators of many classes stay in global memory (L2): ask for all of this pass's fragments now, one cache line per
                    // lane and round, so that the later matrix blocks find them in L1 instead of paying an L2 round trip each
                    const char *pb = reinterpret_cast<const char *>(op);
                    const int bytes = (1 + nin) * OPS * (int)sizeof(double);
                    for (int off = lane * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + off));
                }
                // what the descriptor says about the inflow faces: the wall faces of this lane's element, and the faces where some
                // element of the tile has an interior upwind neighbour
                unsigned wall = 0, interior = 0;
                {
                    // rows 1..kMaxIn of a descriptor always exist (rows = 2 + F, F >= 3): read them all at once, then vote
                    int code[kMaxIn];
#pragma unroll
                    for (int m = 0; m < kMaxIn; ++m) code[m] = sd[(1 + m) * kTile + q];
#pragma unroll
                    for (int m = 0; m < kMaxIn; ++m)
                    {
                        wall |= (unsigned)(code[m] == kSrcWallCode && m < nin) << m;
                        interior |= (unsigned)(__ballot_sync(0xffffffffu, code[m] >= 0 && m < nin) != 0u) << m;
                    }
                    for (int m = kMaxIn; m < nin; ++m)  // elements with more inflow faces than any supported shape has today
                    {
                        const int c2m = sd[(1 + m) * kTile + q];
                        wall |= (unsigned)(c2m == kSrcWallCode) << m;
                        interior |= (unsigned)(__ballot_sync(0xffffffffu, c2m >= 0) != 0u) << m;
                    }
                }
                double acc[MTM > 0 ? MTM : 1][2], accr[REMP > 0 ? REMP : 1];
#pragma unroll
                for (int I = 0; I < MTM; ++I) acc[I][0] = acc[I][1] = 0.0;
#pragma unroll
                for (int cc = 0; cc < REMP; ++cc) accr[cc] = 0.0;
                auto block = [&](size_t voff, bool source, const double *opb) {
#pragma unroll
                    for (int k0 = 0; k0 < KS; k0 += KC)
                    {
                        double2 bq[KC][P2 > 0 ? P2 : 1];
                        double bs[KC], br[KC][REMP > 0 ? REMP : 1], xa[KC];
#pragma unroll
                        for (int k = 0; k < KC; ++k)
                            if (k0 + k < KS)
                            {
                                const int kk = k0 + k;
                                if constexpr (STAGE)
                                {
                                    const double *bk = opb + kk * (32 * MTM);
#pragma unroll
                                    for (int p = PB; p < PE; ++p) bq[k][p] = *reinterpret_cast<const double2 *>(bk + p * 64 + lane * 2);
                                    if constexpr (SINGLE) bs[k] = bk[P2 * 64 + lane];
                                    if constexpr (MYREM > 0)
                                    {
                                        const double *rk = opb + MAIN + (kk * 4 + r) * REMP;
#pragma unroll
                                        for (int cc = 0; cc < REMP; cc += 2)
                                        {
                                            const double2 t = *reinterpret_cast<const double2 *>(rk + cc);
                                            br[k][cc] = t.x;
                                            br[k][cc + 1] = t.y;
                                        }
                                    }
                                }
                                else
                                {
                                    // compact row-major operator: this lane's B entries are R[row][col]
                                    const int col = 4 * kk + r;
#pragma unroll
                                    for (int p = PB; p < PE; ++p)
                                    {
                                        const int row = 16 * p + q;
                                        bq[k][p].x = (row < N && col < N) ? __ldg(opb + row * N + col) : 0.0;
                                        bq[k][p].y = (row + 8 < N && col < N) ? __ldg(opb + (row + 8) * N + col) : 0.0;
                                    }
                                    if constexpr (SINGLE)
                                    {
                                        const int row = 16 * P2 + q;
                                        bs[k] = (row < N && col < N) ? __ldg(opb + row * N + col) : 0.0;
                                    }
#pragma unroll
                                    for (int cc = 0; cc < MYREM; ++cc) br[k][cc] = col < N ? __ldg(opb + (8 * MTM + cc) * N + col) : 0.0;
                                }
                                const double *va = vec + voff;
                                xa[k] = source ? fma(c2, va[4 * kk], c1 * va[(size_t)kTile * NS + 4 * kk]) : va[4 * kk];
                            }
#pragma unroll
                        for (int k = 0; k < KC; ++k)
                            if (k0 + k < KS)
                            {
#pragma unroll
                                for (int I = IB; I < IE; ++I)
                                {
                                    double b;
                                    if (I < 2 * P2) b = (I & 1) ? bq[k][I >> 1].y : bq[k][I >> 1].x;
                                    else b = bs[k];
                                    dmma_m8n8k4(acc[I][0], acc[I][1], xa[k], b);
                                }
#pragma unroll
                                for (int cc = 0; cc < MYREM; ++cc) accr[cc] = fma(xa[k], br[k][cc], accr[cc]);
                            }
                    }
                };
                PBTE_STAMP(i, 9);
                // source / pseudo-time term: X = c1 T_old + c2 u_old, formed here from vectors 0 and 1 of the slot or (march_xsrc) by the producer
                constexpr int kFaceVec = march_xsrc(N) ? 1 : 2;
                block(0, !march_xsrc(N), op);
                for (int m = 0; m < nin; ++m)
                    if (interior >> m & 1u) block((size_t)(kFaceVec + m) * kTile * NS, false, op + (size_t)(1 + m) * OPS);
                // remainder columns: the four lanes of an element row add their partial sums
#pragma unroll
                for (int cc = 0; cc < MYREM; ++cc)
                {
                    accr[cc] += __shfl_xor_sync(0xffffffffu, accr[cc], 1);
                    accr[cc] += __shfl_xor_sync(0xffffffffu, accr[cc], 2);
                }
                // lane r stores the remainder columns 8 MTM + 2r, + 1
                double rem0 = 0.0, rem1 = 0.0;
                if constexpr (MYREM > 0)
                {
                    rem0 = r == 0 ? accr[0] : (REMP > 2 ? accr[REMP > 2 ? 2 : 0] : 0.0);
                    rem1 = r == 0 ? (REM > 1 ? accr[REM > 1 ? 1 : 0] : 0.0) : (REM > 3 ? accr[REM > 3 ? 3 : 0] : 0.0);
                }
#ifdef PBTE_TRACE
                if (acc[IB][0] == 1.2345e300) PBTE_STAMP(i, 15);  // the stamp below waits for the accumulators
#endif
                PBTE_STAMP(i, 2);
                // isothermal walls on inflow faces: + T_bc h_j   (this lane owns element q, dofs 8I + 2r + {0,1})
                if (wall)
                {
                    const double *tbq = reinterpret_cast<const double *>(sl + rows * 32) + (size_t)q * Fp;  // wall temperatures of element q
                    unsigned rest = inm;
                    for (int m = 0; rest; ++m, rest &= rest - 1)
                        if (wall >> m & 1u)
                        {
                            const int jf = __ffs(rest) - 1;
                            const double tw = tbq[jf];
                            const double *hj = oph + ((size_t)o * F + jf) * N;
#pragma unroll
                            for (int I = IB; I < IE; ++I)
                            {
                                const int d0 = 8 * I + 2 * r;
                                if (d0 < N) acc[I][0] = fma(tw, hj[d0], acc[I][0]);
                                if (d0 + 1 < N) acc[I][1] = fma(tw, hj[d0 + 1], acc[I][1]);
                            }
                            if constexpr (MYREM > 0)
                            {
                                const int d0 = 8 * MTM + 2 * r;
                                if (d0 < N) rem0 = fma(tw, hj[d0], rem0);
                                if (d0 + 1 < N) rem1 = fma(tw, hj[d0 + 1], rem1);
                            }
                        }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + g % R);  // this warp has consumed every staged value it needs of the tile
                if (e >= 0)
                {
                    // 16-byte stores; an odd ndof has one padding column per row, which stays zero
                    double *dst = u + (size_t)e * NP;
#pragma unroll
                    for (int I = IB; I < IE; ++I)
                    {
                        const int d0 = 8 * I + 2 * r;
#ifdef PBTE_STORE8
                        if (d0 < N) dst[d0] = acc[I][0];
                        if (d0 + 1 < N) dst[d0 + 1] = acc[I][1];
#else
                        if (d0 < NP) *reinterpret_cast<double2 *>(dst + d0) = make_double2(acc[I][0], d0 + 1 < N ? acc[I][1] : 0.0);
#endif
                    }
                    if constexpr (MYREM > 0)
                    {
                        const int d0 = 8 * MTM + 2 * r;
                        if (d0 < NP) *reinterpret_cast<double2 *>(dst + d0) = make_double2(rem0, d0 + 1 < N ? rem1 : 0.0);
                    }
                }
                PBTE_STAMP(i, 10);
                __syncwarp();  // orders the stores of all lanes before lane 0's release
                if (lane == 0)
                {
                    __threadfence_block();
                    atomicOr(s_done + (HALF == 1 ? nwords : 0) + (i >> 5), 1u << (i & 31));
                }
                PBTE_STAMP(i, 3);
            };
            const int grp = warp / WPT, half = warp - grp * WPT;
            const int first = (grp - gbase % kTileGroups + kTileGroups) % kTileGroups;
            for (int j = first; j < ntile; j += kTileGroups)
            {
                const int g = gbase + j, i = t0 + j;
                PBTE_STAMP(i, 0);
                mbar_wait(full + g % R, (g / R) & 1);
                PBTE_STAMP(i, 1);
                if constexpr (WPT == 1) tile_pass(std::integral_constant<int, -1>{}, g, i);
                else if (half == 0) tile_pass(std::integral_constant<int, 0>{}, g, i);
                else tile_pass(std::integral_constant<int, 1>{}, g, i);
            }
        }
        gbase += ntile;
        // ---- end of the piece: every warp is through, every tile is stored ----------------------------------------------------------
        __syncthreads();
        if (threadIdx.x == 0 && (t0 > 0 || t1 < NTs))
        {
            __threadfence();
            st_release_gpu(S.progress + s, t1 < NTs ? t1 : 0);  // hand the slot on, or leave the counter ready for the next sweep
        }
    }
}

// bytes of metadata and barriers of the classic kernel: level offsets, inflow masks, operator offsets; full/empty per ring slot; the
// bitmap of finished tiles; h_doubles wall-vector entries staged when they fit
__host__ __device__ inline size_t march_classic_meta_bytes(int max_levels, int noclass, int ring, int max_class_tiles, size_t h_doubles)
{
    size_t ints = (size_t)max_levels + 1 + 2 * (size_t)noclass + (size_t)(max_class_tiles + 31) / 32 + 1;
    ints = (ints + 3) & ~(size_t)3;
    return ints * sizeof(int32_t) + sizeof(uint64_t) * (2 * (size_t)ring) + sizeof(double) * h_doubles;
}

// One TMA bulk copy global -> shared of `bytes` (multiple of 16) issued by one thread, completion on an mbarrier.
__device__ __forceinline__ void classic_bulk_stage(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        // chunks of at most 64 KB keep every copy well inside the bulk-copy size limits
        uint32_t done = 0;
        while (done < bytes)
        {
            const uint32_t chunk = min(bytes - done, 65536u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32((char *)dst + done)),
                         "l"((const char *)src + done), "r"(chunk), "r"(smem_u32(bar))
                         : "memory");
            done += chunk;
        }
    }
    // every thread waits for phase 0 of the barrier
    uint32_t ok = 0;
    while (!ok)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar))
            : "memory");
    }
}

// (Measured and dropped in round 2, profiles/r02_kernel_probes.txt run n: handing the ring slot back before the last matrix block of
// a tile - all A fragments of that block loaded first, the hand-over guarded by a consumer of every load - costs more in exposed load
// latency than the earlier refill returns: C2 0.392 against 0.369 ms, C3 slice 3.29 against 2.73 ms.)
// One CTA per slot, one pass over the slot's tiles (the kernel of round 1): kept beside the persistent kernel because it is the faster of
// the two wherever cutting slots into pieces buys nothing (no more slots than resident CTAs: C2) and for operators read from global
// memory (tetrahedra): measured side by side in profiles/r02_kernel_probes.txt (C2 0.380 against 0.389 ms, C3 slice 3.59 against 3.89 ms;
// C5 slice 2.34 against 2.07 ms the other way round).  Its staged operators are in fragment-pair order (Frag<N>).
// grid = nslots CTAs of march_threads(N) threads.  Dynamic shared memory = [operators (STAGE) | metadata | barriers | tile ring].
//
// Warp-specialised pipeline over the slot's tile sequence i = 0 .. NT-1 (level-major, so round-robin assignment keeps the
// consumers balanced inside every level); ring slot = i mod R:
//   source producer   waits empty[slot]; cp.async of the tile descriptor and of the u_old / T_old rows of its 8 elements.
//                     Independent of the current sweep: runs up to R tiles ahead and hides the HBM latency.
//   neighbour producer waits empty[slot] and until every tile up to the tile's last prerequisite (descriptor header) is stored:
//                     consumers set a bit per finished tile, the producer advances a watermark over the bitmap (no level-wide
//                     barrier: tiles of the next level start as soon as their own upwind tiles are done); cp.async of the
//                     upwind neighbour rows (L2 hits: written by this CTA a level ago).  Both producers complete full[slot].
//   consumer w        tiles w, w + consumers, ...: waits full[slot]; U = X P^T + sum_m Unbr_m Q_m^T by mma.sync m8n8k4 FP64 with A
//                     fragments from the ring and B fragments from the staged operators; releases empty[slot]; adds wall
//                     terms, stores the tile, sets its bit in the bitmap (release).
template <int N, bool STAGE>
__global__ void __launch_bounds__(march_threads(N), march_ctas_per_sm(N)) sweep_march_classic_kernel(DevProblem P, DevMarch S)
{
    constexpr int kConsumerWarps = march_consumers(N);
    constexpr int kSrcWarps = march_src_warps(N), kNbrWarps = march_nbr_warps(N);
#ifndef PBTE_PAIR_TILES
#define PBTE_PAIR_TILES 0  // experiment: a consumer warp takes two tiles of one operator class per pass, every B fragment feeds two MMAs
#endif
    constexpr bool kPairTiles = PBTE_PAIR_TILES != 0;
    using G = Frag<N>;
    constexpr int MT = G::MT, MP = G::MP, KS = G::KS, FR = G::FR;
#ifndef PBTE_COMPACT_OPS
#define PBTE_COMPACT_OPS 1  // 0 (experiment): unstaged operators in fragment-pair order as well (192 instead of 100 doubles for ndof 10, one 16-byte load per k step)
#endif
    constexpr bool FRAG_B = STAGE || PBTE_COMPACT_OPS == 0;
    constexpr int OPS = FRAG_B ? FR : compact_doubles(N);  // doubles per operator: staged operators are in fragment order, others compact
    constexpr int NP = (N + 1) & ~1;        // global row stride (doubles): 16-byte aligned rows
    constexpr int NS = stage_stride(N);     // shared-memory row stride
    constexpr int CH = NP / 2;              // 16-byte chunks per row
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    if (*P.stop) return;  // the iteration before this one met the stop test (csrc/tail.cuh)
    const int s = blockIdx.x;
    const int c = S.slot_class[s];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = S.F, rows = 2 + F, noc = S.noclass, R = S.ring, nv = S.nv;
    const double *ops = S.ops + S.slot_op_base[s];
    const int lb = S.cls_level_begin[c], nlev = S.cls_level_begin[c + 1] - lb - 1;
    const size_t op_bytes = STAGE ? (size_t)S.cls_op_cnt[c] * (FR * sizeof(double)) : 0;

    // ---- shared memory carve-up ---------------------------------------------------------------------------------
    int32_t *ltb = reinterpret_cast<int32_t *>(sm_raw + op_bytes);  // [nlev+1] tile offsets of the levels
    int32_t *s_inm = ltb + nlev + 1;                                 // [noclass] inflow masks
    int32_t *s_opoff = s_inm + noc;                                  // [noclass] operator offsets (units of FR)
    const int tbase0 = S.level_tile_begin[lb], NT0 = S.level_tile_begin[lb + nlev] - tbase0;
    const int nwords = (NT0 + 31) / 32 + 1;
    unsigned *s_done = reinterpret_cast<unsigned *>(s_opoff + noc);  // bit i: tile i of this class is stored
    const size_t meta_ints = ((size_t)nlev + 1 + 2 * (size_t)noc + (size_t)nwords + 3) & ~(size_t)3;
    uint64_t *full = reinterpret_cast<uint64_t *>(sm_raw + op_bytes + meta_ints * sizeof(int32_t));  // [R]
    uint64_t *empty = full + R;                                                                     // [R]
    const int Fp = (F + 1) & ~1;
    const size_t hdr_bytes = (size_t)rows * 32 + (size_t)kTile * Fp * sizeof(double);  // descriptor + wall temperatures
    const size_t slot_bytes = hdr_bytes + (size_t)nv * kTile * NS * sizeof(double);
    double *s_h = reinterpret_cast<double *>(empty + R);  // [noclass][F][N] wall vectors of this slot (when S.h_smem)
    const size_t h_doubles = S.h_smem ? (size_t)noc * F * N : 0;
    unsigned char *ring = reinterpret_cast<unsigned char *>(s_h + h_doubles);
    // 16-byte alignment as an OFFSET from the shared-memory base (an integer round trip of the pointer would make every ring access a
    // generic load instead of LDS)
    ring = sm_raw + ((static_cast<size_t>(ring - sm_raw) + 15) & ~static_cast<size_t>(15));
    for (size_t i = threadIdx.x; i < h_doubles; i += blockDim.x) s_h[i] = S.oph[(size_t)s * noc * F * N + i];

    for (int i = threadIdx.x; i <= nlev; i += blockDim.x) ltb[i] = S.level_tile_begin[lb + i];
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) s_done[i] = 0u;
    for (int i = threadIdx.x; i < noc; i += blockDim.x)
    {
        s_inm[i] = S.cls_inmask[(size_t)c * noc + i];
        s_opoff[i] = S.cls_op_off[(size_t)c * noc + i];
    }
    // the ring starts zeroed: copies write NP doubles per row, the k padding up to NS stays zero for the whole kernel
    for (size_t i = threadIdx.x; i < (size_t)R * slot_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * R; i += blockDim.x) mbar_init(full + i, i < R ? 2 * 32 : 1);  // full: one source and one neighbour warp,
                                                                                                    // every lane; empty: lane 0 of the consumer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (STAGE)
    {
        classic_bulk_stage(sm_raw, ops, (uint32_t)op_bytes, &s_bar);  // contains a CTA barrier
        ops = reinterpret_cast<const double *>(sm_raw);
    }
    __syncthreads();

    const int tbase = ltb[0], NT = ltb[nlev] - tbase;
    double *u = P.flux + (size_t)s * P.ne * NP;

    // Copy pattern of the producers.  A tile needs 16 source rows (u_old and T_old of its 8 elements) and 8 rows per inflow
    // face, NP doubles each.  A lane owns one 16-byte chunk column and an instruction covers RPI = 32 / CH whole rows: rows are
    // contiguous, so one cp.async instruction still touches full 32-byte sectors and few cache lines (a row per lane would be
    // bound by the one-line-per-cycle tag stage of L1TEX: 16 B/clk/SM measured, tools/copy_probe.cu), and per instruction a
    // lane needs just one shuffle (the row's element) and one multiply-add (the address): the shared-memory offset is a
    // per-lane constant plus an immediate.
    constexpr int RPI = 32 / CH;
    static_assert(RPI >= 1, "a row is at most 32 chunks of 16 bytes");
    const int r_sub = lane / CH, lch = lane - r_sub * CH;
    const bool lane_copies = r_sub < RPI;
    const uint32_t lane_off = (uint32_t)(r_sub * NS * 8 + lch * 16);  // offset of the lane's chunk inside the staged vectors
    if (warp >= kConsumerWarps && warp < kConsumerWarps + kSrcWarps)
    {
        // ---- source producer ------------------------------------------------------------------------------------
        constexpr int ITS = (16 + RPI - 1) / RPI;
        const unsigned char *ub = reinterpret_cast<const unsigned char *>(u) + lch * 16, *tb = reinterpret_cast<const unsigned char *>(P.tc_prev) + lch * 16;
        const int q = lane & 7;  // lanes 0-7 hold the elements of the u_old rows, lanes 8-15 the same elements for the T_old rows
        const int first = warp - kConsumerWarps;
        [[maybe_unused]] const uint64_t pol_first = PBTE_L2_HINTS ? l2_policy_evict_first() : 0;
        int e_next = first < NT ? __ldg(S.tile_desc + (size_t)(tbase + first) * rows * kTile + q) : -1;
        [[maybe_unused]] const double xc1 = P.slot_c1[s], xc2 = P.slot_c2[s];
        for (int i = first; i < NT; i += kSrcWarps)
        {
            const int slot = i % R, it = i / R;
            const int e = e_next;
            const int32_t *desc_g = S.tile_desc + (size_t)(tbase + i) * rows * kTile;
            if (i + kSrcWarps < NT) e_next = __ldg(desc_g + (size_t)kSrcWarps * rows * kTile + q);
            if constexpr (march_xsrc(N))
            {
                src_tile_regs<N>(u, P.tc_prev, e, desc_g, S.ctbc, rows, Fp, xc1, xc2, ring + (size_t)slot * slot_bytes, hdr_bytes, empty + slot, it > 0,
                                 (uint32_t)(it - 1) & 1u, full + slot);
                PBTE_STAMP(i, 8);
            }
            else
            {
            if (it > 0) mbar_wait(empty + slot, (it - 1) & 1);
            unsigned char *sl = ring + (size_t)slot * slot_bytes;
            PBTE_STAMP(i, 7);
            const uint32_t vb = smem_u32(sl + hdr_bytes) + lane_off;
#pragma unroll
            for (int k = 0; k < ITS; ++k)
            {
                const int row = k * RPI + r_sub;  // rows 0-7 u_old, 8-15 T_old of the tile's elements
                const int er = __shfl_sync(0xffffffffu, e, row & 7);
                if (lane_copies && row < 16)
                {
                    const unsigned char *src = (row < 8 ? ub : tb) + (size_t)(er < 0 ? 0 : er) * (NP * 8);
                    if ((PBTE_L2_HINTS & 1) && row < 8)
                        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(vb + (uint32_t)(k * RPI * NS * 8)), "l"(src),
                                     "r"(er >= 0 ? 16u : 0u), "l"(pol_first)
                                     : "memory");
                    else
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vb + (uint32_t)(k * RPI * NS * 8)), "l"(src), "r"(er >= 0 ? 16u : 0u) : "memory");
                }
            }
            if (lane < rows * 2) cp_async16(sl + lane * 16, reinterpret_cast<const unsigned char *>(desc_g) + lane * 16, 16u);
            {
                // wall temperatures of the 8 elements: Fp/2 chunks per row (at most 32 chunks in all)
                const int hc = Fp >> 1, row = lane / hc, ch = lane - row * hc;
                const int er = __shfl_sync(0xffffffffu, e, row & 7);
                if (lane < kTile * hc)
                    cp_async16(sl + rows * 32 + lane * 16, S.ctbc + (size_t)(er < 0 ? 0 : er) * Fp + 2 * ch, er >= 0 ? 16u : 0u);
            }
            cp_async_arrive(full + slot);
            PBTE_STAMP(i, 8);
            }
        }
    }
    else if (warp >= kConsumerWarps + kSrcWarps)
    {
        // ---- upwind-neighbour producer --------------------------------------------------------------------------
        constexpr int ITN = (32 + RPI - 1) / RPI;  // up to 4 inflow faces (32 rows) per pass
        const unsigned char *ub = reinterpret_cast<const unsigned char *>(u) + lch * 16;
        const int q = lane & 7, mq = lane >> 3;
        int wm = 0;  // watermark: tiles [0, wm) of this class are known to be stored
        const int first = warp - kConsumerWarps - kSrcWarps;
        const int32_t *d0 = S.tile_desc + (size_t)(tbase + first) * rows * kTile;
        int o_next = first < NT ? __ldg(d0 + (1 + F) * kTile) : 0, L_next = first < NT ? __ldg(d0 + (1 + F) * kTile + 3) : -1;
        int code_next = first < NT ? __ldg(d0 + (1 + mq) * kTile + q) : -1;  // rows 1..4 always exist (rows = 2 + F, F >= 3)
        for (int i = first; i < NT; i += kNbrWarps)
        {
            const int slot = i % R, it = i / R;
            const int o = o_next, dep = L_next, code0 = code_next;
            const int32_t *desc_g = S.tile_desc + (size_t)(tbase + i) * rows * kTile;
            if (i + kNbrWarps < NT)
            {
                const int32_t *dn = desc_g + (size_t)kNbrWarps * rows * kTile;
                o_next = __ldg(dn + (1 + F) * kTile);
                L_next = __ldg(dn + (1 + F) * kTile + 3);
                code_next = __ldg(dn + (1 + mq) * kTile + q);
            }
            if constexpr (!march_xnbr(N))
                if (it > 0) mbar_wait(empty + slot, (it - 1) & 1);
            PBTE_STAMP(i, 4);
            // every tile up to the last prerequisite must be stored: advance the watermark over the bitmap of finished tiles
#ifdef PBTE_NODEP  // timing experiment only (wrong results): how fast is the pipeline when no tile waits for another?
            wm = dep + 1;
#endif
            while (wm <= dep)
            {
                const unsigned w = *reinterpret_cast<volatile unsigned *>(s_done + (wm >> 5)) >> (wm & 31);
                const int ones = (~w == 0u) ? 32 : __ffs(~w) - 1;  // consecutive finished tiles starting at wm (zeros are shifted in)
                if (ones == 0) __nanosleep(32);
                else wm += min(ones, 32 - (wm & 31));
            }
            __threadfence_block();  // acquire: the stores of those tiles are visible to the copies issued below
            PBTE_STAMP(i, 5);
            const int nin = __popc((unsigned)s_inm[o]);
            unsigned char *sl = ring + (size_t)slot * slot_bytes;
            if constexpr (march_xnbr(N))
            {
                // rows into registers now, into the slot once it is free
                nbr_tile_regs<N>(u, code0, nin, desc_g, 1, sl, hdr_bytes, empty + slot, it > 0, (uint32_t)(it - 1) & 1u, full + slot);
                PBTE_STAMP(i, 6);
            }
            else
            {
            for (int m0 = 0; m0 < nin; m0 += 4)
            {
                // lane (mq, q) holds the source code of row mq*8 + q of this pass
                const int code = m0 == 0 ? code0 : ((m0 + mq < nin) ? __ldg(desc_g + (1 + m0 + mq) * kTile + q) : kSrcSkipCode);
                const int nrow = min(nin - m0, 4) * kTile;
                const uint32_t vb = smem_u32(sl + hdr_bytes) + (uint32_t)(((march_xsrc(N) ? 1 : 2) + m0) * kTile * NS * 8) + lane_off;
#pragma unroll
                for (int k = 0; k < ITN; ++k)
                {
                    if (k * RPI >= nrow) break;
                    const int row = k * RPI + r_sub;  // lane `row` holds the source of row `row` (face m0 + row / 8, element row % 8)
                    const int cr = __shfl_sync(0xffffffffu, code, row & 31);
                    if (lane_copies && row < nrow)
                    {
                        const unsigned char *src = ub + (size_t)(cr < 0 ? 0 : cr) * (NP * 8);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vb + (uint32_t)(k * RPI * NS * 8)), "l"(src), "r"(cr >= 0 ? 16u : 0u) : "memory");
                    }
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
        const int q = lane >> 2, r = lane & 3;
        const double c1 = P.slot_c1[s], c2 = P.slot_c2[s];
        const double *oph = S.h_smem ? s_h : S.oph + (size_t)s * noc * F * N;
        [[maybe_unused]] const uint64_t pol_last = PBTE_L2_HINTS ? l2_policy_evict_last() : 0;
        // One pass = T tiles (1 or 2) of one operator class whose ring slots are full: U_t = X_t P^T + sum_m Unbr_t,m Q_m^T.
        // One matrix block = KS k steps; per k step the B fragments of all MT row blocks (two per 16-byte load from the staged
        // operator) feed T * MT MMAs, the A fragments come from the staged vectors.  With T = 2 every operator fragment read from
        // shared memory is used twice: the B fragments are 57 of the 75 KB a tile moves through the shared-memory pipe, and that
        // pipe, not the FP64 tensor pipe, bounds the kernel (profiles/README.md).  The fragments of KC k steps are loaded
        // together into their own registers and only then multiplied: no register is rewritten while MMAs that read it are in
        // flight and every MMA finds its operands ready (interleaved load/MMA code ran at the MMA latency: tools/dmma_probe.cu).
        auto pass = [&](auto tc, int i0) {
            constexpr int T = decltype(tc)::value;
            constexpr int KC0 = (16 / (2 * MP)) < 1 ? 1 : (16 / (2 * MP) > KS ? KS : 16 / (2 * MP));
            constexpr int KC = KC0 / T < 1 ? 1 : KC0 / T;
            const unsigned char *sl[T];
            const int32_t *sd[T];
            const double *vec[T];
            int e[T];
            unsigned wall[T];
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
                sl[t] = ring + (size_t)((i0 + t) % R) * slot_bytes;
                sd[t] = reinterpret_cast<const int32_t *>(sl[t]);
                vec[t] = reinterpret_cast<const double *>(sl[t] + hdr_bytes) + (size_t)q * NS + r;
                e[t] = sd[t][q];
                wall[t] = 0;
            }
            const int o = sd[0][(1 + F) * kTile];
            const unsigned inm = (unsigned)s_inm[o];
            const int nin = __popc(inm);
            const double *op = ops + (size_t)s_opoff[o] * OPS;
            if (!STAGE)
            {
                // operators of many classes stay in global memory (L2): ask for all of this pass's fragments now, one cache line per
                // lane and round, so that the later matrix blocks find them in L1 instead of paying an L2 round trip each (C3 slice:
                // 30.1 -> 26.9 ms per sweep; asking one pass earlier, for the warp's next tile, evicts too much: 29.7 ms)
                const char *pb = reinterpret_cast<const char *>(op);
                const int bytes = (1 + nin) * OPS * (int)sizeof(double);
                for (int off = lane * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + off));
            }
            // what the descriptors say about the inflow faces: the wall faces of this lane's elements, and the faces where some
            // element of the pass has an interior upwind neighbour
            unsigned interior = 0;
            {
                // rows 1..kMaxIn of a descriptor always exist (rows = 2 + F, F >= 3): read them all at once, then vote
                int code[T][kMaxIn];
#pragma unroll
                for (int t = 0; t < T; ++t)
#pragma unroll
                    for (int m = 0; m < kMaxIn; ++m) code[t][m] = sd[t][(1 + m) * kTile + q];
#pragma unroll
                for (int m = 0; m < kMaxIn; ++m)
                {
                    bool any = false;
#pragma unroll
                    for (int t = 0; t < T; ++t)
                    {
                        wall[t] |= (unsigned)(code[t][m] == kSrcWallCode && m < nin) << m;
                        any = any || code[t][m] >= 0;
                    }
                    interior |= (unsigned)(__ballot_sync(0xffffffffu, any && m < nin) != 0u) << m;
                }
                for (int m = kMaxIn; m < nin; ++m)  // elements with more inflow faces than any supported shape has today
#pragma unroll
                    for (int t = 0; t < T; ++t)
                    {
                        const int c2m = sd[t][(1 + m) * kTile + q];
                        wall[t] |= (unsigned)(c2m == kSrcWallCode) << m;
                        interior |= (unsigned)(__ballot_sync(0xffffffffu, c2m >= 0) != 0u) << m;
                    }
            }
            double acc[T][MT][2];
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
                for (int I = 0; I < MT; ++I) acc[t][I][0] = acc[t][I][1] = 0.0;
            auto block = [&](size_t voff, bool source, const double *opb) {
                const double2 *bp = reinterpret_cast<const double2 *>(opb) + lane;
#pragma unroll
                for (int k0 = 0; k0 < KS; k0 += KC)
                {
                    double2 bq[KC][MP];
                    double xa[T][KC];
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        if (k0 + k < KS)
                        {
#pragma unroll
                            for (int Ip = 0; Ip < MP; ++Ip)
                            {
                                if constexpr (FRAG_B) bq[k][Ip] = bp[(Ip * KS + k0 + k) * 32];
                                else
                                {
                                    // compact row-major operator: this lane's B entries of row blocks 2Ip and 2Ip+1 are R[row][col]
                                    const int row = 16 * Ip + q, col = 4 * (k0 + k) + r;
                                    bq[k][Ip].x = (row < N && col < N) ? __ldg(opb + row * N + col) : 0.0;
                                    bq[k][Ip].y = (row + 8 < N && col < N) ? __ldg(opb + (row + 8) * N + col) : 0.0;
                                }
                            }
#pragma unroll
                            for (int t = 0; t < T; ++t)
                            {
                                const double *va = vec[t] + voff;
                                xa[t][k] = source ? fma(c2, va[4 * (k0 + k)], c1 * va[(size_t)kTile * NS + 4 * (k0 + k)]) : va[4 * (k0 + k)];
                            }
                        }
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        if (k0 + k < KS)
                        {
#pragma unroll
                            for (int I = 0; I < MT; ++I)
#pragma unroll
                                for (int t = 0; t < T; ++t) dmma_m8n8k4(acc[t][I][0], acc[t][I][1], xa[t][k], (I & 1) ? bq[k][I >> 1].y : bq[k][I >> 1].x);
                        }
                }
            };
            PBTE_STAMP(i0, 9);
            // source / pseudo-time term: X = c1 T_old + c2 u_old, formed here from vectors 0 and 1 of the slot or (march_xsrc) by the producer
            constexpr int kFaceVec = march_xsrc(N) ? 1 : 2;
#ifndef PBTE_PIPE_B
#define PBTE_PIPE_B 0
#endif
            // Experiment (off).  Operators read from global memory with few k steps (simplices): the B fragments of the NEXT matrix block
            // are requested before the MMAs of the current one, so that a block does not start with an exposed L1 / L2 round trip (tile
            // timeline of the C3 slice: 2000 cycles for 24 MMAs).  Same operands, same order of the MMAs: identical bits.  Measured
            // (profiles/r02_kernel_probes.txt, run y): 3.07 against 2.74 ms on the C3 slice - at the 64 registers two CTAs of 16 warps
            // leave a thread, the second fragment set spills and its copies cost more than the round trip.
            constexpr bool kPipeB = PBTE_PIPE_B != 0 && !FRAG_B && T == 1 && KS <= 4;
            if constexpr (kPipeB)
            {
                auto load_b = [&](const double *opb, double2(&bq)[KS][MP]) {
#pragma unroll
                    for (int k = 0; k < KS; ++k)
#pragma unroll
                        for (int Ip = 0; Ip < MP; ++Ip)
                        {
                            const int row = 16 * Ip + q, col = 4 * k + r;
                            bq[k][Ip].x = (row < N && col < N) ? __ldg(opb + row * N + col) : 0.0;
                            bq[k][Ip].y = (row + 8 < N && col < N) ? __ldg(opb + (row + 8) * N + col) : 0.0;
                        }
                };
                auto mma_block = [&](size_t voff, bool source, const double2(&bq)[KS][MP]) {
                    const double *va = vec[0] + voff;
                    double xa[KS];
#pragma unroll
                    for (int k = 0; k < KS; ++k) xa[k] = source ? fma(c2, va[4 * k], c1 * va[(size_t)kTile * NS + 4 * k]) : va[4 * k];
#pragma unroll
                    for (int k = 0; k < KS; ++k)
#pragma unroll
                        for (int I = 0; I < MT; ++I) dmma_m8n8k4(acc[0][I][0], acc[0][I][1], xa[k], (I & 1) ? bq[k][I >> 1].y : bq[k][I >> 1].x);
                };
                double2 bcur[KS][MP], bnext[KS][MP];
                load_b(op, bcur);
                unsigned rest = interior & ((1u << nin) - 1u);  // interior inflow faces still to do, in ascending order
                size_t voff = 0;
                bool source = !march_xsrc(N);
                while (true)
                {
                    const int mn = rest ? __ffs(rest) - 1 : -1;
                    if (mn >= 0) load_b(op + (size_t)(1 + mn) * OPS, bnext);
                    mma_block(voff, source, bcur);
                    if (mn < 0) break;
                    rest &= rest - 1;
                    voff = (size_t)(kFaceVec + mn) * kTile * NS;
                    source = false;
#pragma unroll
                    for (int k = 0; k < KS; ++k)
#pragma unroll
                        for (int Ip = 0; Ip < MP; ++Ip) bcur[k][Ip] = bnext[k][Ip];
                }
            }
            else
            {
            block(0, !march_xsrc(N), op);
            for (int m = 0; m < nin; ++m)
                if (interior >> m & 1u) block((size_t)(kFaceVec + m) * kTile * NS, false, op + (size_t)(1 + m) * OPS);
            }
#ifdef PBTE_TRACE
            if (acc[0][0][0] == 1.2345e300) PBTE_STAMP(i0, 15);  // the stamp below waits for the accumulators
#endif
            PBTE_STAMP(i0, 2);
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
                // isothermal walls on inflow faces: + T_bc h_j   (this lane owns element q, dofs 8I + 2r + {0,1})
                if (wall[t])
                {
                    const double *tbq = reinterpret_cast<const double *>(sl[t] + rows * 32) + (size_t)q * Fp;  // wall temperatures of element q
                    unsigned rest = inm;
                    for (int m = 0; rest; ++m, rest &= rest - 1)
                        if (wall[t] >> m & 1u)
                        {
                            const int j = __ffs(rest) - 1;
                            const double tw = tbq[j];
                            const double *hj = oph + ((size_t)o * F + j) * N;
#pragma unroll
                            for (int I = 0; I < MT; ++I)
                            {
                                const int d0 = 8 * I + 2 * r;
                                if (d0 < N) acc[t][I][0] = fma(tw, hj[d0], acc[t][I][0]);
                                if (d0 + 1 < N) acc[t][I][1] = fma(tw, hj[d0 + 1], acc[t][I][1]);
                            }
                        }
                }
            }
            __syncwarp();
            if (lane == 0)
#pragma unroll
                for (int t = 0; t < T; ++t) mbar_arrive(empty + (i0 + t) % R);  // every staged value of the tile has been consumed
#pragma unroll
            for (int t = 0; t < T; ++t)
                if (e[t] >= 0)
                {
                    double *dst = u + (size_t)e[t] * NP;
#pragma unroll
                    for (int I = 0; I < MT; ++I)
                    {
                        // 16-byte stores; an odd ndof has one padding column per row, which stays zero
                        const int d0 = 8 * I + 2 * r;
                        if (d0 < NP)
                        {
                            if (PBTE_L2_HINTS & 2) st_global_f64x2_hint(dst + d0, make_double2(acc[t][I][0], d0 + 1 < N ? acc[t][I][1] : 0.0), pol_last);
                            else *reinterpret_cast<double2 *>(dst + d0) = make_double2(acc[t][I][0], d0 + 1 < N ? acc[t][I][1] : 0.0);
                        }
                    }
                }
            PBTE_STAMP(i0, 10);
            __syncwarp();  // orders the stores of all lanes before lane 0's release
            if (lane == 0)
            {
                __threadfence_block();
#pragma unroll
                for (int t = 0; t < T; ++t) atomicOr(s_done + ((i0 + t) >> 5), 1u << ((i0 + t) & 31));
            }
            PBTE_STAMP(i0, 3);
        };
        if constexpr (kPairTiles)
        {
            // warp w takes the tile pairs (2j, 2j+1), j = w, w + consumers, ...  The pair goes through one pass when the tiles share
            // their operators and the second does not depend on the first (its last prerequisite, descriptor header word 3, lies
            // before the pair): then waiting for both cannot block the first.  Otherwise one tile at a time.
            const int32_t *hdr0 = S.tile_desc + ((size_t)tbase * rows + 1 + F) * kTile;  // header row of the class's first tile
            for (int j = warp; 2 * j < NT; j += kConsumerWarps)
            {
                const int i0 = 2 * j, it = i0 / R;  // R is even: both tiles are in the same round of the ring
                bool two = false;
                if (i0 + 1 < NT)
                {
                    const int32_t *h0 = hdr0 + (size_t)i0 * rows * kTile, *h1 = h0 + (size_t)rows * kTile;
                    two = __ldg(h1) == __ldg(h0) && __ldg(h1 + 3) < i0;
                }
                PBTE_STAMP(i0, 0);
                mbar_wait(full + i0 % R, it & 1);
                if (two) mbar_wait(full + (i0 + 1) % R, it & 1);
                PBTE_STAMP(i0, 1);
                if (two) pass(std::integral_constant<int, 2>{}, i0);
                else
                {
                    pass(std::integral_constant<int, 1>{}, i0);
                    if (i0 + 1 < NT)
                    {
                        mbar_wait(full + (i0 + 1) % R, it & 1);
                        pass(std::integral_constant<int, 1>{}, i0 + 1);
                    }
                }
            }
        }
        else
        {
            for (int i = warp; i < NT; i += kConsumerWarps)
            {
                PBTE_STAMP(i, 0);
                mbar_wait(full + i % R, (i / R) & 1);
                PBTE_STAMP(i, 1);
                pass(std::integral_constant<int, 1>{}, i);
            }
        }
    }
}

}  // namespace pbte_dev
