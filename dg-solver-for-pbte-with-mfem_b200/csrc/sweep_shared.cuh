// Shared-operator sweep ("marching" kernel): the fast path for meshes whose elements fall into few operator classes
// (csrc/shared_plan.hpp).  For an operator class o, direction k and band b the local system matrix A is the same for every
// element of the class, so the sweep update is applied through precomputed solution operators
//     P   = A^-1 M^T                         u_e  = P (c1 T_old + c2 u_old)
//     Q_f = -c_f A^-1 K_f                         + sum_{inflow f} Q_f u_nbr(f)
//     h_f = -c_f (C/omega) A^-1 g_f               + sum_{inflow wall f} T_bc(f) h_f
// built once per (class, slot) by build_operators_kernel (pivoted LU, same assembly as the per-element kernels).
//
// sweep_march_kernel: one CTA per slot (direction, band).  The CTA stages the slot's operators in shared memory with
// one TMA bulk copy, then marches through the wavefront levels of its upwind DAG: a warp takes one tile of 8 elements of
// one operator class and evaluates  U_tile(8 x n) = X(8 x n) P^T + sum_f Unbr_f(8 x n) Q_f^T  as FP64 tensor-core MMAs
// (mma.sync m8n8k4: A fragment = element data straight from the flux, B fragment = operator in fragment order from shared
// memory, accumulators in registers), adds the wall terms and writes the tile back in place.  Levels are separated by a
// CTA barrier only: no inter-CTA dependency exists, one launch per sweep.
// moments_from_flux_kernel then forms [Tc | Qc] by one coalesced streaming pass over the flux (fixed summation order).
#pragma once

#include "common.cuh"

namespace pbte_dev
{
constexpr int kTile = 8;             // elements per tile (must equal pbte_host::kTileElems)
constexpr int kMarchThreads = 256;   // 8 warps per marching CTA
constexpr int kSrcWallCode = -2;     // pbte_host::kSrcWall

struct DevMarch
{
    int noclass, F;
    const int32_t *slot_class;        // [nslots] upwind-DAG class of the slot
    const int32_t *cls_level_begin;   // [nclass+1] into level_tile_begin
    const int32_t *level_tile_begin;  // per class nlevels+1 tile offsets
    const int32_t *tile_oclass;       // [ntiles]
    const int32_t *tile_desc;         // [ntiles][1+F][8]
    const uint8_t *cls_inmask;        // [nclass][noclass] canonical inflow faces
    const int32_t *cls_op_off;        // [nclass][noclass] first operator of the class inside a slot's operator block (units of FR)
    const int32_t *cls_op_cnt;        // [nclass] operators per slot
    const unsigned long long *slot_op_base;  // [nslots] offset (doubles) of the slot's operator block in ops
    double *ops;                      // operators in MMA fragment order
    double *oph;                      // [nslots][noclass][F][n] wall vectors h_f
    const double *ctbc;               // [ne][F] wall temperature per canonical face
    // operator construction
    const int32_t *oc_rep;            // [noclass]
    const int32_t *oc_ksrc, *oc_gsrc; // [noclass][F] element*F + local face, or -1
    const uint8_t *cface;             // [ne][F] canonical -> local face
    const int32_t *cnbr;              // [ne][F] canonical face -> neighbour / PBTE_FACE_BOUNDARY / PBTE_FACE_NONE
};

// fragment geometry for ndof = N
template <int N>
struct Frag
{
    static constexpr int MT = (N + 7) / 8;   // 8-row blocks of the operator
    static constexpr int KS = (N + 3) / 4;   // k steps of 4
    static constexpr int FR = MT * KS * 32;  // doubles per operator in fragment order
};
__host__ __device__ inline int frag_doubles(int n) { return ((n + 7) / 8) * ((n + 3) / 4) * 32; }

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

// out[(I*KS + kk)*32 + lane] = R[8I + lane/4][4kk + lane%4] (zero padded): the B fragments of mma m8n8k4 for R^T
__device__ __forceinline__ void store_fragment_order(double *out, const double *R, int n)
{
    const int KS = (n + 3) / 4, FR = frag_doubles(n);
    for (int idx = threadIdx.x; idx < FR; idx += blockDim.x)
    {
        const int lane = idx & 31, blk = idx >> 5, kk = blk % KS, I = blk / KS;
        const int row = 8 * I + (lane >> 2), col = 4 * kk + (lane & 3);
        out[idx] = (row < n && col < n) ? R[row + (size_t)col * n] : 0.0;
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
    const int e = S.oc_rep[o];
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
            const int f = S.cface[(size_t)e * F + t];
            if (S.cnbr[(size_t)e * F + t] != PBTE_FACE_NONE) a = dot_seq(P.fnormal + ((size_t)e * F + f) * 3, om, dim);
        }
        s_af[t] = a;
    }
    if (t == 0) s_fail = 0;
    __syncthreads();

    const double *Me = P.mass + (size_t)e * ms, *Ge = P.stiff + (size_t)e * dim * ms, *Fe = P.fmass + (size_t)e * F * ms;
    for (int idx = t; idx < n * n; idx += T)
    {
        double b = 0.0;
        for (int d = 0; d < dim; ++d) b = fma(-om[d], Ge[(size_t)d * ms + idx], b);
        for (int j = 0; j < F; ++j)
            if (s_af[j] > 0.0) b = fma(s_af[j], Fe[(size_t)S.cface[(size_t)e * F + j] * ms + idx], b);
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

    double *ops = S.ops + S.slot_op_base[s] + (size_t)S.cls_op_off[(size_t)c * S.noclass + o] * FR;
    // P = A^-1 M^T
    {
        const double *MTe = P.massT + (size_t)e * ms;
        for (int idx = t; idx < n * n; idx += T) R[idx] = MTe[idx];
        __syncthreads();
        lu_solve_columns(A, R, s_piv, n, n);
        __syncthreads();
        store_fragment_order(ops, R, n);
        __syncthreads();
    }
    int m = 1;
    for (int j = 0; j < F; ++j)
    {
        if (!(inm >> j & 1u)) continue;
        const double cin = 0.5 * v * (s_af[j] - fabs(s_af[j]));  // c_f (src/PBTESolver.cpp:282)
        const int ks = S.oc_ksrc[(size_t)o * F + j], gs = S.oc_gsrc[(size_t)o * F + j];
        double *out = ops + (size_t)m * FR;
        if (ks >= 0)
        {
            // Q_j = -c_j A^-1 K_j
            const double *K = P.fcoup + (size_t)ks * ms;
            for (int idx = t; idx < n * n; idx += T) R[idx] = -cin * K[idx];
            __syncthreads();
            lu_solve_columns(A, R, s_piv, n, n);
            __syncthreads();
            store_fragment_order(out, R, n);
        }
        else
            for (int idx = t; idx < FR; idx += T) out[idx] = 0.0;  // no element of the class has an interior face j
        __syncthreads();
        if (gs >= 0)
        {
            // h_j = -c_j (C/omega) A^-1 g_j
            for (int i = t; i < n; i += T) R[i] = -cin * cb * P.fint[(size_t)gs * n + i];
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
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One TMA bulk copy global -> shared of `bytes` (multiple of 16) issued by one thread, completion on an mbarrier.
__device__ __forceinline__ void bulk_stage(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
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

// grid = nslots CTAs of kMarchThreads threads; dynamic shared memory = the slot's operator block when STAGE
template <int N, bool STAGE>
__global__ void __launch_bounds__(kMarchThreads, 2) sweep_march_kernel(DevProblem P, DevMarch S)
{
    using G = Frag<N>;
    constexpr int MT = G::MT, KS = G::KS, FR = G::FR;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    const int s = blockIdx.x;
    const int c = S.slot_class[s];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int F = S.F, rows = 1 + F;
    const double *ops = S.ops + S.slot_op_base[s];
    if (STAGE)
    {
        bulk_stage(sm_raw, ops, (uint32_t)S.cls_op_cnt[c] * (uint32_t)(FR * sizeof(double)), &s_bar);
        ops = reinterpret_cast<const double *>(sm_raw);
    }
    const int lb = S.cls_level_begin[c], nlev = S.cls_level_begin[c + 1] - lb - 1;
    const int32_t *ltb = S.level_tile_begin + lb;
    const uint8_t *inmask = S.cls_inmask + (size_t)c * S.noclass;
    const int32_t *opoff = S.cls_op_off + (size_t)c * S.noclass;
    const double c1 = P.slot_c1[s], c2 = P.slot_c2[s];
    double *u = P.flux + (size_t)s * P.ne * N;
    const double *oph = S.oph + (size_t)s * S.noclass * F * N;
    const int q = lane >> 2, r = lane & 3;

    for (int L = 0; L < nlev; ++L)
    {
        const int t1 = ltb[L + 1];
        for (int t = ltb[L] + warp; t < t1; t += nwarps)
        {
            const int32_t *desc = S.tile_desc + (size_t)t * rows * kTile;
            const int e = __ldg(desc + q);
            const int o = __ldg(S.tile_oclass + t);
            const unsigned inm = inmask[o];
            const double *op = ops + (size_t)opoff[o] * FR;
            double acc[MT][2];
#pragma unroll
            for (int I = 0; I < MT; ++I) acc[I][0] = acc[I][1] = 0.0;

            // source / pseudo-time term: X = c1 T_old + c2 u_old
            {
                double x[KS];
                const double *ue = u + (size_t)(e < 0 ? 0 : e) * N, *te = P.tc_prev + (size_t)(e < 0 ? 0 : e) * N;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                {
                    const int row = 4 * kk + r;
                    x[kk] = (e >= 0 && row < N) ? fma(c2, __ldcg(ue + row), c1 * __ldg(te + row)) : 0.0;
                }
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
#pragma unroll
                    for (int I = 0; I < MT; ++I)
                    {
                        const double b = STAGE ? op[(I * KS + kk) * 32 + lane] : __ldg(op + (I * KS + kk) * 32 + lane);
                        dmma_m8n8k4(acc[I][0], acc[I][1], x[kk], b);
                    }
            }
            // upwind couplings of the inflow faces
            int m = 1;
            unsigned wall = 0;
            for (unsigned rest = inm; rest; rest &= rest - 1, ++m)
            {
                const int j = __ffs(rest) - 1;
                const int code = __ldg(desc + (1 + j) * kTile + q);
                if (code == kSrcWallCode) wall |= 1u << j;
                if (__ballot_sync(0xffffffffu, code >= 0) == 0u) continue;
                double x[KS];
                const double *un = u + (size_t)(code < 0 ? 0 : code) * N;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                {
                    const int row = 4 * kk + r;
                    x[kk] = (code >= 0 && row < N) ? __ldcg(un + row) : 0.0;
                }
                const double *oq = op + (size_t)m * FR;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
#pragma unroll
                    for (int I = 0; I < MT; ++I)
                    {
                        const double b = STAGE ? oq[(I * KS + kk) * 32 + lane] : __ldg(oq + (I * KS + kk) * 32 + lane);
                        dmma_m8n8k4(acc[I][0], acc[I][1], x[kk], b);
                    }
            }
            // isothermal walls on inflow faces: + T_bc h_j   (this lane owns element q, dofs 8I + 2r + {0,1})
            for (; wall; wall &= wall - 1)
            {
                const int j = __ffs(wall) - 1;
                const double tb = __ldg(S.ctbc + (size_t)e * F + j);
                const double *hj = oph + ((size_t)o * F + j) * N;
#pragma unroll
                for (int I = 0; I < MT; ++I)
                {
                    const int d0 = 8 * I + 2 * r;
                    if (d0 < N) acc[I][0] = fma(tb, __ldg(hj + d0), acc[I][0]);
                    if (d0 + 1 < N) acc[I][1] = fma(tb, __ldg(hj + d0 + 1), acc[I][1]);
                }
            }
            if (e >= 0)
            {
                double *dst = u + (size_t)e * N;
#pragma unroll
                for (int I = 0; I < MT; ++I)
                {
                    const int d0 = 8 * I + 2 * r;
                    if (d0 < N) dst[d0] = acc[I][0];
                    if (d0 + 1 < N) dst[d0 + 1] = acc[I][1];
                }
            }
        }
        __syncthreads();  // level L complete (its stores are visible to the CTA) before level L+1 reads it
    }
}

// mom[m][x] = sum_s fac_m[s] * flux[s][x], x = (element, dof) flattened; slots in ascending order (fixed summation order)
__global__ void __launch_bounds__(256) moments_from_flux_kernel(DevProblem P, double *__restrict__ mom)
{
    const size_t block = (size_t)P.ne * P.n;
    const size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (x >= block) return;
    const int dim = P.dim, ns = P.nslots;
    double t = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0;
    const double *u = P.flux + x;
    int s = 0;
    for (; s + 8 <= ns; s += 8)
    {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(u + (size_t)(s + k) * block);
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            t = fma(P.slot_mf[s + k], v[k], t);
            q0 = fma(P.slot_qf[3 * (s + k)], v[k], q0);
            q1 = fma(P.slot_qf[3 * (s + k) + 1], v[k], q1);
            q2 = fma(P.slot_qf[3 * (s + k) + 2], v[k], q2);
        }
    }
    for (; s < ns; ++s)
    {
        const double v = __ldcs(u + (size_t)s * block);
        t = fma(P.slot_mf[s], v, t);
        q0 = fma(P.slot_qf[3 * s], v, q0);
        q1 = fma(P.slot_qf[3 * s + 1], v, q1);
        q2 = fma(P.slot_qf[3 * s + 2], v, q2);
    }
    mom[x] = t;
    mom[block + x] = q0;
    if (dim > 1) mom[2 * block + x] = q1;
    if (dim > 2) mom[3 * block + x] = q2;
}

}  // namespace pbte_dev
