// Warp-register sweep kernel (ndof <= 32): the hot kernel of the library.
//
// One CTA per (schedule class, element) of the current wavefront level.  The element's local matrices
// (M, M^T, G_d, the F_f of outflow faces, the K_f of inflow faces) are staged once in shared memory and
// reused by every (direction, band) system of the class.  One warp owns one system at a time; lane i
// holds row i of the local matrix in registers:
//   B_k   = -sum_d Omega_d G_d + sum_{out f} a_f F_f      assembled once per direction, kept in registers
//   A     = nu_max M + v_band B_k                          per band
//   r     = M^T (c1 T_old + c2 u_old) - sum_{in f} c_f K_f u_nbr - boundary terms
//   LU with implicit row pivoting (rows never move: the pivot lane broadcasts its row through a
//   double-buffered shared-memory line, the right-hand side rides along as an extra column), back
//   substitution through warp shuffles, solution written back in place, band/angle-weighted moments
//   accumulated in registers and reduced across the CTA at the end.
// With the FullLU policy the pivot lane sequence of every system is cached (ndof bytes per system) by the
// first sweep and replayed afterwards, which removes the pivot search from the steady state.
#pragma once

#include "common.cuh"

namespace pbte_dev
{
constexpr int kWarpsPerCta = 4;

struct WarpKernelShape
{
    int nF_max, nK_max;  // staged face-mass / coupling matrices per CTA (maxima over all items)
};

// PIV: 0 = search pivots, 1 = search and record, 2 = replay recorded pivots
// Two ways to launch it:
//   one launch per wavefront level   grid = items of the level, item_count = grid size, done = nullptr;
//   ONE launch for the whole sweep   grid = resident CTAs, item_count = all items (level-major), done = [nclass][ne] flags zeroed before
//                                    the launch: a CTA takes the items blockIdx.x, + gridDim.x, ... in order, waits until the upwind
//                                    neighbours of its item are flagged (acquire at GPU scope), and flags its own item after the
//                                    flux is stored (release).  An item waits only for items earlier in the list, every CTA is
//                                    resident and works in order: no deadlock, no launch per level, no idle tail per level.
template <int N, int DIM, int PIV>
#ifndef PBTE_WARP_MINBLOCKS
#define PBTE_WARP_MINBLOCKS 1  // experiment: ask for more CTAs per SM (fewer registers per thread)
#endif
__global__ void __launch_bounds__(kWarpsPerCta * 32, PBTE_WARP_MINBLOCKS) sweep_warp_kernel(DevProblem P, const int2 *__restrict__ items, int item_begin, int item_count,
                                                                      int32_t *done, WarpKernelShape shape)
{
    constexpr int NN = N * N;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double sm[];
    if (*P.stop) return;  // the iteration before this one met the stop test (csrc/tail.cuh)
    const int F = P.F;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool singular = false;
    for (int item = blockIdx.x; item < item_count; item += gridDim.x)
    {
    const int2 it = items[item_begin + item];
    const int c = it.x, e = it.y;
    const int s0 = P.cls_slot_begin[c], s1 = P.cls_slot_begin[c + 1];
    const int ndir_c = (s1 - s0) / P.nband;
    const unsigned outm = P.cls_out[(size_t)c * P.ne + e], inm = P.cls_in[(size_t)c * P.ne + e];

    // ---- shared memory carve-up -----------------------------------------------------------------
    double *sM = sm;                       // [NN]
    double *sMT = sM + NN;                 // [NN]
    double *sG = sMT + NN;                 // [DIM][NN]
    double *sF = sG + DIM * NN;            // [nF_max][NN]
    double *sK = sF + shape.nF_max * NN;   // [nK_max][NN]
    double *sScr = sK + shape.nK_max * NN; // per warp: line[2][N+2] | vec[(1+F)][N]
    const int scr_stride = 2 * (N + 2) + (1 + kMaxFaces) * N;
    double *line = sScr + warp * scr_stride;
    double *vec = line + 2 * (N + 2);
    __shared__ signed char s_fslot[kMaxFaces], s_kslot[kMaxFaces];
    __shared__ double s_red[kWarpsPerCta][1 + kMaxDim][32];

    if (threadIdx.x == 0)
    {
        int nf = 0, nk = 0;
        for (int f = 0; f < kMaxFaces; ++f)
        {
            s_fslot[f] = (f < F && (outm >> f & 1u)) ? (signed char)nf++ : (signed char)-1;
            s_kslot[f] = (f < F && (inm >> f & 1u)) ? (signed char)nk++ : (signed char)-1;
        }
    }
    if (done && threadIdx.x < F && (inm >> threadIdx.x & 1u))
    {
        // whole-sweep launch: the upwind neighbours of this item must be stored (their CTAs flagged them with a release)
        const int nb = P.fnbr[(size_t)e * F + threadIdx.x];
        if (nb >= 0)
        {
            const int32_t *flag = done + (size_t)c * P.ne + nb;
            int v;
            do
            {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (!v) __nanosleep(100);
            } while (!v);
        }
    }
    __syncthreads();
    {
        const size_t ms = P.ms;
        const double *Me = P.mass + (size_t)e * ms, *MTe = P.massT + (size_t)e * ms;
        const double *Ge = P.stiff + (size_t)e * DIM * ms;
        for (int i = threadIdx.x; i < NN; i += blockDim.x)
        {
            sM[i] = Me[i];
            sMT[i] = MTe[i];
#pragma unroll
            for (int d = 0; d < DIM; ++d) sG[d * NN + i] = Ge[(size_t)d * ms + i];
        }
        for (int f = 0; f < F; ++f)
        {
            if (s_fslot[f] >= 0)
            {
                const double *src = P.fmass + ((size_t)e * F + f) * ms;
                double *dst = sF + s_fslot[f] * NN;
                for (int i = threadIdx.x; i < NN; i += blockDim.x) dst[i] = src[i];
            }
            if (s_kslot[f] >= 0)
            {
                const double *src = P.fcoup + ((size_t)e * F + f) * ms;
                double *dst = sK + s_kslot[f] * NN;
                for (int i = threadIdx.x; i < NN; i += blockDim.x) dst[i] = src[i];
            }
        }
    }
    __syncthreads();

    const bool active = lane < N;
    const int row = active ? lane : 0;  // idle lanes shadow row 0 (their results are discarded)
    const double T_old = P.tc_prev[(size_t)e * P.np + row];
    double accT = 0.0, accQ[kMaxDim] = {0.0, 0.0, 0.0};

    for (int kd = warp; kd < ndir_c; kd += kWarpsPerCta)
    {
        const int sd = s0 + kd * P.nband;  // first slot of this direction
        const double om[3] = {P.slot_dir[3 * sd], P.slot_dir[3 * sd + 1], P.slot_dir[3 * sd + 2]};
        double af[kMaxFaces];
#pragma unroll
        for (int f = 0; f < kMaxFaces; ++f)
            af[f] = (f < F && P.fnbr[(size_t)e * F + f] != PBTE_FACE_NONE) ? dot_seq(P.fnormal + ((size_t)e * F + f) * 3, om, DIM) : 0.0;

        // B_k row in registers
        double bk[N];
#pragma unroll
        for (int j = 0; j < N; ++j)
        {
            double x = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) x = fma(-om[d], sG[d * NN + row + j * N], x);
            bk[j] = x;
        }
#pragma unroll
        for (int f = 0; f < kMaxFaces; ++f)
        {
            if (f < F && af[f] > 0.0)  // warp-uniform
            {
                const double *Fm = sF + s_fslot[f] * NN;
#pragma unroll
                for (int j = 0; j < N; ++j) bk[j] = fma(af[f], Fm[row + j * N], bk[j]);
            }
        }

        for (int band = 0; band < P.nband; ++band)
        {
            const int s = sd + band;
            const double v = P.slot_v[s], c1 = P.slot_c1[s], c2 = P.slot_c2[s], cb = P.slot_cb[s];
            const size_t uoff = ((size_t)s * P.ne + e) * P.np;

            // stage the source vector and the upwind neighbour fluxes (lane j holds entry j)
            if (active) vec[lane] = c1 * T_old + c2 * P.flux[uoff + lane];
#pragma unroll
            for (int f = 0; f < kMaxFaces; ++f)
            {
                if (f < F && af[f] < 0.0)
                {
                    const int nb = P.fnbr[(size_t)e * F + f];
                    if (nb >= 0 && active) vec[(1 + f) * N + lane] = __ldcg(P.flux + ((size_t)s * P.ne + nb) * P.np + lane);  // through L2: written by another SM during this launch
                }
            }
            __syncwarp();

            // right-hand side of row `row`
            double r = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) r = fma(sMT[row + j * N], vec[j], r);
#pragma unroll
            for (int f = 0; f < kMaxFaces; ++f)
            {
                if (f < F && af[f] < 0.0)
                {
                    const double cin = 0.5 * v * (af[f] - fabs(af[f]));
                    const int nb = P.fnbr[(size_t)e * F + f];
                    if (nb >= 0)
                    {
                        const double *K = sK + s_kslot[f] * NN;
                        const double *z = vec + (1 + f) * N;
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < N; ++j) acc = fma(K[row + j * N], z[j], acc);
                        r = fma(-cin, acc, r);
                    }
                    else r = fma(-cin * cb * P.ftbc[(size_t)e * F + f], P.fint[((size_t)e * F + f) * N + row], r);
                }
            }

            // local matrix row
            double a[N];
#pragma unroll
            for (int j = 0; j < N; ++j) a[j] = fma(v, bk[j], P.nu_max * sM[row + j * N]);

            // ---- LU, implicit pivoting -----------------------------------------------------------
            int mystep = active ? N : -1;  // N = not yet used as pivot; idle lanes never compete
            double myinv = 0.0;
            uint8_t *pv = P.pivots ? P.pivots + uoff : nullptr;
            int rec = (PIV == 2 && active) ? (int)pv[lane] : 0;  // lane j holds the pivot lane of step j
#pragma unroll
            for (int k = 0; k < N; ++k)
            {
                int p;
                if (PIV == 2) p = __shfl_sync(FULL, rec, k);
                else
                {
                    const double mag = fabs(a[k]);
                    const bool cand = mystep == N;
                    const int hi = cand ? __double2hiint(mag) : -1;
                    const int mh = __reduce_max_sync(FULL, hi);
                    const unsigned lo = (cand && hi == mh) ? (unsigned)__double2loint(mag) : 0u;
                    const unsigned ml = __reduce_max_sync(FULL, lo);
                    const unsigned win = __ballot_sync(FULL, cand && hi == mh && lo == ml);
                    p = __ffs(win) - 1;
                    if (mh == 0 && ml == 0u) singular = true;
                    if (PIV == 1 && lane == k) rec = p;
                }
                double *ln = line + (k & 1) * (N + 2);
                if (lane == p)
                {
                    myinv = 1.0 / a[k];
                    mystep = k;
                    ln[k] = myinv;
#pragma unroll
                    for (int j = k + 1; j < N; ++j) ln[j] = a[j];
                    ln[N] = r;
                }
                __syncwarp();
                if (mystep == N)
                {
                    const double l = a[k] * ln[k];
                    a[k] = l;
#pragma unroll
                    for (int j = k + 1; j < N; ++j) a[j] = fma(-l, ln[j], a[j]);
                    r = fma(-l, ln[N], r);
                }
            }
            if (PIV == 1 && active) pv[lane] = (uint8_t)rec;

            // ---- back substitution: U row of step k lives in the lane that was its pivot ----------
            double x = 0.0;
#pragma unroll
            for (int k = N - 1; k >= 0; --k)
            {
                const unsigned own = __ballot_sync(FULL, mystep == k);
                const double xk = __shfl_sync(FULL, r * myinv, __ffs(own) - 1);
                if (mystep < k && mystep >= 0) r = fma(-a[k], xk, r);
                if (lane == k) x = xk;
            }
            __syncwarp();  // vec / line are reused by the next system

            if (active)
            {
                P.flux[uoff + lane] = x;
                accT = fma(P.slot_mf[s], x, accT);
#pragma unroll
                for (int d = 0; d < DIM; ++d) accQ[d] = fma(P.slot_qf[3 * s + d], x, accQ[d]);
            }
        }
    }

    // ---- moments of this (class, element): fixed-order sum over the warps ----------------------------
    s_red[warp][0][lane] = accT;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s_red[warp][1 + d][lane] = accQ[d];
    __syncthreads();
    for (int idx = threadIdx.x; idx < (1 + DIM) * 32; idx += blockDim.x)
    {
        const int m = idx >> 5, i = idx & 31;
        if (i < N)
        {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < kWarpsPerCta; ++w) sum += s_red[w][m][i];
            P.partial[(((size_t)c * (1 + DIM) + m) * P.ne + e) * N + i] = sum;
        }
    }
    // every warp's flux rows of this item are stored before the item is flagged (and before the next item reuses shared memory)
    __syncthreads();
    if (done && threadIdx.x == 0)
    {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(done + (size_t)c * P.ne + e), "r"(1) : "memory");
    }
    }  // items of this CTA
    if (singular && lane == 0) atomicCAS(P.error_flag, 0, PBTE_ERR_SINGULAR);
}

inline size_t sweep_warp_smem(int n, int dim, WarpKernelShape sh)
{
    const size_t nn = (size_t)n * n;
    return sizeof(double) * ((2 + dim + sh.nF_max + sh.nK_max) * nn + (size_t)kWarpsPerCta * (2 * (n + 2) + (1 + kMaxFaces) * n));
}

}  // namespace pbte_dev
