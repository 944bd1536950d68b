// Generic sweep kernel: any ndof, matrices read through the read-only cache, the local system held in
// shared memory and factored cooperatively by the CTA (row-pivoted LU, MFEM LUFactors operation order:
// multiply by the reciprocal pivot, eliminate column by column).  It is the functional baseline for
// ndof > 32 and a cross-check for the warp-register kernel; one CTA per (schedule class, element).
#pragma once

#include "common.cuh"

namespace pbte_dev
{
// dynamic shared memory: A[n*n] | b[n] | vec[(1+F)*n] | coef (3*F doubles) | ints
__global__ void __launch_bounds__(128) sweep_generic_kernel(DevProblem P, const int2 *__restrict__ items, int item_begin)
{
    extern __shared__ double sm[];
    if (*P.stop) return;  // the iteration before this one met the stop test (csrc/tail.cuh)
    const int n = P.n, F = P.F, dim = P.dim, T = blockDim.x, t = threadIdx.x;
    double *A = sm;
    double *b = A + (size_t)n * n;
    double *vec = b + n;                 // [(1+F)][n]  source vector and neighbour fluxes
    double *coef = vec + (size_t)(1 + F) * n;  // out[F], in[F], bdr[F]
    __shared__ int s_piv;
    __shared__ int s_fail;

    const int2 it = items[item_begin + blockIdx.x];
    const int c = it.x, e = it.y;
    const int s0 = P.cls_slot_begin[c], s1 = P.cls_slot_begin[c + 1];
    const size_t ms = P.ms;
    const double *Me = P.mass + (size_t)e * ms;
    const double *MTe = P.massT + (size_t)e * ms;
    const double *Ge = P.stiff + (size_t)e * dim * ms;
    const double *Fe = P.fmass + (size_t)e * F * ms;
    const double *Ke = P.fcoup + (size_t)e * F * ms;
    const int32_t *nbr = P.fnbr + (size_t)e * F;

    double accT = 0.0, accQ[kMaxDim] = {0.0, 0.0, 0.0};
    if (t == 0) s_fail = 0;

    for (int s = s0; s < s1; ++s)
    {
        const double om[3] = {P.slot_dir[3 * s], P.slot_dir[3 * s + 1], P.slot_dir[3 * s + 2]};
        const double v = P.slot_v[s], c1 = P.slot_c1[s], c2 = P.slot_c2[s], cb = P.slot_cb[s];
        __syncthreads();  // previous system fully consumed
        if (t < F)
        {
            double outc = 0.0, inc = 0.0, bdr = 0.0;
            if (nbr[t] != PBTE_FACE_NONE)
            {
                const double a = dot_seq(P.fnormal + ((size_t)e * F + t) * 3, om, dim);
                if (a > 0.0) outc = v * a;
                const double cin = 0.5 * v * (a - fabs(a));  // = v*a on inflow faces, 0 otherwise
                if (nbr[t] >= 0) inc = cin;
                else bdr = -cin * cb * P.ftbc[(size_t)e * F + t];
            }
            coef[t] = outc;
            coef[F + t] = inc;
            coef[2 * F + t] = bdr;
        }
        // source vector c1*T_old + c2*u_old and the upwind neighbour fluxes
        for (int i = t; i < n; i += T)
            vec[i] = c1 * P.tc_prev[(size_t)e * P.np + i] + c2 * P.flux[((size_t)s * P.ne + e) * P.np + i];
        for (int f = 0; f < F; ++f)
        {
            const int nb = nbr[f];
            if (nb >= 0)
                for (int i = t; i < n; i += T) vec[(size_t)(1 + f) * n + i] = P.flux[((size_t)s * P.ne + nb) * P.np + i];
        }
        __syncthreads();
        // A = nu_max*M - v*sum_d Omega_d G_d + sum_{outflow f} v*a_f F_f
        for (int idx = t; idx < n * n; idx += T)
        {
            double a = P.nu_max * Me[idx];
            for (int d = 0; d < dim; ++d) a += (-v * om[d]) * Ge[(size_t)d * ms + idx];
            for (int f = 0; f < F; ++f)
                if (coef[f] != 0.0) a += coef[f] * Fe[(size_t)f * ms + idx];
            A[idx] = a;
        }
        // rhs
        for (int i = t; i < n; i += T)
        {
            double r = 0.0;
            for (int j = 0; j < n; ++j) r += MTe[i + (size_t)j * n] * vec[j];
            for (int f = 0; f < F; ++f)
            {
                if (coef[F + f] != 0.0)
                {
                    double acc = 0.0;
                    const double *K = Ke + (size_t)f * ms;
                    const double *z = vec + (size_t)(1 + f) * n;
                    for (int j = 0; j < n; ++j) acc += K[i + (size_t)j * n] * z[j];
                    r -= coef[F + f] * acc;
                }
                if (coef[2 * F + f] != 0.0) r += coef[2 * F + f] * P.fint[((size_t)e * F + f) * n + i];
            }
            b[i] = r;
        }
        __syncthreads();
        // LU with row pivoting, right-hand side carried along
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
                s_piv = piv;
                if (best == 0.0) s_fail = 1;
            }
            __syncthreads();
            const int piv = s_piv;
            if (piv != k)
            {
                for (int j = t; j <= n; j += T)
                {
                    double *x = j < n ? &A[k + (size_t)j * n] : &b[k];
                    double *y = j < n ? &A[piv + (size_t)j * n] : &b[piv];
                    const double tmp = *x;
                    *x = *y;
                    *y = tmp;
                }
                __syncthreads();
            }
            const double inv = 1.0 / A[k + (size_t)k * n];
            __syncthreads();
            for (int i = k + 1 + t; i < n; i += T) A[i + (size_t)k * n] *= inv;
            __syncthreads();
            const int w = n - k - 1;
            for (int idx = t; idx < w * (w + 1); idx += T)
            {
                const int i = k + 1 + idx % w, jj = idx / w;  // jj == w -> rhs column
                const double l = A[i + (size_t)k * n];
                if (jj < w) A[i + (size_t)(k + 1 + jj) * n] -= l * A[k + (size_t)(k + 1 + jj) * n];
                else b[i] -= l * b[k];
            }
            __syncthreads();
        }
        // back substitution (column oriented)
        for (int j = n - 1; j >= 0; --j)
        {
            if (t == 0) b[j] /= A[j + (size_t)j * n];
            __syncthreads();
            const double xj = b[j];
            for (int i = t; i < j; i += T) b[i] -= A[i + (size_t)j * n] * xj;
            __syncthreads();
        }
        for (int i = t; i < n; i += T)
        {
            const double x = b[i];
            P.flux[((size_t)s * P.ne + e) * P.np + i] = x;
            if (i == t)  // thread i owns dof i when n <= blockDim; larger n handled below
            {
                accT += P.slot_mf[s] * x;
                for (int d = 0; d < dim; ++d) accQ[d] += P.slot_qf[3 * s + d] * x;
            }
        }
        if (n > T)
        {
            // rare path (ndof > 128): accumulate straight into the partial buffer
            __syncthreads();
            for (int i = t + T; i < n; i += T)
            {
                double *pp = P.partial + ((size_t)c * (1 + dim)) * P.ne * n + (size_t)e * n + i;
                const double x = b[i];
                if (s == s0) { pp[0] = 0.0; for (int d = 0; d < dim; ++d) pp[(size_t)(1 + d) * P.ne * n] = 0.0; }
                pp[0] += P.slot_mf[s] * x;
                for (int d = 0; d < dim; ++d) pp[(size_t)(1 + d) * P.ne * n] += P.slot_qf[3 * s + d] * x;
            }
        }
    }
    if (t < n)
    {
        double *pp = P.partial + ((size_t)c * (1 + dim)) * P.ne * n + (size_t)e * n + t;
        pp[0] = accT;
        for (int d = 0; d < dim; ++d) pp[(size_t)(1 + d) * P.ne * n] = accQ[d];
    }
    __syncthreads();
    if (t == 0 && s_fail) atomicCAS(P.error_flag, 0, PBTE_ERR_SINGULAR);
}

inline size_t sweep_generic_smem(int n, int F) { return sizeof(double) * ((size_t)n * n + n + (size_t)(1 + F) * n + 3 * F); }

}  // namespace pbte_dev
