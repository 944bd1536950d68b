// MacroscopicQuantities on the device, stand-alone calls (pbte_cuda_macro_*): block accumulation, cell integrals, residual.
// Sums over the dof index use warp-shuffle segmented reductions; every reduction has a fixed order (bit-reproducible).
// Inside the source iteration the same quantities come from iteration_tail_kernel (tail.cuh).
#pragma once

#include "common.cuh"

namespace pbte_dev
{
// Tc += f*x, Qc_d += f*v*Omega_d*x for one coefficient block x [ne][n]   (MacroscopicQuantities.cpp:104-128)
__global__ void accumulate_block_kernel(size_t block /*ne*n*/, int dim, double f, double q0, double q1, double q2,
                                        const double *__restrict__ x, double *__restrict__ mom)
{
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= block) return;
    const double v = x[idx];
    mom[idx] += f * v;
    mom[block + idx] += q0 * v;
    if (dim > 1) mom[2 * block + idx] += q1 * v;
    if (dim > 2) mom[3 * block + idx] += q2 * v;
}

// Segmented warp reduction: `seg` consecutive lanes (power of two, >= n) share one (moment, element) pair.
// cell[m][e] = sum_i mom[m][e][i] * basis[e][i]          (MacroscopicQuantities.cpp:130-157)
__global__ void finalize_kernel(int ne, int n, int nmom, int seg, const double *__restrict__ mom, const double *__restrict__ basis,
                                double *__restrict__ cell /*[nmom][ne]*/)
{
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t pair = gid / seg;
    const int i0 = (int)(gid % seg);
    const bool valid = pair < (size_t)nmom * ne;
    double s = 0.0;
    if (valid)
    {
        const size_t e = pair % ne;
        for (int i = i0; i < n; i += seg) s += mom[pair * n + i] * basis[e * n + i];
    }
    for (int off = seg >> 1; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off, seg);
    if (valid && i0 == 0) cell[pair] = s;
}

// out[0] = sum (tv - tv_prev)^2, out[1] = sum tv^2, out[2] = sum tc^2   (one CTA, fixed order)
__global__ void __launch_bounds__(1024) norms_kernel(int ne, size_t ntc, const double *__restrict__ tv, const double *__restrict__ tv_prev,
                                                    const double *__restrict__ tc, double *__restrict__ out)
{
    __shared__ double sh[3][32];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int e = threadIdx.x; e < ne; e += blockDim.x)
    {
        const double d = tv[e] - tv_prev[e];
        a += d * d;
        b += tv[e] * tv[e];
    }
    for (size_t i = threadIdx.x; i < ntc; i += blockDim.x) c += tc[i] * tc[i];
    for (int off = 16; off > 0; off >>= 1)
    {
        a += __shfl_down_sync(0xffffffffu, a, off);
        b += __shfl_down_sync(0xffffffffu, b, off);
        c += __shfl_down_sync(0xffffffffu, c, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        sh[0][warp] = a;
        sh[1][warp] = b;
        sh[2][warp] = c;
    }
    __syncthreads();
    if (warp == 0)
    {
        const int nw = blockDim.x >> 5;
        a = lane < nw ? sh[0][lane] : 0.0;
        b = lane < nw ? sh[1][lane] : 0.0;
        c = lane < nw ? sh[2][lane] : 0.0;
        for (int off = 16; off > 0; off >>= 1)
        {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
            c += __shfl_down_sync(0xffffffffu, c, off);
        }
        if (lane == 0)
        {
            out[0] = a;
            out[1] = b;
            out[2] = c;
        }
    }
}

}  // namespace pbte_dev
