// MacroscopicQuantities on the device: class-partial reduction, cell integrals, residual and norms.
// Sums over the dof index use warp-shuffle segmented reductions; every reduction has a fixed order, so a
// sweep is bit-reproducible from run to run.
#pragma once

#include "common.cuh"

namespace pbte_dev
{
// mom[m][e][i] = sum_c partial[c][m][e][i]      (classes in ascending order)
__global__ void reduce_partials_kernel(size_t count /*(1+dim)*ne*n*/, int nclass, const double *__restrict__ partial, double *__restrict__ mom)
{
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= count) return;
    double s = 0.0;
    for (int c = 0; c < nclass; ++c) s += partial[(size_t)c * count + idx];
    mom[idx] = s;
}

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

// ---- fused end of a source iteration (pbte_cuda_finish_iteration) -------------------------------------------------
// finish_cells_kernel: cell integrals of all moments (as finalize_kernel) and, for the temperature moment, the three
// squared norms of the stop test as per-CTA partial sums, plus the hand-over to the next iteration in the same pass:
// tv_prev <- Tv and tc_prev <- Tc (padded rows).  Every (moment, element) pair is owned by one lane segment, which reads
// tv_prev[e] before it overwrites it.  finish_norms_kernel adds the per-CTA partials in a fixed order (bit-reproducible).
__global__ void __launch_bounds__(256) finish_cells_kernel(int ne, int n, int np, int nmom, int seg, const double *__restrict__ mom,
                                                           const double *__restrict__ basis, double *__restrict__ cell /*[nmom][ne]*/,
                                                           double *__restrict__ tv_prev, double *__restrict__ tc_prev /*[ne][np]*/,
                                                           double *__restrict__ block_part /*[gridDim.x][3]*/)
{
    __shared__ double sh[3][8];
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t pair = gid / seg;
    const int i0 = (int)(gid % seg);
    const bool valid = pair < (size_t)nmom * ne;
    const bool is_t = valid && pair < (size_t)ne;  // moment 0 = Tc
    double s = 0.0, c2 = 0.0;
    if (valid)
    {
        const size_t e = pair % ne;
        for (int i = i0; i < n; i += seg)
        {
            const double v = mom[pair * n + i];
            s += v * basis[e * n + i];
            if (is_t)
            {
                c2 += v * v;
                tc_prev[e * np + i] = v;
            }
        }
    }
    for (int off = seg >> 1; off > 0; off >>= 1)
    {
        s += __shfl_down_sync(0xffffffffu, s, off, seg);
        c2 += __shfl_down_sync(0xffffffffu, c2, off, seg);
    }
    double d2 = 0.0, t2 = 0.0;
    if (valid && i0 == 0)
    {
        cell[pair] = s;
        if (is_t)
        {
            const double d = s - tv_prev[pair];
            d2 = d * d;
            t2 = s * s;
            tv_prev[pair] = s;
        }
    }
    if (!(is_t && i0 == 0)) c2 = 0.0;
    // CTA partial sums in a fixed order: warp shuffle tree, then the 8 warps in ascending order
    for (int off = 16; off > 0; off >>= 1)
    {
        d2 += __shfl_down_sync(0xffffffffu, d2, off);
        t2 += __shfl_down_sync(0xffffffffu, t2, off);
        c2 += __shfl_down_sync(0xffffffffu, c2, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        sh[0][warp] = d2;
        sh[1][warp] = t2;
        sh[2][warp] = c2;
    }
    __syncthreads();
    if (threadIdx.x < 3)
    {
        double a = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[threadIdx.x][w];
        block_part[(size_t)blockIdx.x * 3 + threadIdx.x] = a;
    }
}

// out[k] = sum_b block_part[b][k], k = 0..2: strided partial sums per thread, then a fixed tree (one CTA of 256 threads)
__global__ void __launch_bounds__(256) finish_norms_kernel(int nblocks, const double *__restrict__ block_part, double *__restrict__ out)
{
    __shared__ double sh[3][8];
    double a[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += block_part[(size_t)b * 3 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) sh[k][warp] = a[k];
    __syncthreads();
    if (threadIdx.x < 3)
    {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        out[threadIdx.x] = t;
    }
}

}  // namespace pbte_dev
