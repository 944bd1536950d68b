// AngularSweepOrder on the device: upwind masks of every (direction, element), then the two longest-path
// labelings of the upwind DAG — `pass` (reproduces the reference's repeated ascending scan,
// src/AngularSweepOrder.cpp:103-143: order = elements sorted by (pass, id)) and `level` (wavefront
// schedule) — by monotone parallel relaxation.
#pragma once

#include "common.cuh"

namespace pbte_dev
{
// One thread per (direction, element).  Bit f of in_mask: face f is interior and n_f . Omega < 0 (the
// neighbour must be swept first); bit f of out_mask: n_f . Omega > 0 (outflow face, any kind);
// bit f of bin_mask: boundary face with n_f . Omega < 0.
__global__ void upwind_mask_kernel(int ne, int F, int dim, int ndir, const double *__restrict__ dirs /*[ndir][3]*/,
                                   const double *__restrict__ fnormal, const int32_t *__restrict__ fnbr,
                                   uint8_t *__restrict__ in_mask, uint8_t *__restrict__ out_mask, uint8_t *__restrict__ bin_mask)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)ndir * ne) return;
    const int k = (int)(idx / ne), e = (int)(idx % ne);
    const double om[3] = {dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]};
    unsigned in = 0, out = 0, bin = 0;
    for (int f = 0; f < F; ++f)
    {
        const int nb = fnbr[(size_t)e * F + f];
        if (nb == PBTE_FACE_NONE) continue;
        const double a = dot_seq(fnormal + ((size_t)e * F + f) * 3, om, dim);
        if (a < 0.0)
        {
            if (nb >= 0) in |= 1u << f;
            else bin |= 1u << f;
        }
        else if (a > 0.0) out |= 1u << f;
    }
    in_mask[idx] = (uint8_t)in;
    out_mask[idx] = (uint8_t)out;
    bin_mask[idx] = (uint8_t)bin;
}

// One relaxation sweep over every (class, element):
//   pass(e)  <- max(pass(e),  max_{m upwind of e} pass(m) + [m > e])
//   level(e) <- max(level(e), max_{m upwind of e} level(m) + 1)
// Values only grow and are bounded by the longest path of the DAG, so chaotic in-place updates converge to
// the same fixed point as the sequential scan.  *changed counts the updates of this sweep.
__global__ void relax_kernel(int ne, int F, int nclass, const uint8_t *__restrict__ cls_mask /*[nclass][ne]*/,
                             const int32_t *__restrict__ fnbr, int32_t *pass, int32_t *level, int32_t *changed)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)nclass * ne) return;
    const int c = (int)(idx / ne), e = (int)(idx % ne);
    const unsigned m = cls_mask[idx];
    if (!m) return;
    volatile int32_t *vp = pass + (size_t)c * ne;
    volatile int32_t *vl = level + (size_t)c * ne;
    int p = vp[e], l = vl[e];
    const int p0 = p, l0 = l;
    for (int f = 0; f < F; ++f)
    {
        if (!(m >> f & 1u)) continue;
        const int nb = fnbr[(size_t)e * F + f];
        p = max(p, vp[nb] + (nb > e ? 1 : 0));
        l = max(l, vl[nb] + 1);
    }
    if (p != p0 || l != l0)
    {
        vp[e] = p;
        vl[e] = l;
        atomicAdd(changed, 1);
    }
}

}  // namespace pbte_dev
