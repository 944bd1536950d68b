// Shared device-side definitions of libpbte_cuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "pbte_cuda.h"

namespace pbte_dev
{
constexpr int kMaxFaces = PBTE_MAX_FACES;
constexpr int kMaxDim = 3;
constexpr int kHistRing = 1024;  // entries of the iteration history ring (pinned host memory written by the device)

// Iteration control in device memory: the finish step of an iteration compares the residual with `tol` and raises `stop`;
// every kernel of a later iteration returns at once when it finds `stop` set (csrc/tail.cuh).
struct DevCtrl
{
    double tol;    // stop when residual < tol (0: never)
    int32_t stop;  // raised by the finish step of the first iteration with residual < tol
    int32_t iter;  // finished iterations since the last reset
};

// Device image of the problem plus the sweep state, passed by value to the kernels.
//
// HBM layout (all FP64 unless noted):
//   mass/massT [ne][ms]            ms = ndof*ndof rounded up to an even count (16-byte rows for bulk copies)
//   stiff      [ne][dim][ms]
//   fmass/fcoup[ne][F][ms]
//   fint       [ne][F][ndof]       fnormal [ne][F][3]   ftbc [ne][F]   fnbr int32 [ne][F]   basis [ne][ndof]
//   flux       [nslots][ne][np]    persistent angular flux; slot = (owned direction in schedule-class order) * nband + band.
//                                  One slot = one reference block coeff[k][b][s] (ndof x ne, column-major), so uploads and
//                                  downloads are plain copies, a marching CTA streams one contiguous block and the moment
//                                  pass is perfectly coalesced
//   tc_prev    [ne][np]            previous-iteration pseudo-temperature (source of the sweep); np = ndof rounded up to even so
//                                  that every row is 16-byte aligned (asynchronous 16-byte copies); the padding stays zero
//   partial    [nclass][1+dim][ne][ndof]   per-class moment partial sums written by the sweep (every entry, every sweep)
//   mom        [1+dim][ne][ndof]   Tc | Qc_0.. : the vector that is all-reduced across GPUs
//   slot_*     per-slot scalars: direction (3), v, c1 = nu*C/omega, c2 = nu_max - nu, cb = C/omega, moment factors
struct DevProblem
{
    int dim, ne, n, np, F, nband, nslots, ms, nclass;  // np = row stride (doubles) of flux and tc_prev: n rounded up to even
    double nu_max;
    const double *mass, *massT, *stiff, *fmass, *fcoup, *fint, *fnormal, *ftbc, *basis;
    const int32_t *fnbr;
    const double *slot_dir;  // [nslots][3]
    const double *slot_v, *slot_c1, *slot_c2, *slot_cb, *slot_mf;  // [nslots]
    const double *slot_qf;   // [nslots][3]
    double *flux;
    const double *tc_prev;
    double *partial;
    const int32_t *cls_slot_begin;  // [nclass+1]
    const uint8_t *cls_out;         // [nclass][ne] faces with n.Omega > 0 for some direction of the class
    const uint8_t *cls_in;          // [nclass][ne] interior faces with n.Omega < 0 for some direction of the class
    uint8_t *pivots;                // [nslots][ne][ndof] cached pivot lanes (FullLU policy) or nullptr
    int32_t *error_flag;            // device int: first PBTE_ERR_* raised by a kernel
    const int32_t *stop;            // DevCtrl::stop: sweep and tail kernels return at once when it is set
};

// n . Omega with the reference's operation order: sequential sum over d < dim, no FMA contraction
// (mfem::Vector::operator*, used at src/AngularSweepOrder.cpp:128 and src/PBTESolver.cpp:161,281).
__device__ __forceinline__ double dot_seq(const double *nrm, const double *omega, int dim)
{
    double s = __dmul_rn(nrm[0], omega[0]);
    if (dim > 1) s = __dadd_rn(s, __dmul_rn(nrm[1], omega[1]));
    if (dim > 2) s = __dadd_rn(s, __dmul_rn(nrm[2], omega[2]));
    return s;
}

}  // namespace pbte_dev
