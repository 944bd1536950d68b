/*
 * host_capi.h — plain-C access to the host layer (mesh + FE setup through the kept pbte:: classes) for
 * the Python test-suite and bench.py.  A "case" owns the host objects (SpatialMesh, AngleQuadrature,
 * PhononProperties, element integrals) and the flattened problem image that goes to libpbte_cuda.
 */
#ifndef PBTE_HOST_CAPI_H
#define PBTE_HOST_CAPI_H

#include "pbte_cuda.h"

#define PBTE_HOST_API __attribute__((visibility("default")))

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbte_host_case pbte_host_case_t;

typedef struct pbte_host_spec
{
    const char *mesh;         /* mesh file or builtin name (unit-square[-tri|-quad], unit-cube[-tet|-hex]) */
    int32_t order;            /* DG polynomial order */
    int32_t refine;           /* uniform refinement levels */
    int32_t angle_dim;        /* 2 or 3 */
    int32_t polar_points, azimuth_points;
    int32_t polar_gauss, azimuth_gauss; /* 1 = Gauss-Legendre, 0 = uniform midpoint */
    int32_t face_rule;        /* 0 = reference-compatible face points (quirk Q1), 1 = consistent traces */
    int32_t n_bc;             /* entries in bc_attr / bc_temp */
    const int32_t *bc_attr;
    const double *bc_temp;
    /* material (config/si.yaml keys) */
    int32_t num_spectral;
    double ref_temp, ref_len, lattice_dist, Ai, BL, BT, BU;
    double C_LA[2], C_TA[2];
} pbte_host_spec_t;

PBTE_HOST_API const char *pbte_host_last_error(void);
PBTE_HOST_API int pbte_host_case_create(const pbte_host_spec_t *spec, pbte_host_case_t **out);
PBTE_HOST_API void pbte_host_case_destroy(pbte_host_case_t *c);
/* The flattened problem image (valid until the case is destroyed). */
PBTE_HOST_API const pbte_problem_t *pbte_host_case_problem(const pbte_host_case_t *c);
/* 1-D tables for tests: directions incl. polar/azimuth angles, band tables. */
PBTE_HOST_API int pbte_host_case_band_table(const pbte_host_case_t *c, int which /*0 k,1 w,2 dw,3 vg,4 invkn,5 dos,6 heatcap*/, double *out);
/* Runs the kept C++ API end to end on the GPU: AngularSweepOrder::Build, PBTESolver ctor,
 * CreateInitialCoefficients, Solve(coeff, macro).  Outputs (any may be NULL): Tc [ndof*ne], Qc [dim][ndof*ne],
 * Tv [ne], Qv[dim][ne], orders [ndir][ne], coeff [ndir][nband][ndof*ne].  Returns the residual via *residual;
 * cache_policy 0 FullLU / 1 OnTheFly. */
PBTE_HOST_API int pbte_host_case_solve(pbte_host_case_t *c, int32_t max_iter, double tol, int32_t cache_policy, double *residual, int32_t *iters,
                         double *Tc, double *Qc, double *Tv, double *Qv, int32_t *orders, double *coeff);
/* Persistent solver objects for repeated Solve calls (bench.py end-to-end leg): construction = AngularSweepOrder::Build +
 * PBTESolver constructor; each solve call is PBTESolver::Solve(coeff, macro) on HOST coefficient blocks (uploaded at entry,
 * downloaded at exit, src/PBTESolver.cpp:208-332) continuing from the state the previous call returned. */
typedef struct pbte_host_solver pbte_host_solver_t;
PBTE_HOST_API int pbte_host_solver_create(pbte_host_case_t *c, int32_t cache_policy, pbte_host_solver_t **out);
PBTE_HOST_API int pbte_host_solver_solve(pbte_host_solver_t *s, int32_t max_iter, double tol, double *residual, int32_t *iters, double *device_ms);
PBTE_HOST_API int pbte_host_solver_moments(pbte_host_solver_t *s, double *Tc, double *Tv);
PBTE_HOST_API void pbte_host_solver_destroy(pbte_host_solver_t *s);
/* MacroscopicQuantities stand-alone API on host coefficient blocks: Reset, Accumulate (all k,band), Finalize,
 * Residual(prev_Tv).  coeff layout as above. */
PBTE_HOST_API int pbte_host_case_macro(pbte_host_case_t *c, const double *coeff, const double *prev_tv, double *Tc, double *Qc, double *Tv, double *Qv,
                         double *residual);
/* Evaluate Tc at the vertices of every element (what a ParaView level-of-detail-1 dump holds): out [ne][nverts]. */
PBTE_HOST_API int pbte_host_case_vertex_values(const pbte_host_case_t *c, const double *field /*[ndof*ne]*/, double *out, int32_t *nverts);

#ifdef __cplusplus
}
#endif
#endif
