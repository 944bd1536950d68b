/*
 * pbte_cuda.h — C ABI of the B200-native PBTE sweep library (libpbte_cuda.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no FFI of its own:
 * its boundary is the C++ class API that src/PhononBTE.cpp:385-398 calls.  The kept host classes
 * (pbte::PBTESolver, pbte::AngularSweepOrder, pbte::MacroscopicQuantities — see
 * dg-solver-for-pbte-with-mfem_b200/host/) flatten their inputs into the POD "problem image" below
 * and call these entry points; nothing but plain pointers, sizes and status codes crosses the boundary.
 *
 * Every array is FP64 / int32, matrices are column-major (entry (i,j) at [i + j*ndof]) exactly like
 * mfem::DenseMatrix, indices are 0-based.  All functions return PBTE_OK (0) or a negative status;
 * pbte_cuda_last_error() gives the message of the last failure on the calling thread.
 * There is no CPU fallback: every compute entry point fails with PBTE_ERR_CUDA when no device is usable.
 *
 * Replaces (reference file:line)                                         entry point
 *   AngularSweepOrder::Build           src/AngularSweepOrder.cpp:78-147  pbte_cuda_build_schedule
 *   AngularSweepOrder::Order           include/AngularSweepOrder.hpp:28  pbte_cuda_get_order
 *   PBTESolver::PBTESolver (BuildCellData, AssembleA, LU policy)
 *                                      src/PBTESolver.cpp:10-206         pbte_cuda_create
 *   PBTESolver::CreateInitialCoefficients  src/PBTESolver.cpp:76-95      pbte_cuda_reset_state
 *   PBTESolver::Solve                  src/PBTESolver.cpp:208-332        pbte_cuda_solve
 *     one pass of its k/b/s/element loop  :238-312                       pbte_cuda_sweep
 *   MacroscopicQuantities::Reset / AccumulateDirectionalCoeff
 *                                      src/MacroscopicQuantities.cpp:93-128   (fused into pbte_cuda_sweep)
 *   MacroscopicQuantities::Finalize / Residual  :130-166                 pbte_cuda_finish_iteration
 *   accessors Tc/Tv/Qc/Qv              include/MacroscopicQuantities.hpp:67-70   pbte_cuda_download_moments
 *   coeff[k][b][s] in/out of Solve     src/PBTESolver.cpp:208,252,308    pbte_cuda_upload_coeff / download_coeff
 */
#ifndef PBTE_CUDA_H
#define PBTE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBTE_ABI_VERSION 1
#define PBTE_MAX_FACES 8 /* local faces per element (tri 3, quad 4, tet 4, hex 6) */

enum pbte_status
{
    PBTE_OK = 0,
    PBTE_ERR_INVALID = -1,      /* bad argument / size mismatch         (MFEM_VERIFY in the reference)        */
    PBTE_ERR_CUDA = -2,         /* CUDA runtime failure or no device                                           */
    PBTE_ERR_SWEEP_STALL = -3,  /* cyclic upwind dependency             (std::runtime_error, AngularSweepOrder.cpp:138-142) */
    PBTE_ERR_MISSING_BC = -4,   /* boundary face without a temperature  (std::out_of_range, PBTESolver.cpp:287)   */
    PBTE_ERR_ZERO_NORM = -5,    /* ||Tv|| == 0 in the residual          (MFEM_VERIFY, MacroscopicQuantities.cpp:162) */
    PBTE_ERR_SINGULAR = -6,     /* zero pivot in a local system                                                */
    PBTE_ERR_NOMEM = -7,
    PBTE_ERR_STATE = -8         /* call order violated (e.g. sweep before build_schedule)                      */
};

/* face_nbr special values */
#define PBTE_FACE_BOUNDARY (-1) /* isothermal wall; face_tbc holds its temperature */
#define PBTE_FACE_NONE (-2)     /* padding slot (element has fewer than nfaces faces) */

/*
 * Problem image: everything PBTESolver's constructor receives, flattened.
 * Element-local faces are in the order of ElementIntegralData::face_couplings (ascending global face id,
 * src/ElementIntegrator.cpp:118-257).  Per-element strides: ndof*ndof for matrices, ndof for vectors.
 */
typedef struct pbte_problem
{
    int32_t abi_version; /* PBTE_ABI_VERSION */
    int32_t dim;         /* mesh dimension = number of stiffness matrices and of direction components used */
    int32_t ne;          /* elements */
    int32_t ndof;        /* local degrees of freedom n */
    int32_t nfaces;      /* face slots per element F (<= PBTE_MAX_FACES) */
    int32_t ndir;        /* directions N_a */
    int32_t nband;       /* bands N_b = nbranch * nspec, band = branch*nspec + spec */
    int32_t reserved0;

    const double *basis_integrals; /* [ne][ndof]            ElementIntegralData::basis_integrals     */
    const double *mass;            /* [ne][ndof*ndof]       ::mass_matrix                            */
    const double *stiff;           /* [ne][dim][ndof*ndof]  ::stiffness_matrices                     */
    const double *face_mass;       /* [ne][F][ndof*ndof]    ::face_mass_matrices                     */
    const double *face_coupling;   /* [ne][F][ndof*ndof]    FaceCoupling::coupling (interior faces)  */
    const double *face_integral;   /* [ne][F][ndof]         FaceCoupling::isothermal_rhs (boundary)  */
    const double *face_normal;     /* [ne][F][3]            utils::ComputeOutwardFaceNormal          */
    const int32_t *face_nbr;       /* [ne][F]  neighbour element, PBTE_FACE_BOUNDARY or PBTE_FACE_NONE */
    const double *face_tbc;        /* [ne][F]  wall temperature iso_bc_.at(attr) for boundary faces  */

    const double *dir;     /* [ndir][3]  AngleDirection::direction */
    const double *dir_w;   /* [ndir]     AngleDirection::weight    */
    double omega;          /* AngleQuadrature::TotalWeight()       */

    const double *inv_kn;   /* [nband] PhononProperties::InvKn          (nu)    */
    const double *vg;       /* [nband] ::GroupVelocity                  (v)     */
    const double *heat_cap; /* [nband] ::HeatCapacity                   (C)     */
    const double *freq_w;   /* [nband] ::FrequencyWeight                (Delta) */
    double heat_cap_v;      /* ::avgHeatCapacity()                      (C_V)   */
} pbte_problem_t;

typedef struct pbte_options
{
    int32_t device;     /* CUDA device ordinal; -1 = current device                                  */
    int32_t dir_begin;  /* this handle sweeps directions [dir_begin, dir_end); 0,0 = all             */
    int32_t dir_end;    /*   (the (angle,band) partition of SURVEY §8e: one handle per GPU)          */
    int32_t cache_policy; /* 0 = FullLU (cache pivot sequences, CachePolicy::FullLU), 1 = OnTheFly    */
    int32_t kernel;     /* 0 = auto, 1 = generic shared-memory kernel, 2 = warp-register kernel (both: one pivoted LU per
                           (element, direction, band), any mesh), 3 = shared-operator marching kernel (elements grouped
                           into operator classes, FP64 tensor-core MMAs; PBTE_ERR_INVALID when the mesh has too many classes) */
    int32_t use_graph;  /* 1 = replay the per-iteration launch sequence as a CUDA graph (default), 0 = plain launches */
    void *stream;       /* cudaStream_t to run on; NULL = a stream owned by the handle               */
    double class_tol;   /* relative max-norm tolerance under which two local matrices count as equal when elements are
                           grouped into operator classes; <= 0 = default 1e-12                      */
    int32_t max_classes; /* give up on the shared-operator path beyond this many classes; <= 0 = default 512 */
    int32_t reserved[5];
} pbte_options_t;

typedef struct pbte_handle pbte_handle_t;

/* ---- life cycle -------------------------------------------------------------------------------- */
int pbte_cuda_abi_version(void);
const char *pbte_cuda_last_error(void);
int pbte_cuda_device_count(int *count);
void pbte_cuda_default_options(pbte_options_t *opt);
/* Copies the problem image to the device (the caller may free its arrays afterwards).  Elements are grouped into operator
 * classes here (host side); when the shared-operator marching path is certain to run (kernel = 3, or kernel = 0 on an image
 * above PBTE_IMAGE_LIMIT_MB = 1024 MB with few classes) only the class representatives' local matrices go to the device
 * instead of the per-element image (pbte_stats_t::image_bytes, image_reps_only; PBTE_IMAGE=full|reps forces). */
int pbte_cuda_create(const pbte_problem_t *problem, const pbte_options_t *opt, pbte_handle_t **out);
int pbte_cuda_destroy(pbte_handle_t *h);

/* ---- AngularSweepOrder ------------------------------------------------------------------------- */
/* Upwind DAG of every owned direction -> pass numbers (reference list order) and wavefront levels,
 * computed on the device; PBTE_ERR_SWEEP_STALL on a cyclic dependency. */
int pbte_cuda_build_schedule(pbte_handle_t *h);
/* order[ne] = AngularSweepOrder::Order(k), bit-exact.  k is a global direction index owned by h. */
int pbte_cuda_get_order(pbte_handle_t *h, int32_t k, int32_t *order);
/* level[ne] = wavefront level of every element for direction k (the GPU execution schedule). */
int pbte_cuda_get_levels(pbte_handle_t *h, int32_t k, int32_t *level);
/* number of distinct upwind DAGs among the owned directions and the longest level count */
int pbte_cuda_schedule_info(pbte_handle_t *h, int32_t *n_classes, int32_t *max_levels);

/* ---- state (warm start in, results out) --------------------------------------------------------- */
/* coeff = 0, Tc = Tv = 0: PBTESolver::CreateInitialCoefficients + MacroscopicQuantities::Reset. */
int pbte_cuda_reset_state(pbte_handle_t *h);
/* One block coeff[k][band] as an (ndof x ne) column-major matrix, host memory. */
int pbte_cuda_upload_coeff(pbte_handle_t *h, int32_t k, int32_t band, const double *host);
int pbte_cuda_download_coeff(pbte_handle_t *h, int32_t k, int32_t band, double *host);
/* All owned blocks in one call: host = [dir_end - dir_begin][nband][ndof*ne], direction-major, contiguous (the layout of
 * PBTESolver::CreateInitialCoefficients here).  Contiguous transfers through a device staging buffer, padding and slot
 * permutation on the device, copies overlapped with the repacking; full PCIe rate when `host` is pinned memory
 * (pbte_cuda_host_alloc: cudaHostAlloc behind a plain-C call, so that the C++ host layer needs no CUDA headers). */
int pbte_cuda_upload_coeff_all(pbte_handle_t *h, const double *host);
int pbte_cuda_download_coeff_all(pbte_handle_t *h, double *host);
int pbte_cuda_host_alloc(void **ptr, size_t bytes);
int pbte_cuda_host_free(void *ptr);
/* Previous-iteration moments used as the source of the next sweep (macro.Tc()/macro.Tv() on entry). */
int pbte_cuda_upload_moments(pbte_handle_t *h, const double *Tc /*[ndof*ne]*/, const double *Tv /*[ne]*/);
/* Tc [ndof*ne], Qc [dim][ndof*ne], Tv [ne], Qv [dim][ne]; any pointer may be NULL. */
int pbte_cuda_download_moments(pbte_handle_t *h, double *Tc, double *Qc, double *Tv, double *Qv);

/* ---- iteration --------------------------------------------------------------------------------- */
/* One pass of the (k, band, element) loop over the owned directions.  Leaves this handle's PARTIAL
 * moment vector [Tc | Qc_0 .. Qc_{dim-1}] ((1+dim)*ndof*ne doubles, element-major [e][i]) in device memory. */
int pbte_cuda_sweep(pbte_handle_t *h);
/* Device pointer / length of the partial moment vector: multi-GPU callers all-reduce (sum) it in place
 * between pbte_cuda_sweep and pbte_cuda_finish_iteration (one collective per iteration). */
int pbte_cuda_moment_buffer(pbte_handle_t *h, double **device_ptr, size_t *count);
/* Finalize (Tv, Qv), residual ||Tv - Tv_prev|| / ||Tv||, then Tv_prev <- Tv, Tc_prev <- Tc.  Waits for the result. */
int pbte_cuda_finish_iteration(pbte_handle_t *h, double *residual);
/* PBTESolver::Solve on one device: up to max_iter iterations, stops when residual < tol.
 * residual_history (may be NULL) receives one value per executed iteration. */
int pbte_cuda_solve(pbte_handle_t *h, int32_t max_iter, double tol, int32_t *iters_done, double *residual,
                    double *residual_history);

/* One whole source iteration ENQUEUED on the handle's stream and not waited for: sweep, the all-reduce of [Tc | Qc] when a
 * communicator is attached, Finalize + Residual + hand-over (the body of the loop at src/PBTESolver.cpp:235-330).  On a
 * single GPU this is two kernel launches (the marching sweep and a tail that streams the flux once and finishes the
 * iteration), replayed as one CUDA graph.  pbte_cuda_iterate_sync waits for everything enqueued and returns the number of
 * iterations finished since the last solve / reset of the counters and the residual of the last one. */
int pbte_cuda_iterate_async(pbte_handle_t *h);
int pbte_cuda_iterate_sync(pbte_handle_t *h, int64_t *iterations, double *last_residual);
/* pbte_cuda_solve with a per-iteration callback (the line the reference prints, src/PBTESolver.cpp:318-326) and,
 * for n_handles > 1, over several handles on several devices of ONE process (pbte_cuda_comm_attach_all first).
 * The loop never synchronises between iterations: up to PBTE_LOOKAHEAD (default 4) iterations are in flight, the stop
 * test res < tol is also evaluated on the device, and iterations enqueued beyond the stop are no-ops there, so the
 * state left behind is that of the first iteration with res < tol, as in the reference. */
typedef void (*pbte_iteration_callback_t)(void *user, int32_t iteration /*1-based*/, double residual, double tv_norm, double tc_fnorm);
int pbte_cuda_solve_ex(pbte_handle_t *const *handles, int32_t n_handles, int32_t max_iter, double tol, int32_t *iters_done,
                       double *residual, double *residual_history, pbte_iteration_callback_t on_iteration, void *user);

/* ---- multi-GPU: the solver owns its communication ------------------------------------ */
/* The (direction, band) space is partitioned by direction blocks (pbte_options_t::dir_begin/dir_end); the one exchange
 * per iteration is an all-reduce (sum, FP64) of the per-DOF moment vector [Tc | Qc], issued by the library through NCCL
 * on the handle's stream once a communicator is attached.  Replaces the gather + broadcast of Tc / residual of the
 * reference's predecessor ("Reference Project/src/DGSolver/PBTE_NonGraySMRT_MPI.cpp":268-315) and the dead
 * ExchangeFaceNbrData skeleton (src/PBTESolver.cpp:622-632).
 *   one process per GPU:  rank 0 calls pbte_cuda_comm_get_unique_id and ships the 128 bytes to the other ranks (the
 *                         launcher's own channel: MPI_Bcast, a file, torch.distributed ...); every rank then calls
 *                         pbte_cuda_comm_attach(h, id, rank, world_size)   [collective]
 *   one process, N GPUs:  pbte_cuda_comm_attach_all(handles, n) on n handles created on n different devices, then
 *                         pbte_cuda_solve_ex(handles, n, ...)
 * With a communicator attached pbte_cuda_finish_iteration / pbte_cuda_iterate_async / pbte_cuda_solve include the
 * all-reduce, and pbte_cuda_moment_buffer is refused. */
#define PBTE_COMM_ID_BYTES 128
int pbte_cuda_comm_get_unique_id(void *id /*[PBTE_COMM_ID_BYTES]*/);
int pbte_cuda_comm_attach(pbte_handle_t *h, const void *id, int32_t rank, int32_t world_size);
int pbte_cuda_comm_attach_all(pbte_handle_t *const *handles, int32_t n_handles);
int pbte_cuda_comm_detach(pbte_handle_t *h);

/* ||Tv||_2 and ||Tc||_F of the last finish_iteration (the two norms the reference prints per iteration,
 * src/PBTESolver.cpp:318-326). */
int pbte_cuda_last_norms(pbte_handle_t *h, double *tv_norm, double *tc_fnorm);

/* ---- stand-alone MacroscopicQuantities calls (outside Solve) ------------------------------------ */
/* These need only sizes, directions and band tables in the problem image (matrix pointers may be NULL). */
int pbte_cuda_macro_reset(pbte_handle_t *h);                                                   /* ::Reset                      :93-102  */
int pbte_cuda_macro_accumulate(pbte_handle_t *h, int32_t k, int32_t band, const double *coeff); /* ::AccumulateDirectionalCoeff :104-128; coeff = host (ndof x ne) */
int pbte_cuda_macro_finalize(pbte_handle_t *h, const double *basis_integrals /*[ne][ndof] host or NULL*/); /* ::Finalize :130-157 */
int pbte_cuda_macro_residual(pbte_handle_t *h, const double *prev_tv /*[ne] host*/, double *residual);     /* ::Residual :159-166 */

/* ---- introspection for benchmarks --------------------------------------------------------------- */
typedef struct pbte_stats
{
    int64_t kernel_launches;  /* launches of this library's kernels since create                      */
    int64_t sweeps;           /* completed pbte_cuda_sweep calls                                       */
    double last_sweep_ms;     /* device time of the last sweep (CUDA events on the handle's stream)    */
    double last_finish_ms;    /* device time of the last finish_iteration                              */
    int64_t flux_bytes;       /* bytes of persistent angular flux held by this handle                  */
    int64_t launches_per_sweep;
    int32_t kernel_used;      /* 1 generic, 2 warp-register, 3 shared-operator marching                */
    int32_t n_classes;        /* distinct upwind DAGs among the owned directions                       */
    int32_t max_levels;
    int32_t n_operator_classes; /* element classes with identical local operators (0 = not classified)  */
    int32_t march_tiles;      /* 8-element tiles of the marching schedule (all DAG classes)            */
    int32_t march_variant;    /* marching path: 0 operators in shared memory, 1 operators in registers */
    double march_fill;        /* elements / (8 * tiles): occupancy of the MMA row blocks               */
    int64_t operator_bytes;   /* device bytes of the precomputed shared operators                      */
    double last_march_ms;     /* device time of the marching kernel alone inside the last sweep (0 = n/a) */
    int32_t march_persistent; /* marching path: 1 = persistent CTAs with work lists (slots cut into equal shares), 0 = one CTA per slot */
    int32_t march_ctas;       /* CTAs of the marching kernel */
    int64_t image_bytes;      /* device bytes of local matrices held: the per-element image and/or the class representatives */
    double create_ms;         /* host time of pbte_cuda_create (classification + uploads) */
    int32_t image_reps_only;  /* 1 = only the operator-class representatives' matrices were uploaded (marching path only) */
    int32_t march_mask_mismatch; /* (upwind DAG, element) pairs whose n.Omega sign differs from their operator class's for a direction
                                    tangential to a face within class_tol: their inflow term of relative size <= class_tol is dropped */
} pbte_stats_t;
int pbte_cuda_get_stats(pbte_handle_t *h, pbte_stats_t *stats);
/* The cudaStream_t the handle launches on (for CUDA-event timing by the caller). */
int pbte_cuda_get_stream(pbte_handle_t *h, void **stream);
/* Micro-benchmarks for the roofline denominators MEASURED_PEAKS.json lacks: FP64 FMA pipe, FP64 tensor pipe
 * (DMMA m8n8k4) in TFLOP/s and a plain device copy in GB/s (read+write).  Any pointer may be NULL. */
int pbte_cuda_measure_peaks(int device, double *fp64_fma_tflops, double *fp64_mma_tflops, double *hbm_copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* PBTE_CUDA_H */
