/*
 * pbte_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's hot path, operating on the same flat problem image
 * (include/pbte_cuda.h) that the CUDA library consumes.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's library; the product never does.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle_golden.py) against
 *   (1) the reference's own sources compiled unmodified against the mfem_lite stand-in (oracle/_ref), which
 *       reproduce every golden log the reference ships (output/log/*, the decoded VTU, T_slice.txt), and
 *   (2) the committed golden fixtures generated from that build (tests/golden/).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).  Arithmetic
 * keeps the reference's operation order (no FMA contraction: compile without -march=native / -ffast-math).
 */
#include "pbte_cuda.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>

#define NN(p) ((size_t)(p)->ndof * (p)->ndof)

/* mfem::Vector::operator* on (normal, direction): sequential sum over the space dimension. */
static double dot_dim(const double *a, const double *b, int dim)
{
    double s = 0.0;
    for (int d = 0; d < dim; ++d) s += a[d] * b[d];
    return s;
}

/* AngularSweepOrder::Build for one direction — src/AngularSweepOrder.cpp:103-143 (repeated ascending scan;
 * an element is ready when no unprocessed interior neighbour has normal.dir < 0).  Returns PBTE_ERR_SWEEP_STALL
 * when a pass makes no progress (:138-142). */
int oracle_sweep_order(const pbte_problem_t *p, int k, int32_t *order)
{
    const int ne = p->ne, F = p->nfaces;
    const double *d = p->dir + 3 * (size_t)k;
    char *processed = (char *)calloc((size_t)ne > 0 ? ne : 1, 1);
    int count = 0;
    while (count < ne)
    {
        int progressed = 0;
        for (int e = 0; e < ne; ++e)
        {
            if (processed[e]) continue;
            int ready = 1;
            for (int f = 0; f < F; ++f)
            {
                const int nb = p->face_nbr[(size_t)e * F + f];
                if (nb < 0) continue; /* boundary or padding */
                if (processed[nb]) continue;
                if (dot_dim(p->face_normal + ((size_t)e * F + f) * 3, d, p->dim) < 0.0)
                {
                    ready = 0;
                    break;
                }
            }
            if (ready)
            {
                order[count++] = e;
                processed[e] = 1;
                progressed = 1;
            }
        }
        if (!progressed)
        {
            free(processed);
            return PBTE_ERR_SWEEP_STALL;
        }
    }
    free(processed);
    return PBTE_OK;
}

/* PBTESolver::AssembleA — src/PBTESolver.cpp:146-168:  A = dt_inv*M - sum_d vg*dir_d*G_d + sum_{f: n.dir>0} vg*(n.dir)*F_f */
void oracle_assemble_A(const pbte_problem_t *p, double dt_inv, int k, int band, int e, double *A)
{
    const size_t nn = NN(p);
    const int F = p->nfaces, dim = p->dim;
    const double vg = p->vg[band];
    const double *dir = p->dir + 3 * (size_t)k;
    const double *M = p->mass + (size_t)e * nn;
    for (size_t i = 0; i < nn; ++i) A[i] = M[i];
    for (size_t i = 0; i < nn; ++i) A[i] *= dt_inv;
    for (int d = 0; d < dim; ++d)
    {
        const double c = -vg * dir[d];
        const double *G = p->stiff + ((size_t)e * dim + d) * nn;
        for (size_t i = 0; i < nn; ++i) A[i] += c * G[i];
    }
    for (int f = 0; f < F; ++f)
    {
        if (p->face_nbr[(size_t)e * F + f] == PBTE_FACE_NONE) continue;
        const double f_dot = dot_dim(p->face_normal + ((size_t)e * F + f) * 3, dir, dim);
        if (f_dot > 0.0)
        {
            const double c = vg * f_dot;
            const double *Fm = p->face_mass + ((size_t)e * F + f) * nn;
            for (size_t i = 0; i < nn; ++i) A[i] += c * Fm[i];
        }
    }
}

/* mfem::LUFactors::Factor as used by DenseMatrixInverse::Factor (src/PBTESolver.cpp:195-200): in-place,
 * first maximal pivot in the column, rows swapped across the whole matrix, multiply by reciprocal pivot. */
int oracle_lu_factor(int m, double *data, int *ipiv)
{
    for (int i = 0; i < m; ++i)
    {
        int piv = i;
        double a = fabs(data[piv + (size_t)i * m]);
        for (int j = i + 1; j < m; ++j)
        {
            const double b = fabs(data[j + (size_t)i * m]);
            if (b > a)
            {
                a = b;
                piv = j;
            }
        }
        ipiv[i] = piv;
        if (piv != i)
            for (int j = 0; j < m; ++j)
            {
                const double t = data[i + (size_t)j * m];
                data[i + (size_t)j * m] = data[piv + (size_t)j * m];
                data[piv + (size_t)j * m] = t;
            }
        if (data[i + (size_t)i * m] == 0.0) return PBTE_ERR_SINGULAR;
        const double a_ii_inv = 1.0 / data[i + (size_t)i * m];
        for (int j = i + 1; j < m; ++j) data[j + (size_t)i * m] *= a_ii_inv;
        for (int k = i + 1; k < m; ++k)
        {
            const double a_ik = data[i + (size_t)k * m];
            for (int j = i + 1; j < m; ++j) data[j + (size_t)k * m] -= a_ik * data[j + (size_t)i * m];
        }
    }
    return PBTE_OK;
}

/* mfem::LUFactors::Solve (DenseMatrixInverse::Mult, src/PBTESolver.cpp:202-206): permute, unit-lower forward
 * sweep, upper backward sweep, both column oriented. */
void oracle_lu_solve(int m, const double *data, const int *ipiv, double *x)
{
    for (int i = 0; i < m; ++i)
    {
        const double t = x[i];
        x[i] = x[ipiv[i]];
        x[ipiv[i]] = t;
    }
    for (int j = 0; j < m; ++j)
    {
        const double x_j = x[j];
        for (int i = j + 1; i < m; ++i) x[i] -= data[i + (size_t)j * m] * x_j;
    }
    for (int j = m - 1; j >= 0; --j)
    {
        const double x_j = (x[j] /= data[j + (size_t)j * m]);
        for (int i = 0; i < j; ++i) x[i] -= data[i + (size_t)j * m] * x_j;
    }
}

/* y = A x, column by column (mfem::DenseMatrix::Mult) */
static void matvec(int n, const double *A, const double *x, double *y)
{
    for (int r = 0; r < n; ++r) y[r] = x[0] * A[r];
    for (int c = 1; c < n; ++c)
        for (int r = 0; r < n; ++r) y[r] += x[c] * A[r + (size_t)c * n];
}

static void matvec_t(int n, const double *A, const double *x, double *y) /* y = A^T x via an explicit transpose copy */
{
    /* the reference multiplies by the stored transpose mass_t (src/PBTESolver.cpp:104-106,269,275): same column order */
    for (int r = 0; r < n; ++r) y[r] = x[0] * A[(size_t)r * n];
    for (int c = 1; c < n; ++c)
        for (int r = 0; r < n; ++r) y[r] += x[c] * A[c + (size_t)r * n];
}

/* One (k, band) sweep over the mesh in the given element order — src/PBTESolver.cpp:256-308.
 * lu/ipiv: cached factors of this (k,band) ([ne][n*n], [ne][n]) or NULL to factor on the fly. */
static int sweep_block(const pbte_problem_t *p, double dt_inv, int k, int band, const int32_t *order, const double *prev_Tc, double *coeff,
                       const double *lu, const int *ipiv, double *work)
{
    const int n = p->ndof, F = p->nfaces, dim = p->dim, ne = p->ne;
    const size_t nn = NN(p);
    const double invKn = p->inv_kn[band], Cwp = p->heat_cap[band], vg = p->vg[band];
    const double *dir = p->dir + 3 * (size_t)k;
    double *rhs = work, *tmp = rhs + n, *A = tmp + n;
    int *piv_local = (int *)(A + nn);
    for (int idx = 0; idx < ne; ++idx)
    {
        const int e = order[idx];
        const double *M = p->mass + (size_t)e * nn;
        for (int i = 0; i < n; ++i) rhs[i] = 0.0;
        /* invKn*Cwp/omega * M^T Tc_prev */
        matvec_t(n, M, prev_Tc + (size_t)e * n, tmp);
        {
            const double a = invKn * Cwp / p->omega;
            if (a != 0.0)
                for (int i = 0; i < n; ++i) rhs[i] += a * tmp[i];
        }
        /* (dt_inv - invKn) * M^T u_old */
        matvec_t(n, M, coeff + (size_t)e * n, tmp);
        {
            const double a = dt_inv - invKn;
            if (a != 0.0)
                for (int i = 0; i < n; ++i) rhs[i] += a * tmp[i];
        }
        for (int f = 0; f < F; ++f)
        {
            const int nb = p->face_nbr[(size_t)e * F + f];
            if (nb == PBTE_FACE_NONE) continue;
            const double f_dot = dot_dim(p->face_normal + ((size_t)e * F + f) * 3, dir, dim);
            const double coeff_in = 0.5 * vg * (f_dot - fabs(f_dot));
            if (nb == PBTE_FACE_BOUNDARY)
            {
                const double Tbc = p->face_tbc[(size_t)e * F + f];
                if (isnan(Tbc)) return PBTE_ERR_MISSING_BC; /* iso_bc_.at(attr) throws, :287 */
                const double a = -coeff_in * Cwp / p->omega * Tbc;
                const double *g = p->face_integral + ((size_t)e * F + f) * n;
                if (a != 0.0)
                    for (int i = 0; i < n; ++i) rhs[i] += a * g[i];
            }
            else
            {
                matvec(n, p->face_coupling + ((size_t)e * F + f) * nn, coeff + (size_t)nb * n, tmp);
                const double a = -coeff_in;
                if (a != 0.0)
                    for (int i = 0; i < n; ++i) rhs[i] += a * tmp[i];
            }
        }
        if (lu) oracle_lu_solve(n, lu + (size_t)e * nn, ipiv + (size_t)e * n, rhs);
        else
        {
            oracle_assemble_A(p, dt_inv, k, band, e, A);
            const int st = oracle_lu_factor(n, A, piv_local);
            if (st) return st;
            oracle_lu_solve(n, A, piv_local, rhs);
        }
        memcpy(coeff + (size_t)e * n, rhs, sizeof(double) * n);
    }
    return PBTE_OK;
}

/* MacroscopicQuantities::AccumulateDirectionalCoeff — src/MacroscopicQuantities.cpp:104-128 */
void oracle_accumulate(const pbte_problem_t *p, int k, int band, const double *coeff, double *Tc, double *Qc)
{
    const size_t block = (size_t)p->ndof * p->ne;
    const double factor = p->inv_kn[band] * p->dir_w[k] * p->freq_w[band] / p->heat_cap_v;
    for (size_t i = 0; i < block; ++i) Tc[i] += factor * coeff[i];
    for (int d = 0; d < p->dim; ++d)
    {
        const double ff = factor * p->vg[band] * p->dir[3 * (size_t)k + d];
        double *Q = Qc + d * block;
        for (size_t i = 0; i < block; ++i) Q[i] += ff * coeff[i];
    }
}

/* MacroscopicQuantities::Finalize — :130-157 */
void oracle_finalize(const pbte_problem_t *p, const double *Tc, const double *Qc, double *Tv, double *Qv)
{
    const int n = p->ndof, ne = p->ne;
    const size_t block = (size_t)n * ne;
    for (int e = 0; e < ne; ++e)
    {
        const double *m = p->basis_integrals + (size_t)e * n;
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += Tc[(size_t)e * n + i] * m[i];
        Tv[e] = t;
        for (int d = 0; d < p->dim; ++d)
        {
            double q = 0.0;
            for (int i = 0; i < n; ++i) q += Qc[d * block + (size_t)e * n + i] * m[i];
            Qv[(size_t)d * ne + e] = q;
        }
    }
}

/* mfem::Vector::Norml2 (scaled accumulation) */
static double norml2(const double *v, int n)
{
    if (n == 0) return 0.0;
    if (n == 1) return fabs(v[0]);
    double scale = 0.0, sum = 0.0;
    for (int i = 0; i < n; ++i)
    {
        if (v[i] != 0.0)
        {
            const double a = fabs(v[i]);
            if (scale <= a)
            {
                const double r = scale / a;
                sum = 1.0 + sum * (r * r);
                scale = a;
                continue;
            }
            const double r = a / scale;
            sum += r * r;
        }
    }
    return scale * sqrt(sum);
}

/* MacroscopicQuantities::Residual — :159-166; returns NaN-free value or -1 when ||Tv|| == 0 (MFEM_VERIFY abort). */
double oracle_residual(const double *Tv, const double *prev_Tv, int ne)
{
    const double denom = norml2(Tv, ne);
    if (!(denom > 0.0)) return -1.0;
    double *diff = (double *)malloc(sizeof(double) * (size_t)(ne > 0 ? ne : 1));
    for (int e = 0; e < ne; ++e) diff[e] = Tv[e] - prev_Tv[e];
    const double r = norml2(diff, ne) / denom;
    free(diff);
    return r;
}

/* The (k, band) sweeps of one iteration are mutually independent (they read only prev_Tc and their own block,
 * src/PBTESolver.cpp:238-312); worker threads pull (k, band) pairs from a shared counter. */
struct job
{
    const pbte_problem_t *p;
    double dt_inv;
    const int32_t *orders;
    double *coeff;
    const double *prev_Tc;
    double *lu;
    int *ipiv;
    double *work;
    size_t work_sz;
    int phase; /* 0: fill the LU cache, 1: sweep */
    atomic_int next;
    atomic_int status;
};
struct job_arg
{
    struct job *j;
    int tid;
};

static void *job_worker(void *argp)
{
    struct job_arg *a = (struct job_arg *)argp;
    struct job *j = a->j;
    const pbte_problem_t *p = j->p;
    const int n = p->ndof, ne = p->ne, nband = p->nband;
    const size_t nn = NN(p), block = (size_t)n * ne;
    const int total = p->ndir * nband;
    for (;;)
    {
        const int kb = atomic_fetch_add(&j->next, 1);
        if (kb >= total) break;
        const int k = kb / nband, b = kb % nband;
        if (j->phase == 0)
        {
            for (int e = 0; e < ne; ++e)
            {
                double *A = j->lu + ((size_t)kb * ne + e) * nn;
                oracle_assemble_A(p, j->dt_inv, k, b, e, A);
                if (oracle_lu_factor(n, A, j->ipiv + ((size_t)kb * ne + e) * n)) atomic_store(&j->status, PBTE_ERR_SINGULAR);
            }
        }
        else
        {
            const int st = sweep_block(p, j->dt_inv, k, b, j->orders + (size_t)k * ne, j->prev_Tc, j->coeff + (size_t)kb * block,
                                       j->lu ? j->lu + (size_t)kb * ne * nn : NULL, j->ipiv ? j->ipiv + (size_t)kb * ne * n : NULL,
                                       j->work + j->work_sz * a->tid);
            if (st) atomic_store(&j->status, st);
        }
    }
    return NULL;
}

static void run_jobs(struct job *j, int nthreads)
{
    atomic_store(&j->next, 0);
    struct job_arg args[256];
    pthread_t th[256];
    if (nthreads > 256) nthreads = 256;
    for (int t = 0; t < nthreads; ++t)
    {
        args[t].j = j;
        args[t].tid = t;
    }
    if (nthreads == 1)
    {
        job_worker(&args[0]);
        return;
    }
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, job_worker, &args[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/*
 * PBTESolver::Solve — src/PBTESolver.cpp:208-332, with the constructor's FullLU cache (:50-73) when cache_lu != 0.
 *   coeff  in/out [ndir][nband][ndof*ne]      Tc/Tv in/out (warm start + result), Qc/Qv out
 *   k_begin/k_end: directions swept (whole range for the serial reference; a sub-range models one GPU's share;
 *                  the moments are then PARTIAL sums and no residual is formed when partial != 0)
 *   threads > 1: the independent (k,band) sweeps run on worker threads (pthreads); moments are still accumulated serially in
 *                the reference's (k, b, s) order afterwards, so results are identical to threads == 1.
 * Returns the number of iterations executed, or a negative status.
 */
int oracle_solve(const pbte_problem_t *p, int max_iter, double tol, int cache_lu, int threads, double *coeff, double *Tc, double *Qc, double *Tv,
                 double *Qv, double *residual_history, double *last_residual)
{
    const int n = p->ndof, ne = p->ne, ndir = p->ndir, nband = p->nband, dim = p->dim;
    const size_t nn = NN(p), block = (size_t)n * ne;
    double dt_inv = 0.0;
    for (int b = 0; b < nband; ++b)
        if (p->inv_kn[b] > dt_inv) dt_inv = p->inv_kn[b]; /* :39-46 */
    int status = PBTE_OK;

    int32_t *orders = (int32_t *)malloc(sizeof(int32_t) * (size_t)ndir * (ne > 0 ? ne : 1));
    for (int k = 0; k < ndir && status == PBTE_OK; ++k) status = oracle_sweep_order(p, k, orders + (size_t)k * ne);
    if (status)
    {
        free(orders);
        return status;
    }
    double *lu = NULL;
    int *ipiv = NULL;
    const int nthreads = threads > 0 ? threads : 1;
    const size_t work_sz = 2 * (size_t)n + nn + (size_t)n; /* rhs, tmp, A, pivots (ints fit in doubles) */
    double *work = (double *)malloc(sizeof(double) * work_sz * nthreads);
    struct job job;
    memset(&job, 0, sizeof(job));
    job.p = p;
    job.dt_inv = dt_inv;
    job.orders = orders;
    job.coeff = coeff;
    job.work = work;
    job.work_sz = work_sz;
    if (cache_lu)
    {
        lu = (double *)malloc(sizeof(double) * (size_t)ndir * nband * ne * nn);
        ipiv = (int *)malloc(sizeof(int) * (size_t)ndir * nband * ne * n);
        if (!lu || !ipiv)
        {
            free(orders);
            free(lu);
            free(ipiv);
            free(work);
            return PBTE_ERR_NOMEM;
        }
        job.phase = 0;
        job.lu = lu;
        job.ipiv = ipiv;
        run_jobs(&job, nthreads);
        if (job.status) status = job.status;
    }
    double *prev_Tv = (double *)malloc(sizeof(double) * (size_t)(ne > 0 ? ne : 1));
    double *prev_Tc = (double *)malloc(sizeof(double) * (block > 0 ? block : 1));
    memcpy(prev_Tv, Tv, sizeof(double) * ne);
    memcpy(prev_Tc, Tc, sizeof(double) * block);

    double res = 0.0;
    int iters = 0;
    for (int iter = 0; iter < max_iter && status == PBTE_OK; ++iter)
    {
        /* macro.Reset() :237 */
        memset(Tc, 0, sizeof(double) * block);
        memset(Qc, 0, sizeof(double) * block * dim);
        job.phase = 1;
        job.lu = lu;
        job.ipiv = ipiv;
        job.prev_Tc = prev_Tc;
        run_jobs(&job, nthreads);
        if (job.status) status = job.status;
        if (status) break;
        for (int k = 0; k < ndir; ++k) /* :311, in (k, b, s) order */
            for (int b = 0; b < nband; ++b) oracle_accumulate(p, k, b, coeff + ((size_t)k * nband + b) * block, Tc, Qc);
        oracle_finalize(p, Tc, Qc, Tv, Qv);
        res = oracle_residual(Tv, prev_Tv, ne);
        if (residual_history) residual_history[iter] = res;
        ++iters;
        if (res < 0.0)
        {
            status = PBTE_ERR_ZERO_NORM;
            break;
        }
        if (res < tol) break; /* :327 */
        memcpy(prev_Tv, Tv, sizeof(double) * ne);
        memcpy(prev_Tc, Tc, sizeof(double) * block);
    }
    if (last_residual) *last_residual = res;
    free(work);
    free(prev_Tv);
    free(prev_Tc);
    free(lu);
    free(ipiv);
    free(orders);
    return status ? status : iters;
}
