// TEST INFRASTRUCTURE (never linked into the product): CPU emulation of the shared-operator marching sweep.
//
// Builds the operator classes and the tile schedule with the product's own host-side planner (csrc/shared_plan.hpp),
// constructs the operators P, Q_f, h_f with a plain dense LU and then executes the sweep exactly as sweep_march_kernel
// walks it (levels -> tiles -> tile_desc source codes -> compact operator numbering).  tests/test_plan_cpu.py compares the
// result with the oracle, which pins the planner, the schedule tables and the operator algebra without a GPU.
#include "shared_plan.hpp"

#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace
{
// n . Omega in the reference's operation order (src/AngularSweepOrder.cpp:128); compiled with -ffp-contract=off
double dot_seq(const double *a, const double *b, int dim)
{
    double s = a[0] * b[0];
    if (dim > 1) s = s + a[1] * b[1];
    if (dim > 2) s = s + a[2] * b[2];
    return s;
}

// solve A X = R in place (R holds ncol columns, column-major), partial pivoting; returns false on a zero pivot
bool lu_solve(std::vector<double> A, double *R, int n, int ncol)
{
    for (int k = 0; k < n; ++k)
    {
        int piv = k;
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(A[i + (size_t)k * n]) > std::fabs(A[piv + (size_t)k * n])) piv = i;
        if (A[piv + (size_t)k * n] == 0.0) return false;
        if (piv != k)
        {
            for (int j = 0; j < n; ++j) std::swap(A[k + (size_t)j * n], A[piv + (size_t)j * n]);
            for (int j = 0; j < ncol; ++j) std::swap(R[k + (size_t)j * n], R[piv + (size_t)j * n]);
        }
        for (int i = k + 1; i < n; ++i)
        {
            const double l = A[i + (size_t)k * n] / A[k + (size_t)k * n];
            if (l == 0.0) continue;
            for (int j = k + 1; j < n; ++j) A[i + (size_t)j * n] -= l * A[k + (size_t)j * n];
            for (int j = 0; j < ncol; ++j) R[i + (size_t)j * n] -= l * R[k + (size_t)j * n];
        }
    }
    for (int j = 0; j < ncol; ++j)
        for (int i = n - 1; i >= 0; --i)
        {
            double s = R[i + (size_t)j * n];
            for (int c = i + 1; c < n; ++c) s -= A[i + (size_t)c * n] * R[c + (size_t)j * n];
            R[i + (size_t)j * n] = s / A[i + (size_t)i * n];
        }
    return true;
}
}  // namespace

extern "C" int plan_check_run(const pbte_problem_t *p, double tol, int max_classes, int iters, double *Tc_out, double *Qc_out,
                              double *coeff_out, int32_t *info, double *fill)
{
    using namespace pbte_host;
    const int ne = p->ne, n = p->ndof, F = p->nfaces, dim = p->dim, ndir = p->ndir, nband = p->nband;
    const size_t nn = (size_t)n * n, block = (size_t)ne * n;
    const SharedPlan plan = classify_elements(*p, tol, max_classes);
    if (info) info[0] = plan.ok ? plan.noclass : -1, info[1] = plan.naclass;
    if (!plan.ok) return 1;

    // upwind masks, DAG classes, wavefront levels
    std::vector<uint8_t> in((size_t)ndir * ne), bin((size_t)ndir * ne);
    for (int k = 0; k < ndir; ++k)
        for (int e = 0; e < ne; ++e)
        {
            unsigned mi = 0, mb = 0;
            for (int f = 0; f < F; ++f)
            {
                const int nb = p->face_nbr[(size_t)e * F + f];
                if (nb == PBTE_FACE_NONE) continue;
                const double a = dot_seq(p->face_normal + ((size_t)e * F + f) * 3, p->dir + 3 * (size_t)k, dim);
                if (a < 0.0) (nb >= 0 ? mi : mb) |= 1u << f;
            }
            in[(size_t)k * ne + e] = (uint8_t)mi;
            bin[(size_t)k * ne + e] = (uint8_t)mb;
        }
    std::map<std::string, int> seen;
    std::vector<int> class_of_dir(ndir), rep;
    for (int k = 0; k < ndir; ++k)
    {
        std::string key((const char *)in.data() + (size_t)k * ne, ne);
        key.append((const char *)bin.data() + (size_t)k * ne, ne);
        auto it = seen.find(key);
        if (it == seen.end())
        {
            it = seen.emplace(key, (int)rep.size()).first;
            rep.push_back(k);
        }
        class_of_dir[k] = it->second;
    }
    const int nclass = (int)rep.size();
    std::vector<uint8_t> cin((size_t)nclass * ne), cbin((size_t)nclass * ne);
    std::vector<int32_t> level((size_t)nclass * ne, 0);
    for (int c = 0; c < nclass; ++c)
    {
        std::memcpy(cin.data() + (size_t)c * ne, in.data() + (size_t)rep[c] * ne, ne);
        std::memcpy(cbin.data() + (size_t)c * ne, bin.data() + (size_t)rep[c] * ne, ne);
        int32_t *lv = level.data() + (size_t)c * ne;
        for (int sweep = 0;; ++sweep)
        {
            bool changed = false;
            for (int e = 0; e < ne; ++e)
                for (int f = 0; f < F; ++f)
                    if (cin[(size_t)c * ne + e] >> f & 1u)
                    {
                        const int nb = p->face_nbr[(size_t)e * F + f];
                        if (lv[nb] + 1 > lv[e]) lv[e] = lv[nb] + 1, changed = true;
                    }
            if (!changed) break;
            if (sweep > ne + 1) return 2;  // cyclic
        }
    }
    const MarchSchedule S = build_march_schedule(plan, nclass, level.data(), cin.data(), cbin.data());
    if (info) info[2] = nclass, info[3] = S.ntiles, info[4] = S.max_levels;
    if (info)
    {
        // how close behind its last prerequisite a tile follows (tiles the marching pipeline may have to wait for)
        int close8 = 0, close16 = 0;
        for (int t = 0; t < S.ntiles; ++t)
        {
            const int32_t *hdr = S.tile_desc.data() + ((size_t)t * S.rows + 1 + F) * kTileElems;
            if (hdr[3] >= 0 && hdr[2] - hdr[3] <= 8) ++close8;
            if (hdr[3] >= 0 && hdr[2] - hdr[3] <= 16) ++close16;
        }
        info[5] = close8;
        info[6] = close16;
        info[7] = (int32_t)std::min<int64_t>(S.mask_mismatch, INT32_MAX);
    }
    // descriptor headers: operator class, level, index in class, last prerequisite tile (strictly earlier, covers every neighbour)
    for (int c = 0; c < nclass; ++c)
    {
        const int lb = S.cls_level_begin[c], nlev = S.cls_level_begin[c + 1] - lb - 1;
        const int t_begin = S.level_tile_begin[lb], t_end = S.level_tile_begin[lb + nlev];
        std::vector<int> tile_of(ne, -1);
        for (int t = t_begin; t < t_end; ++t)
            for (int k = 0; k < kTileElems; ++k)
                if (S.tile_desc[(size_t)t * S.rows * kTileElems + k] >= 0) tile_of[S.tile_desc[(size_t)t * S.rows * kTileElems + k]] = t - t_begin;
        for (int t = t_begin; t < t_end; ++t)
        {
            const int32_t *d = S.tile_desc.data() + (size_t)t * S.rows * kTileElems, *hdr = d + (size_t)(1 + F) * kTileElems;
            if (hdr[0] != S.tile_oclass[t] || hdr[2] != t - t_begin || hdr[3] >= t - t_begin) return 4;
            for (int m = 0; m < F; ++m)
                for (int k = 0; k < kTileElems; ++k)
                {
                    const int code = d[(size_t)(1 + m) * kTileElems + k];
                    if (code >= 0 && (tile_of[code] < 0 || tile_of[code] > hdr[3])) return 4;
                }
        }
    }
    if (fill) *fill = S.fill;

    // slot tables
    double nu_max = 0.0;
    for (int b = 0; b < nband; ++b) nu_max = std::max(nu_max, p->inv_kn[b]);
    const int noc = plan.noclass, rows = S.rows;
    std::vector<int> op_off((size_t)nclass * noc), op_cnt(nclass);
    for (int c = 0; c < nclass; ++c)
    {
        int cnt = 0;
        for (int o = 0; o < noc; ++o)
        {
            op_off[(size_t)c * noc + o] = cnt;
            cnt += 1 + __builtin_popcount(S.cls_inmask[(size_t)c * noc + o]);
        }
        op_cnt[c] = cnt;
    }

    std::vector<double> flux((size_t)ndir * nband * block, 0.0), tc_prev(block, 0.0), mom((1 + (size_t)dim) * block, 0.0);
    std::vector<double> ops, oph((size_t)noc * F * n), A(nn), R(nn), x(n), acc(n);
    for (int it = 0; it < iters; ++it)
    {
        std::fill(mom.begin(), mom.end(), 0.0);
        for (int k = 0; k < ndir; ++k)
        {
            const int c = class_of_dir[k];
            const double *om = p->dir + 3 * (size_t)k;
            for (int b = 0; b < nband; ++b)
            {
                const double nu = p->inv_kn[b], v = p->vg[b], C = p->heat_cap[b];
                const double c1 = nu * C / p->omega, c2 = nu_max - nu, cb = C / p->omega;
                // operators of this slot (build_operators_kernel)
                ops.assign((size_t)op_cnt[c] * nn, 0.0);
                std::fill(oph.begin(), oph.end(), 0.0);
                for (int o = 0; o < noc; ++o)
                {
                    const int e = plan.oc_rep[o];
                    const unsigned inm = S.cls_inmask[(size_t)c * noc + o];
                    double af[PBTE_MAX_FACES] = {0};
                    for (int j = 0; j < F; ++j)
                        if (plan.cnbr[(size_t)e * F + j] != PBTE_FACE_NONE)
                            af[j] = dot_seq(p->face_normal + ((size_t)e * F + plan.cface[(size_t)e * F + j]) * 3, om, dim);
                    for (size_t idx = 0; idx < nn; ++idx)
                    {
                        double bsum = 0.0;
                        for (int d = 0; d < dim; ++d) bsum += -om[d] * p->stiff[((size_t)e * dim + d) * nn + idx];
                        for (int j = 0; j < F; ++j)
                            if (af[j] > 0.0) bsum += af[j] * p->face_mass[((size_t)e * F + plan.cface[(size_t)e * F + j]) * nn + idx];
                        A[idx] = nu_max * p->mass[(size_t)e * nn + idx] + v * bsum;
                    }
                    double *dst = ops.data() + (size_t)op_off[(size_t)c * noc + o] * nn;
                    for (int i = 0; i < n; ++i)
                        for (int j = 0; j < n; ++j) dst[i + (size_t)j * n] = p->mass[(size_t)e * nn + j + (size_t)i * n];  // M^T
                    if (!lu_solve(A, dst, n, n)) return 3;
                    int m = 1;
                    for (int j = 0; j < F; ++j)
                    {
                        if (!(inm >> j & 1u)) continue;
                        const double cin_f = 0.5 * v * (af[j] - std::fabs(af[j]));
                        const int ks = plan.oc_ksrc[(size_t)o * F + j], gs = plan.oc_gsrc[(size_t)o * F + j];
                        double *q = dst + (size_t)m * nn;
                        if (ks >= 0)
                        {
                            for (size_t idx = 0; idx < nn; ++idx) q[idx] = -cin_f * p->face_coupling[(size_t)ks * nn + idx];
                            if (!lu_solve(A, q, n, n)) return 3;
                        }
                        if (gs >= 0)
                        {
                            double *hv = oph.data() + ((size_t)o * F + j) * n;
                            for (int i = 0; i < n; ++i) hv[i] = -cin_f * cb * p->face_integral[(size_t)gs * n + i];
                            if (!lu_solve(A, hv, n, 1)) return 3;
                        }
                        ++m;
                    }
                }
                // march (sweep_march_kernel)
                double *u = flux.data() + ((size_t)k * nband + b) * block;
                const int lb = S.cls_level_begin[c], nlev = S.cls_level_begin[c + 1] - lb - 1;
                for (int L = 0; L < nlev; ++L)
                    for (int t = S.level_tile_begin[lb + L]; t < S.level_tile_begin[lb + L + 1]; ++t)
                    {
                        const int32_t *desc = S.tile_desc.data() + (size_t)t * rows * kTileElems;
                        const int o = S.tile_oclass[t];
                        const unsigned inm = S.cls_inmask[(size_t)c * noc + o];
                        const double *op = ops.data() + (size_t)op_off[(size_t)c * noc + o] * nn;
                        for (int qd = 0; qd < kTileElems; ++qd)
                        {
                            const int e = desc[qd];
                            if (e < 0) continue;
                            for (int i = 0; i < n; ++i) x[i] = c1 * tc_prev[(size_t)e * n + i] + c2 * u[(size_t)e * n + i];
                            for (int i = 0; i < n; ++i)
                            {
                                double s = 0.0;
                                for (int j = 0; j < n; ++j) s += op[i + (size_t)j * n] * x[j];
                                acc[i] = s;
                            }
                            int m = 1;
                            for (int j = 0; j < F; ++j)
                            {
                                if (!(inm >> j & 1u)) continue;
                                const int code = desc[m * kTileElems + qd];  // row 1 + (m-1): m counts from 1 here
                                if (code >= 0)
                                {
                                    const double *q = op + (size_t)m * nn, *un = u + (size_t)code * n;
                                    for (int i = 0; i < n; ++i)
                                    {
                                        double s = 0.0;
                                        for (int jj = 0; jj < n; ++jj) s += q[i + (size_t)jj * n] * un[jj];
                                        acc[i] += s;
                                    }
                                }
                                else if (code == kSrcWall)
                                {
                                    const double tb = plan.ctbc[(size_t)e * F + j];
                                    for (int i = 0; i < n; ++i) acc[i] += tb * oph[((size_t)o * F + j) * n + i];
                                }
                                ++m;
                            }
                            for (int i = 0; i < n; ++i) u[(size_t)e * n + i] = acc[i];
                        }
                    }
                // moments (MacroscopicQuantities.cpp:104-128)
                const double f = nu * p->dir_w[k] * p->freq_w[b] / p->heat_cap_v;
                for (size_t i = 0; i < block; ++i)
                {
                    mom[i] += f * u[i];
                    for (int d = 0; d < dim; ++d) mom[(1 + (size_t)d) * block + i] += f * v * om[d] * u[i];
                }
            }
        }
        std::memcpy(tc_prev.data(), mom.data(), block * sizeof(double));
    }
    if (Tc_out) std::memcpy(Tc_out, mom.data(), block * sizeof(double));
    if (Qc_out) std::memcpy(Qc_out, mom.data() + block, dim * block * sizeof(double));
    if (coeff_out) std::memcpy(coeff_out, flux.data(), flux.size() * sizeof(double));
    return 0;
}
