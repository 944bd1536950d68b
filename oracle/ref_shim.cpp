// ref_shim.cpp — TEST INFRASTRUCTURE.  Thin extern "C" access to the REFERENCE's own classes, compiled from the
// unmodified sources under /root/reference/src against the mfem_lite stand-in (recipe: oracle/Makefile, output
// oracle/_ref/libpbte_ref.so).  Used to (a) pin the C restatement in pbte_oracle.c, (b) generate the golden
// fixtures under tests/golden/, (c) time the reference's CPU path ("cpu_baseline.kind": "reference").
// Nothing here is linked into the product libraries.
#include "AngularQuadrature.hpp"
#include "AngularSweepOrder.hpp"
#include "ElementIntegrator.hpp"
#include "MacroscopicQuantities.hpp"
#include "PBTESolver.hpp"
#include "PhononProperties.hpp"
#include "SpatialMesh.hpp"
#include "Utils.hpp"

#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

namespace
{
thread_local std::string g_err = "no error";
struct Silence
{
    std::streambuf *old;
    std::ostringstream sink;
    Silence() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Silence() { std::cout.rdbuf(old); }
};
}  // namespace

struct ref_case
{
    pbte::SpatialMesh spatial;
    pbte::AngleQuadrature quad;
    pbte::PhononProperties props;
    std::vector<pbte::ElementIntegralData> elem_data;
};

#define REF_TRY try {
#define REF_CATCH                          \
    }                                      \
    catch (const std::out_of_range &ex)    \
    {                                      \
        g_err = ex.what();                 \
        return -4;                         \
    }                                      \
    catch (const std::exception &ex)       \
    {                                      \
        g_err = ex.what();                 \
        return std::strstr(ex.what(), "stalled") ? -3 : -1; \
    }

#define REF_API __attribute__((visibility("default")))
extern "C" {
REF_API const char *ref_last_error(void) { return g_err.c_str(); }

// Same sequence as main(), src/PhononBTE.cpp:84-271 (mesh from config incl. boundary temperatures, scale by
// reference_length, refine, L2 space, angles from config, element integrals, band tables).
REF_API int ref_case_create(const char *config_yaml, const char *si_yaml, int order, int refine, ref_case **out)
{
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    Silence quiet;
    REF_TRY
    auto c = std::make_unique<ref_case>();
    c->spatial.LoadMeshFromConfig(config_yaml);
    const auto material = pbte::PhononProperties::LoadMaterial(si_yaml);
    c->spatial.ScaleCoordinates(material.ref_len);
    c->spatial.UniformRefine(refine);
    c->spatial.BuildDGSpace(order);
    c->quad = pbte::AngleQuadrature::Build(pbte::AngleQuadrature::LoadOptionsFromConfig(config_yaml));
    pbte::DGElementIntegrator integrator(c->spatial.FESpace());
    c->elem_data = integrator.AssembleAll();
    c->props = pbte::PhononProperties::Build(material);
    *out = c.release();
    return 0;
    REF_CATCH
}

REF_API void ref_case_destroy(ref_case *c) { delete c; }

REF_API int ref_case_sizes(const ref_case *c, int *dim, int *ne, int *ndof, int *nfaces, int *ndir, int *nbranch, int *nspec)
{
    const mfem::Mesh &m = c->spatial.Mesh();
    *dim = m.Dimension();
    *ne = m.GetNE();
    *ndof = c->spatial.FESpace().GetFE(0)->GetDof();
    size_t F = 0;
    for (const auto &ed : c->elem_data) F = std::max(F, ed.face_couplings.size());
    *nfaces = static_cast<int>(F);
    *ndir = static_cast<int>(c->quad.Directions().size());
    *nbranch = static_cast<int>(c->props.Frequency().size());
    *nspec = *nbranch > 0 ? static_cast<int>(c->props.Frequency(0).size()) : 0;
    return 0;
}

// Flat problem image in the layout of include/pbte_cuda.h, taken from the reference's own objects
// (ElementIntegralData, utils::ComputeOutwardFaceNormal as called by PBTESolver::BuildCellData :116, iso BC map).
REF_API int ref_case_fill(const ref_case *c, double *basis, double *mass, double *stiff, double *fmass, double *fcoup, double *fint, double *fnormal,
                  int *fnbr, double *ftbc, double *dir, double *dirw, double *omega, double *invkn, double *vg, double *hc, double *fw,
                  double *hcv, double *polar, double *azimuth)
{
    REF_TRY
    int dim, ne, n, F, ndir, nbranch, nspec;
    ref_case_sizes(c, &dim, &ne, &n, &F, &ndir, &nbranch, &nspec);
    const mfem::Mesh &mesh = c->spatial.Mesh();
    const size_t nn = static_cast<size_t>(n) * n;
    const auto &bcs = c->spatial.IsothermalBoundaryTemps();
    for (int e = 0; e < ne; ++e)
    {
        const auto &ed = c->elem_data[e];
        std::memcpy(basis + static_cast<size_t>(e) * n, ed.basis_integrals.GetData(), sizeof(double) * n);
        std::memcpy(mass + e * nn, ed.mass_matrix.Data(), sizeof(double) * nn);
        for (int d = 0; d < dim; ++d) std::memcpy(stiff + (static_cast<size_t>(e) * dim + d) * nn, ed.stiffness_matrices[d].Data(), sizeof(double) * nn);
        for (int f = 0; f < F; ++f)
        {
            const size_t slot = static_cast<size_t>(e) * F + f;
            fnbr[slot] = -2;
            ftbc[slot] = 0.0;
            std::memset(fmass + slot * nn, 0, sizeof(double) * nn);
            std::memset(fcoup + slot * nn, 0, sizeof(double) * nn);
            std::memset(fint + slot * n, 0, sizeof(double) * n);
            std::memset(fnormal + slot * 3, 0, sizeof(double) * 3);
            if (f >= static_cast<int>(ed.face_couplings.size())) continue;
            const auto &fc = ed.face_couplings[f];
            fnbr[slot] = fc.neighbor_elem >= 0 ? fc.neighbor_elem : -1;
            const mfem::Vector nrm = utils::ComputeOutwardFaceNormal(mesh, fc.face_id, e);
            for (int d = 0; d < nrm.Size(); ++d) fnormal[slot * 3 + d] = nrm[d];
            std::memcpy(fmass + slot * nn, ed.face_mass_matrices[f].Data(), sizeof(double) * nn);
            if (fc.neighbor_elem >= 0) std::memcpy(fcoup + slot * nn, fc.coupling.Data(), sizeof(double) * nn);
            else
            {
                std::memcpy(fint + slot * n, fc.isothermal_rhs.GetData(), sizeof(double) * n);
                auto it = bcs.find(fc.boundary_attr);
                ftbc[slot] = it != bcs.end() ? it->second : std::numeric_limits<double>::quiet_NaN();
            }
        }
    }
    const auto &dirs = c->quad.Directions();
    for (int k = 0; k < ndir; ++k)
    {
        for (int d = 0; d < 3; ++d) dir[3 * k + d] = dirs[k].direction[d];
        dirw[k] = dirs[k].weight;
        if (polar) polar[k] = dirs[k].polar;
        if (azimuth) azimuth[k] = dirs[k].azimuth;
    }
    *omega = c->quad.TotalWeight();
    for (int b = 0; b < nbranch; ++b)
        for (int s = 0; s < nspec; ++s)
        {
            invkn[b * nspec + s] = c->props.InvKn(b)[s];
            vg[b * nspec + s] = c->props.GroupVelocity(b)[s];
            hc[b * nspec + s] = c->props.HeatCapacity(b)[s];
            fw[b * nspec + s] = c->props.FrequencyWeight(b)[s];
        }
    *hcv = c->props.avgHeatCapacity();
    return 0;
    REF_CATCH
}

// AngularSweepOrder::Build -> orders [ndir][ne]
REF_API int ref_case_orders(const ref_case *c, int *orders)
{
    REF_TRY
    const pbte::AngularSweepOrder sweep = pbte::AngularSweepOrder::Build(c->spatial.Mesh(), c->quad);
    const int ne = c->spatial.Mesh().GetNE();
    for (int k = 0; k < sweep.NumDirections(); ++k) std::memcpy(orders + static_cast<size_t>(k) * ne, sweep.Order(k).data(), sizeof(int) * ne);
    return 0;
    REF_CATCH
}

// PBTESolver ctor + CreateInitialCoefficients (or caller's warm start) + Solve.  coeff [ndir][nbranch*nspec][ndof*ne] in/out
// (all zeros in = cold start); Tc/Qc/Tv/Qv out.  policy 0 FullLU, 1 OnTheFly.  seconds[0] = constructor (LU cache),
// seconds[1] = Solve.
REF_API int ref_case_solve(ref_case *c, int max_iter, double tol, int policy, double *coeff_io, double *Tc, double *Qc, double *Tv, double *Qv,
                   double *residual, double *seconds)
{
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    Silence quiet;
    REF_TRY
    const mfem::Mesh &mesh = c->spatial.Mesh();
    pbte::MacroscopicQuantities macro(c->spatial.FESpace(), c->props, c->quad);
    macro.Reset();
    const auto t0 = std::chrono::steady_clock::now();
    pbte::AngularSweepOrder sweep = pbte::AngularSweepOrder::Build(mesh, c->quad);
    pbte::PBTESolver solver(mesh, c->spatial.FESpace(), c->quad, sweep, c->elem_data, c->props, c->spatial.IsothermalBoundaryTemps(),
                            policy == 0 ? pbte::CachePolicy::FullLU : pbte::CachePolicy::OnTheFly, tol, max_iter);
    auto coeff = solver.CreateInitialCoefficients();
    const int ne = mesh.GetNE(), n = macro.Tc().Height(), dim = mesh.Dimension();
    const size_t block = static_cast<size_t>(n) * ne;
    if (coeff_io)
    {
        size_t o = 0;
        for (auto &pd : coeff)
            for (auto &pb : pd)
                for (auto &m : pb)
                {
                    std::memcpy(m.Data(), coeff_io + o, sizeof(double) * block);
                    o += block;
                }
    }
    const auto t1 = std::chrono::steady_clock::now();
    const double res = solver.Solve(coeff, macro);
    const auto t2 = std::chrono::steady_clock::now();
    if (seconds)
    {
        seconds[0] = std::chrono::duration<double>(t1 - t0).count();
        seconds[1] = std::chrono::duration<double>(t2 - t1).count();
    }
    if (residual) *residual = res;
    if (coeff_io)
    {
        size_t o = 0;
        for (auto &pd : coeff)
            for (auto &pb : pd)
                for (auto &m : pb)
                {
                    std::memcpy(coeff_io + o, m.Data(), sizeof(double) * block);
                    o += block;
                }
    }
    if (Tc) std::memcpy(Tc, macro.Tc().Data(), sizeof(double) * block);
    if (Tv) std::memcpy(Tv, macro.Tv().GetData(), sizeof(double) * ne);
    for (int d = 0; d < dim; ++d)
    {
        if (Qc) std::memcpy(Qc + d * block, macro.Qc()[d].Data(), sizeof(double) * block);
        if (Qv) std::memcpy(Qv + static_cast<size_t>(d) * ne, macro.Qv()[d].GetData(), sizeof(double) * ne);
    }
    return 0;
    REF_CATCH
}

// Persistent reference solver for timing: the constructor (LU cache of every (dir,band,cell), src/PBTESolver.cpp:50-73) runs once,
// then Solve(coeff, macro) is called repeatedly with max_iter fixed at construction, continuing from the returned state.
struct ref_solver
{
    ref_case *c;
    std::unique_ptr<pbte::MacroscopicQuantities> macro;
    std::unique_ptr<pbte::AngularSweepOrder> sweep;
    std::unique_ptr<pbte::PBTESolver> solver;
    std::vector<std::vector<std::vector<mfem::DenseMatrix>>> coeff;
};

REF_API int ref_solver_create(ref_case *c, int iters_per_call, int policy, ref_solver **out, double *ctor_seconds)
{
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    Silence quiet;
    REF_TRY
    auto s = std::make_unique<ref_solver>();
    s->c = c;
    const auto t0 = std::chrono::steady_clock::now();
    s->macro = std::make_unique<pbte::MacroscopicQuantities>(c->spatial.FESpace(), c->props, c->quad);
    s->macro->Reset();
    s->sweep = std::make_unique<pbte::AngularSweepOrder>(pbte::AngularSweepOrder::Build(c->spatial.Mesh(), c->quad));
    s->solver = std::make_unique<pbte::PBTESolver>(c->spatial.Mesh(), c->spatial.FESpace(), c->quad, *s->sweep, c->elem_data, c->props,
                                                   c->spatial.IsothermalBoundaryTemps(),
                                                   policy == 0 ? pbte::CachePolicy::FullLU : pbte::CachePolicy::OnTheFly, 0.0, iters_per_call);
    s->coeff = s->solver->CreateInitialCoefficients();
    if (ctor_seconds) *ctor_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *out = s.release();
    return 0;
    REF_CATCH
}

REF_API int ref_solver_solve(ref_solver *s, double *residual, double *seconds)
{
    Silence quiet;
    REF_TRY
    const auto t0 = std::chrono::steady_clock::now();
    const double res = s->solver->Solve(s->coeff, *s->macro);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (residual) *residual = res;
    return 0;
    REF_CATCH
}

REF_API void ref_solver_destroy(ref_solver *s) { delete s; }
}  // extern "C"
