#include "host_capi.h"

#include "AngularQuadrature.hpp"
#include "AngularSweepOrder.hpp"
#include "ElementIntegrator.hpp"
#include "MacroscopicQuantities.hpp"
#include "PBTESolver.hpp"
#include "PhononProperties.hpp"
#include "ProblemImage.hpp"
#include "SpatialMesh.hpp"

#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

namespace
{
thread_local std::string g_host_err = "no error";
}

struct pbte_host_case
{
    pbte::SpatialMesh spatial;
    pbte::AngleQuadrature quad;
    pbte::PhononProperties props;
    std::vector<pbte::ElementIntegralData> elem_data;
    pbte::ProblemImage image;
};

struct pbte_host_solver
{
    pbte_host_case *c;
    std::unique_ptr<pbte::MacroscopicQuantities> macro;
    std::unique_ptr<pbte::AngularSweepOrder> sweep;
    std::unique_ptr<pbte::PBTESolver> solver;
    pbte::PBTESolver::CoefficientBlocks coeff;
    int policy = 0;
};

#define HOST_TRY try {
#define HOST_CATCH                                   \
    }                                                \
    catch (const std::out_of_range &ex)              \
    {                                                \
        g_host_err = ex.what();                      \
        return PBTE_ERR_MISSING_BC;                  \
    }                                                \
    catch (const mfem::ErrorException &ex)           \
    {                                                \
        g_host_err = ex.what();                      \
        return PBTE_ERR_INVALID;                     \
    }                                                \
    catch (const std::invalid_argument &ex)          \
    {                                                \
        g_host_err = ex.what();                      \
        return PBTE_ERR_INVALID;                     \
    }                                                \
    catch (const std::exception &ex)                 \
    {                                                \
        g_host_err = ex.what();                      \
        return std::strstr(ex.what(), "stalled") ? PBTE_ERR_SWEEP_STALL : PBTE_ERR_CUDA; \
    }

extern "C" {

const char *pbte_host_last_error(void) { return g_host_err.c_str(); }

int pbte_host_case_create(const pbte_host_spec_t *s, pbte_host_case_t **out)
{
    if (!s || !out) return PBTE_ERR_INVALID;
    *out = nullptr;
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    HOST_TRY
    auto c = std::make_unique<pbte_host_case>();
    c->spatial.LoadMesh(s->mesh ? s->mesh : "");
    for (int i = 0; i < s->n_bc; ++i) c->spatial.SetIsothermalBoundary(s->bc_attr[i], s->bc_temp[i]);
    c->spatial.ScaleCoordinates(s->ref_len);
    c->spatial.UniformRefine(s->refine);
    c->spatial.BuildDGSpace(s->order, mfem::Ordering::byNODES, "-");

    pbte::AngleDiscretizationOptions ao;
    ao.dimension = s->angle_dim;
    ao.polar_points = s->polar_points;
    ao.azimuth_points = s->azimuth_points;
    ao.polar_scheme = s->polar_gauss ? pbte::AngleDiscretizationType::GaussLegendre : pbte::AngleDiscretizationType::Uniform;
    ao.azimuth_scheme = s->azimuth_gauss ? pbte::AngleDiscretizationType::GaussLegendre : pbte::AngleDiscretizationType::Uniform;
    c->quad = pbte::AngleQuadrature::Build(ao);

    pbte::PhononMaterial m;
    m.num_branches = 2;
    m.C_LA = {s->C_LA[0], s->C_LA[1]};
    m.C_TA = {s->C_TA[0], s->C_TA[1]};
    m.lattice_dist = s->lattice_dist;
    m.K_range = {0.0, 2.0 * 3.14159265358979323846 / s->lattice_dist};
    m.Ai = s->Ai;
    m.BL = s->BL;
    m.BT = s->BT;
    m.BU = s->BU;
    m.num_spectral = s->num_spectral;
    m.ref_temp = s->ref_temp;
    m.ref_len = s->ref_len;
    c->props = pbte::PhononProperties::Build(m);

    pbte::DGElementIntegrator integ(c->spatial.FESpace());
    integ.SetFacePointRule(s->face_rule ? pbte::FacePointRule::Consistent : pbte::FacePointRule::ReferenceCompatible);
    c->elem_data = integ.AssembleAll();
    c->image = pbte::ProblemImage::Full(c->spatial.Mesh(), c->quad, c->elem_data, c->props, c->spatial.IsothermalBoundaryTemps());
    *out = c.release();
    return PBTE_OK;
    HOST_CATCH
}

void pbte_host_case_destroy(pbte_host_case_t *c) { delete c; }

const pbte_problem_t *pbte_host_case_problem(const pbte_host_case_t *c) { return c ? &c->image.View() : nullptr; }

int pbte_host_case_band_table(const pbte_host_case_t *c, int which, double *out)
{
    if (!c || !out) return PBTE_ERR_INVALID;
    const pbte::PhononProperties &p = c->props;
    const pbte::PhononProperties::Table *t = nullptr;
    switch (which)
    {
        case 0: t = &p.WaveVector(); break;
        case 1: t = &p.Frequency(); break;
        case 2: t = &p.FrequencyWeight(); break;
        case 3: t = &p.GroupVelocity(); break;
        case 4: t = &p.InvKn(); break;
        case 5: t = &p.DensityOfStates(); break;
        case 6: t = &p.HeatCapacity(); break;
        default: return PBTE_ERR_INVALID;
    }
    size_t o = 0;
    for (const auto &row : *t)
        for (double v : row) out[o++] = v;
    return PBTE_OK;
}

int pbte_host_case_solve(pbte_host_case_t *c, int32_t max_iter, double tol, int32_t cache_policy, double *residual, int32_t *iters,
                         double *Tc, double *Qc, double *Tv, double *Qv, int32_t *orders, double *coeff_out)
{
    if (!c) return PBTE_ERR_INVALID;
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    HOST_TRY
    const mfem::Mesh &mesh = c->spatial.Mesh();
    // the call sequence of the reference driver, src/PhononBTE.cpp:279-398
    pbte::MacroscopicQuantities macro(c->spatial.FESpace(), c->props, c->quad);
    macro.Reset();
    pbte::AngularSweepOrder sweep = pbte::AngularSweepOrder::Build(mesh, c->quad);
    pbte::PBTESolver solver(mesh, c->spatial.FESpace(), c->quad, sweep, c->elem_data, c->props, c->spatial.IsothermalBoundaryTemps(),
                            cache_policy == 0 ? pbte::CachePolicy::FullLU : pbte::CachePolicy::OnTheFly, tol, max_iter);
    auto coeff = solver.CreateInitialCoefficients();
    const double res = solver.Solve(coeff, macro);
    if (residual) *residual = res;
    if (iters) *iters = solver.LastSolveIterations();
    const int ne = mesh.GetNE(), n = macro.Tc().Height(), dim = mesh.Dimension();
    const size_t block = static_cast<size_t>(n) * ne;
    if (Tc) std::memcpy(Tc, macro.Tc().Data(), sizeof(double) * block);
    if (Tv) std::memcpy(Tv, macro.Tv().GetData(), sizeof(double) * ne);
    for (int d = 0; d < dim; ++d)
    {
        if (Qc) std::memcpy(Qc + d * block, macro.Qc()[d].Data(), sizeof(double) * block);
        if (Qv) std::memcpy(Qv + static_cast<size_t>(d) * ne, macro.Qv()[d].GetData(), sizeof(double) * ne);
    }
    if (orders)
        for (int k = 0; k < sweep.NumDirections(); ++k) std::memcpy(orders + static_cast<size_t>(k) * ne, sweep.Order(k).data(), sizeof(int32_t) * ne);
    if (coeff_out)
    {
        size_t o = 0;
        for (auto &pd : coeff)
            for (auto &pb : pd)
                for (auto &m : pb)
                {
                    std::memcpy(coeff_out + o, m.Data(), sizeof(double) * block);
                    o += block;
                }
    }
    return PBTE_OK;
    HOST_CATCH
}

int pbte_host_solver_create(pbte_host_case_t *c, int32_t cache_policy, pbte_host_solver_t **out)
{
    if (!c || !out) return PBTE_ERR_INVALID;
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    HOST_TRY
    auto s = std::make_unique<pbte_host_solver>();
    s->c = c;
    s->policy = cache_policy;
    s->macro = std::make_unique<pbte::MacroscopicQuantities>(c->spatial.FESpace(), c->props, c->quad);
    s->sweep = std::make_unique<pbte::AngularSweepOrder>(pbte::AngularSweepOrder::Build(c->spatial.Mesh(), c->quad));
    s->solver = std::make_unique<pbte::PBTESolver>(c->spatial.Mesh(), c->spatial.FESpace(), c->quad, *s->sweep, c->elem_data, c->props,
                                                   c->spatial.IsothermalBoundaryTemps(),
                                                   cache_policy == 0 ? pbte::CachePolicy::FullLU : pbte::CachePolicy::OnTheFly, 0.0, 1);
    s->coeff = s->solver->CreateInitialCoefficients();
    *out = s.release();
    return PBTE_OK;
    HOST_CATCH
}

int pbte_host_solver_solve(pbte_host_solver_t *s, int32_t max_iter, double tol, double *residual, int32_t *iters, double *device_ms)
{
    if (!s) return PBTE_ERR_INVALID;
    HOST_TRY
    s->solver->SetIterationControl(tol, max_iter);
    std::ostringstream sink;  // the per-iteration log lines of Solve are not part of the measurement
    std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
    double res = 0.0;
    try
    {
        res = s->solver->Solve(s->coeff, *s->macro);
    }
    catch (...)
    {
        std::cout.rdbuf(old);
        throw;
    }
    std::cout.rdbuf(old);
    if (residual) *residual = res;
    if (iters) *iters = s->solver->LastSolveIterations();
    if (device_ms) *device_ms = s->solver->LastSolveSweepMs();
    return PBTE_OK;
    HOST_CATCH
}

int pbte_host_solver_moments(pbte_host_solver_t *s, double *Tc, double *Tv)
{
    if (!s) return PBTE_ERR_INVALID;
    const size_t block = static_cast<size_t>(s->macro->Tc().Height()) * s->macro->Tc().Width();
    if (Tc) std::memcpy(Tc, s->macro->Tc().Data(), sizeof(double) * block);
    if (Tv) std::memcpy(Tv, s->macro->Tv().GetData(), sizeof(double) * s->macro->Tv().Size());
    return PBTE_OK;
}

void pbte_host_solver_destroy(pbte_host_solver_t *s) { delete s; }

int pbte_host_case_macro(pbte_host_case_t *c, const double *coeff, const double *prev_tv, double *Tc, double *Qc, double *Tv, double *Qv,
                         double *residual)
{
    if (!c || !coeff) return PBTE_ERR_INVALID;
    mfem::set_error_action(mfem::MFEM_ERROR_THROW);
    HOST_TRY
    pbte::MacroscopicQuantities macro(c->spatial.FESpace(), c->props, c->quad);
    macro.Reset();
    const mfem::Mesh &mesh = c->spatial.Mesh();
    const int ne = mesh.GetNE(), n = macro.Tc().Height(), dim = mesh.Dimension();
    const int ndir = static_cast<int>(c->quad.Directions().size());
    const int nbranch = static_cast<int>(c->props.Frequency().size()), nspec = static_cast<int>(c->props.Frequency(0).size());
    const size_t block = static_cast<size_t>(n) * ne;
    mfem::DenseMatrix blk(n, ne);
    size_t o = 0;
    for (int k = 0; k < ndir; ++k)
        for (int b = 0; b < nbranch; ++b)
            for (int s = 0; s < nspec; ++s, o += block)
            {
                std::memcpy(blk.Data(), coeff + o, sizeof(double) * block);
                macro.AccumulateDirectionalCoeff(k, b, s, blk);
            }
    macro.Finalize(c->elem_data);
    if (Tc) std::memcpy(Tc, macro.Tc().Data(), sizeof(double) * block);
    if (Tv) std::memcpy(Tv, macro.Tv().GetData(), sizeof(double) * ne);
    for (int d = 0; d < dim; ++d)
    {
        if (Qc) std::memcpy(Qc + d * block, macro.Qc()[d].Data(), sizeof(double) * block);
        if (Qv) std::memcpy(Qv + static_cast<size_t>(d) * ne, macro.Qv()[d].GetData(), sizeof(double) * ne);
    }
    if (prev_tv && residual)
    {
        mfem::Vector prev(ne);
        std::memcpy(prev.GetData(), prev_tv, sizeof(double) * ne);
        *residual = macro.Residual(prev);
    }
    return PBTE_OK;
    HOST_CATCH
}

int pbte_host_case_vertex_values(const pbte_host_case_t *c, const double *field, double *out, int32_t *nverts)
{
    if (!c || !field || !out) return PBTE_ERR_INVALID;
    HOST_TRY
    const mfem::Mesh &mesh = c->spatial.Mesh();
    const mfem::FiniteElementSpace &fes = c->spatial.FESpace();
    static const double ref_tri[3][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    static const double ref_sq[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
    static const double ref_tet[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static const double ref_cube[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    const mfem::Geometry::Type g = mesh.GetElementBaseGeometry(0);
    const double(*ref)[3] = g == mfem::Geometry::TRIANGLE ? ref_tri : g == mfem::Geometry::SQUARE ? ref_sq : g == mfem::Geometry::TETRAHEDRON ? ref_tet : ref_cube;
    const int nv = mfem::Geometry::NumVerts[g];
    if (nverts) *nverts = nv;
    const mfem::FiniteElement *fe = fes.GetFE(0);
    const int n = fe->GetDof();
    mfem::Vector shape(n);
    for (int k = 0; k < nv; ++k)
    {
        mfem::IntegrationPoint ip;
        ip.Set3(ref[k][0], ref[k][1], ref[k][2]);
        fe->CalcShape(ip, shape);
        for (int e = 0; e < mesh.GetNE(); ++e)
        {
            double v = 0.0;
            for (int j = 0; j < n; ++j) v += field[static_cast<size_t>(e) * n + j] * shape[j];
            out[static_cast<size_t>(e) * nv + k] = v;
        }
    }
    return PBTE_OK;
    HOST_CATCH
}
}  // extern "C"
