#include "PBTESolver.hpp"

#include <cstdlib>

#include "CudaCheck.hpp"
#include "ProblemImage.hpp"

#include <cmath>
#include <iostream>

namespace pbte
{
PBTESolver::PBTESolver(const mfem::Mesh &mesh, const mfem::FiniteElementSpace &fes, const AngleQuadrature &angle_quad,
                       const AngularSweepOrder &sweep_order, const std::vector<ElementIntegralData> &elem_data,
                       const PhononProperties &props, const std::map<int, double> &isothermal_bcs, CachePolicy policy, double tol,
                       int max_iter)
    : quad_(angle_quad), sweep_(sweep_order), elem_data_(elem_data), iso_bc_(isothermal_bcs), tol_(tol), max_iter_(max_iter)
{
    dim_ = mesh.Dimension();
    ne_ = mesh.GetNE();
    ndof_ = fes.GetFE(0)->GetDof();
    ndir_ = static_cast<int>(quad_.Directions().size());
    nbranch_ = static_cast<int>(props.Frequency().size());
    nspec_ = nbranch_ > 0 ? static_cast<int>(props.Frequency(0).size()) : 0;

    // One flat image of everything the reference's constructor copies into CellData / the LU cache
    // (src/PBTESolver.cpp:10-139); uploaded once, then owned by the device library.
    const ProblemImage image = ProblemImage::Full(mesh, quad_, elem_data_, props, iso_bc_);
    missing_attr_ = image.MissingBoundaryAttr();
    pbte_options_t opt;
    pbte_cuda_default_options(&opt);
    opt.cache_policy = policy == CachePolicy::FullLU ? 0 : 1;
    // GPU knobs stay out of config.yaml (the reference's parsers must keep working): environment variables
    if (const char *k = std::getenv("PBTE_KERNEL")) opt.kernel = std::atoi(k);          // 0 auto, 1 generic, 2 warp, 3 shared-operator
    if (const char *t = std::getenv("PBTE_CLASS_TOL")) opt.class_tol = std::atof(t);
    ThrowOnStatus(pbte_cuda_create(&image.View(), &opt, &dev_), "pbte_cuda_create");
    ThrowOnStatus(pbte_cuda_build_schedule(dev_), "pbte_cuda_build_schedule");
}

PBTESolver::~PBTESolver()
{
    if (dev_) pbte_cuda_destroy(dev_);
}

PBTESolver::CoefficientBlocks PBTESolver::CreateInitialCoefficients() const
{
    CoefficientBlocks coeff(ndir_, std::vector<std::vector<mfem::DenseMatrix>>(nbranch_, std::vector<mfem::DenseMatrix>(nspec_)));
    for (auto &per_dir : coeff)
        for (auto &per_branch : per_dir)
            for (mfem::DenseMatrix &m : per_branch)
            {
                m.SetSize(ndof_, ne_);
                m = 0.0;
            }
    return coeff;
}

double PBTESolver::Solve(CoefficientBlocks &coeff, MacroscopicQuantities &macro)
{
    MFEM_VERIFY(static_cast<int>(coeff.size()) == ndir_, "coeff size mismatch (dir)");
    MFEM_VERIFY(static_cast<int>(coeff[0].size()) == nbranch_, "coeff size mismatch (branch)");
    MFEM_VERIFY(static_cast<int>(coeff[0][0].size()) == nspec_, "coeff size mismatch (spec)");

    // same one-time diagnostic as the reference (src/PBTESolver.cpp:218-232)
    if (iso_bc_.empty()) std::cout << "[Serial dbg] iso_bc_ is empty; no isothermal boundaries configured." << std::endl;
    else
    {
        std::cout << "[Serial dbg] iso_bc_ size=" << iso_bc_.size() << " keys:";
        for (const auto &kv : iso_bc_) std::cout << " attr" << kv.first << "->T=" << kv.second;
        std::cout << std::endl;
    }
    // the reference hits iso_bc_.at(attr) on the first boundary face it visits (:287)
    if (missing_attr_ >= 0 && max_iter_ > 0) throw std::out_of_range("map::at (no isothermal temperature for boundary attribute " + std::to_string(missing_attr_) + ")");

    // warm start: caller's flux and moments go to the device
    for (int k = 0; k < ndir_; ++k)
        for (int b = 0; b < nbranch_; ++b)
            for (int s = 0; s < nspec_; ++s)
                ThrowOnStatus(pbte_cuda_upload_coeff(dev_, k, b * nspec_ + s, coeff[k][b][s].Data()), "pbte_cuda_upload_coeff");
    ThrowOnStatus(pbte_cuda_upload_moments(dev_, macro.Tc().Data(), macro.Tv().GetData()), "pbte_cuda_upload_moments");

    double res = 0.0;
    sweep_ms_ = 0.0;
    iters_ = 0;
    for (int iter = 0; iter < max_iter_; ++iter)
    {
        ThrowOnStatus(pbte_cuda_sweep(dev_), "pbte_cuda_sweep");
        const int st = pbte_cuda_finish_iteration(dev_, &res);
        if (st == PBTE_ERR_ZERO_NORM) MFEM_VERIFY(false, "Tv norm is zero; cannot compute residual.");
        ThrowOnStatus(st, "pbte_cuda_finish_iteration");
        pbte_stats_t stats;
        pbte_cuda_get_stats(dev_, &stats);
        sweep_ms_ += stats.last_sweep_ms + stats.last_finish_ms;
        ++iters_;
        if (iter == 0 || res > tol_)
        {
            double tv_norm = 0.0, tc_norm = 0.0;
            pbte_cuda_last_norms(dev_, &tv_norm, &tc_norm);
            std::cout.setf(std::ios::scientific);
            std::cout.precision(6);
            std::cout << "[Serial] iter " << (iter + 1) << ", residual = " << res << ";  Tv_norm=" << tv_norm << " Tc_norm=" << tc_norm
                      << std::endl;
            std::cout.unsetf(std::ios::scientific);
        }
        if (res < tol_) break;
    }

    // results back into the caller-owned objects
    for (int k = 0; k < ndir_; ++k)
        for (int b = 0; b < nbranch_; ++b)
            for (int s = 0; s < nspec_; ++s)
                ThrowOnStatus(pbte_cuda_download_coeff(dev_, k, b * nspec_ + s, coeff[k][b][s].Data()), "pbte_cuda_download_coeff");
    if (iters_ > 0) macro.PullFromDevice(dev_);
    return res;
}
}  // namespace pbte
