#include "PBTESolver.hpp"

#include "CudaCheck.hpp"
#include "ProblemImage.hpp"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace pbte
{
namespace
{
// PBTE_DEVICES: "all", or a comma-separated list of CUDA device ordinals; unset = the current device only
std::vector<int> DeviceList()
{
    std::vector<int> out;
    const char *ev = std::getenv("PBTE_DEVICES");
    if (!ev || !*ev) return {-1};
    const std::string spec(ev);
    if (spec == "all")
    {
        int n = 0;
        pbte_cuda_device_count(&n);
        for (int i = 0; i < n; ++i) out.push_back(i);
    }
    else
    {
        std::stringstream ss(spec);
        std::string item;
        while (std::getline(ss, item, ',')) out.push_back(std::atoi(item.c_str()));
    }
    if (out.empty()) out.push_back(-1);
    return out;
}

struct PrintState
{
    double tol;
};

// the line the reference prints per iteration (src/PBTESolver.cpp:316-326)
void PrintIteration(void *user, int32_t iteration, double res, double tv_norm, double tc_norm)
{
    const PrintState *st = static_cast<const PrintState *>(user);
    if (iteration != 1 && !(res > st->tol)) return;
    std::cout.setf(std::ios::scientific);
    std::cout.precision(6);
    std::cout << "[Serial] iter " << iteration << ", residual = " << res << ";  Tv_norm=" << tv_norm << " Tc_norm=" << tc_norm << std::endl;
    std::cout.unsetf(std::ios::scientific);
}
}  // namespace

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
    // (src/PBTESolver.cpp:10-139); uploaded once per device, then owned by the device library.
    const ProblemImage image = ProblemImage::Full(mesh, quad_, elem_data_, props, iso_bc_);
    missing_attr_ = image.MissingBoundaryAttr();
    pbte_options_t opt;
    pbte_cuda_default_options(&opt);
    opt.cache_policy = policy == CachePolicy::FullLU ? 0 : 1;
    // GPU knobs stay out of config.yaml (the reference's parsers must keep working): environment variables
    if (const char *k = std::getenv("PBTE_KERNEL")) opt.kernel = std::atoi(k);          // 0 auto, 1 generic, 2 warp, 3 shared-operator
    if (const char *t = std::getenv("PBTE_CLASS_TOL")) opt.class_tol = std::atof(t);
    std::vector<int> devices = DeviceList();
    if (static_cast<int>(devices.size()) > std::max(ndir_, 1)) devices.resize(std::max(ndir_, 1));  // at least one direction per device
    const int nd = static_cast<int>(devices.size());
    dir_begin_.assign(nd + 1, 0);
    for (int r = 0; r < nd; ++r) dir_begin_[r + 1] = dir_begin_[r] + ndir_ / nd + (r < ndir_ % nd ? 1 : 0);
    try
    {
        for (int r = 0; r < nd; ++r)
        {
            opt.device = devices[r];
            opt.dir_begin = dir_begin_[r];
            opt.dir_end = dir_begin_[r + 1];
            pbte_handle *h = nullptr;
            ThrowOnStatus(pbte_cuda_create(&image.View(), &opt, &h), "pbte_cuda_create");
            devs_.push_back(h);
            ThrowOnStatus(pbte_cuda_build_schedule(h), "pbte_cuda_build_schedule");
        }
        if (nd > 1) ThrowOnStatus(pbte_cuda_comm_attach_all(devs_.data(), nd), "pbte_cuda_comm_attach_all");
    }
    catch (...)
    {
        for (pbte_handle *h : devs_) pbte_cuda_destroy(h);
        devs_.clear();
        throw;
    }
}

PBTESolver::~PBTESolver()
{
    for (pbte_handle *h : devs_) pbte_cuda_destroy(h);
    for (void *p : slabs_) pbte_cuda_host_free(p);
}

PBTESolver::CoefficientBlocks PBTESolver::CreateInitialCoefficients() const
{
    CoefficientBlocks coeff(ndir_, std::vector<std::vector<mfem::DenseMatrix>>(nbranch_, std::vector<mfem::DenseMatrix>(nspec_)));
    // One pinned slab for all blocks when it is of a size worth pinning (PBTE_PINNED_LIMIT_MB, default 16384): Solve then moves
    // the whole set with two bulk transfers at PCIe rate.  The slab belongs to the solver; the blocks are views into it.
    const size_t block = static_cast<size_t>(ndof_) * ne_, total = block * ndir_ * nbranch_ * nspec_;
    double limit_mb = 16384.0;
    if (const char *ev = std::getenv("PBTE_PINNED_LIMIT_MB")) limit_mb = std::atof(ev);
    double *slab = nullptr;
    if (total > 0 && static_cast<double>(total) * sizeof(double) <= limit_mb * 1048576.0)
    {
        void *p = nullptr;
        if (pbte_cuda_host_alloc(&p, total * sizeof(double)) == PBTE_OK && p)
        {
            slab = static_cast<double *>(p);
            slabs_.push_back(p);
        }
    }
    size_t o = 0;
    for (auto &per_dir : coeff)
        for (auto &per_branch : per_dir)
            for (mfem::DenseMatrix &m : per_branch)
            {
                if (slab) m.UseExternalData(slab + o, ndof_, ne_);
                else m.SetSize(ndof_, ne_);
                m = 0.0;
                o += block;
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

    // do the caller's blocks lie back to back in direction-major order (CreateInitialCoefficients' slab, or any such array)?
    const size_t block = static_cast<size_t>(ndof_) * ne_;
    const int nband = nbranch_ * nspec_;
    bool contiguous = true;
    double *base = coeff[0][0][0].Data();
    for (int k = 0; k < ndir_ && contiguous; ++k)
        for (int b = 0; b < nbranch_ && contiguous; ++b)
            for (int s = 0; s < nspec_; ++s)
            {
                MFEM_VERIFY(coeff[k][b][s].Height() == ndof_ && coeff[k][b][s].Width() == ne_, "coeff block shape mismatch");
                if (coeff[k][b][s].Data() != base + (static_cast<size_t>(k) * nband + b * nspec_ + s) * block) contiguous = false;
            }
    const int nd = static_cast<int>(devs_.size());

    // warm start: caller's flux and moments go to the device(s)
    for (int r = 0; r < nd; ++r)
    {
        if (contiguous) ThrowOnStatus(pbte_cuda_upload_coeff_all(devs_[r], base + static_cast<size_t>(dir_begin_[r]) * nband * block), "pbte_cuda_upload_coeff_all");
        else
            for (int k = dir_begin_[r]; k < dir_begin_[r + 1]; ++k)
                for (int b = 0; b < nbranch_; ++b)
                    for (int s = 0; s < nspec_; ++s)
                        ThrowOnStatus(pbte_cuda_upload_coeff(devs_[r], k, b * nspec_ + s, coeff[k][b][s].Data()), "pbte_cuda_upload_coeff");
        ThrowOnStatus(pbte_cuda_upload_moments(devs_[r], macro.Tc().Data(), macro.Tv().GetData()), "pbte_cuda_upload_moments");
    }

    // the iteration loop of src/PBTESolver.cpp:235-330 runs inside the library: iterations back to back on the device, the
    // stop test evaluated there as well, the reference's per-iteration line printed as the residuals arrive
    double res = 0.0;
    int32_t iters = 0;
    PrintState print{tol_};
    const auto t0 = std::chrono::steady_clock::now();
    const int st = pbte_cuda_solve_ex(devs_.data(), nd, max_iter_, tol_, &iters, &res, nullptr, &PrintIteration, &print);
    sweep_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    iters_ = iters;
    if (st == PBTE_ERR_ZERO_NORM) MFEM_VERIFY(false, "Tv norm is zero; cannot compute residual.");
    ThrowOnStatus(st, "pbte_cuda_solve_ex");

    // results back into the caller-owned objects
    for (int r = 0; r < nd; ++r)
    {
        if (contiguous) ThrowOnStatus(pbte_cuda_download_coeff_all(devs_[r], base + static_cast<size_t>(dir_begin_[r]) * nband * block), "pbte_cuda_download_coeff_all");
        else
            for (int k = dir_begin_[r]; k < dir_begin_[r + 1]; ++k)
                for (int b = 0; b < nbranch_; ++b)
                    for (int s = 0; s < nspec_; ++s)
                        ThrowOnStatus(pbte_cuda_download_coeff(devs_[r], k, b * nspec_ + s, coeff[k][b][s].Data()), "pbte_cuda_download_coeff");
    }
    if (iters_ > 0) macro.PullFromDevice(devs_[0]);
    return res;
}
}  // namespace pbte
