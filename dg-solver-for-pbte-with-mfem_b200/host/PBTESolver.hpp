// Source-iteration solver for the non-gray SMRT phonon BTE.  Same public API as the reference's
// include/PBTESolver.hpp:17-46 (CachePolicy, constructor, Solve, CreateInitialCoefficients); the work
// happens in libpbte_cuda through the C ABI of include/pbte_cuda.h.  There is no CPU path.
#pragma once

#include "AngularQuadrature.hpp"
#include "AngularSweepOrder.hpp"
#include "ElementIntegrator.hpp"
#include "MacroscopicQuantities.hpp"
#include "PhononProperties.hpp"

#include "mfem.hpp"

#include <map>
#include <vector>

struct pbte_handle;

namespace pbte
{
enum class CachePolicy
{
    FullLU,   // reference: factor every (dir,band,cell) once and keep it.  Here: keep the pivot sequences
              // on the device (n bytes per system) and re-factor in registers each sweep.
    OnTheFly  // nothing cached; pivots searched in every sweep
};

class PBTESolver
{
public:
    typedef std::vector<std::vector<std::vector<mfem::DenseMatrix>>> CoefficientBlocks;  // [dir][branch][spec] -> ndof x ne

    PBTESolver(const mfem::Mesh &mesh, const mfem::FiniteElementSpace &fes, const AngleQuadrature &angle_quad,
               const AngularSweepOrder &sweep_order, const std::vector<ElementIntegralData> &elem_data,
               const PhononProperties &props, const std::map<int, double> &isothermal_bcs,
               CachePolicy policy = CachePolicy::FullLU, double tol = 1e-8, int max_iter = 200);
    ~PBTESolver();
    PBTESolver(const PBTESolver &) = delete;
    PBTESolver &operator=(const PBTESolver &) = delete;

    /// In-place source iteration from the caller's coeff / macro (warm start); returns the last residual.
    double Solve(CoefficientBlocks &coeff, MacroscopicQuantities &macro);
    CoefficientBlocks CreateInitialCoefficients() const;

    /// Extension: change tolerance / iteration cap between Solve calls (the reference fixes them at construction).
    void SetIterationControl(double tol, int max_iter)
    {
        tol_ = tol;
        max_iter_ = max_iter;
    }
    /// Time the iteration loop of the last Solve took (ms, host clock around the loop: the iterations run back to back on the
    /// device without host synchronisation) and its iteration count; DOF*angle*band updates per sweep.
    double LastSolveSweepMs() const { return sweep_ms_; }
    int LastSolveIterations() const { return iters_; }
    double UpdatesPerSweep() const { return static_cast<double>(ndof_) * ne_ * ndir_ * nbranch_ * nspec_; }

private:
    const AngleQuadrature &quad_;
    const AngularSweepOrder &sweep_;
    const std::vector<ElementIntegralData> &elem_data_;
    std::map<int, double> iso_bc_;
    double tol_;
    int max_iter_;
    int dim_ = 0, ndof_ = 0, ne_ = 0, nbranch_ = 0, nspec_ = 0, ndir_ = 0;
    int missing_attr_ = -1;
    // One handle per device.  PBTE_DEVICES=0,1,.. (or "all") spreads the directions over several GPUs of this process: every handle
    // owns a contiguous direction block, the library all-reduces [Tc | Qc] once per iteration (NCCL, include/pbte_cuda.h).
    std::vector<pbte_handle *> devs_;
    std::vector<int> dir_begin_;  // [devs_.size() + 1]
    // Coefficient storage handed out by CreateInitialCoefficients: one pinned slab per call, [dir][branch][spec][ne][ndof], the
    // blocks are views into it (mfem::DenseMatrix::UseExternalData).  Solve moves such a set with two bulk transfers.
    mutable std::vector<void *> slabs_;
    double sweep_ms_ = 0.0;
    int iters_ = 0;
};
}  // namespace pbte
