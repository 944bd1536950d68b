// Angular/spectral moments of the flux: pseudo-temperature Tc, heat flux Qc, cell integrals Tv/Qv and
// the relative residual.  Same public API as the reference's include/MacroscopicQuantities.hpp:20-88.
// All arithmetic runs on the GPU through the C ABI: inside PBTESolver::Solve the moments are fused into
// the sweep kernels; the stand-alone Reset/Accumulate/Finalize/Residual calls below use a small device
// context of their own (created lazily).  The host matrices are mirrors of the device results.
#pragma once

#include "AngularQuadrature.hpp"
#include "ElementIntegrator.hpp"
#include "PhononProperties.hpp"

#include "mfem.hpp"

#include <memory>
#include <string>
#include <vector>

struct pbte_handle;

namespace pbte
{
class PBTESolver;

class MacroscopicQuantities
{
public:
    MacroscopicQuantities(const mfem::FiniteElementSpace &fes, const PhononProperties &props, const AngleQuadrature &quad);
    ~MacroscopicQuantities();
    MacroscopicQuantities(const MacroscopicQuantities &) = delete;
    MacroscopicQuantities &operator=(const MacroscopicQuantities &) = delete;

    void Reset();
    /// Tc += f*coeff, Qc_d += f*v*Omega_d*coeff with f = nu*w_k*Delta/C_V  (coeff: ndof x ne)
    void AccumulateDirectionalCoeff(int dir_idx, int branch, int spec, const mfem::DenseMatrix &coeff);
    void Finalize(const std::vector<ElementIntegralData> &elem_data);
    /// ||Tv - prev_Tv|| / ||Tv||; aborts through MFEM_VERIFY when ||Tv|| = 0.
    double Residual(const mfem::Vector &prev_Tv) const;

    void WriteParaView(const std::string &prefix, bool high_order = true) const;
    void Write2DSliceTemperature(const std::string &out_path, int Nx = 50, int Ny = 50, double clamp_tol = 1e-12) const;

    const mfem::DenseMatrix &Tc() const { return Tc_; }
    const mfem::Vector &Tv() const { return Tv_; }
    const std::vector<mfem::DenseMatrix> &Qc() const { return Qc_; }
    const std::vector<mfem::Vector> &Qv() const { return Qv_; }
    double HeatCapV() const { return heat_cap_v_; }

private:
    friend class PBTESolver;
    void EnsureDevice() const;
    void PullFromDevice(pbte_handle *h);  // refresh the host mirrors from a handle's moment state

    const mfem::FiniteElementSpace &fes_;
    const PhononProperties &props_;
    const AngleQuadrature &quad_;
    int dim_ = 0, ndof_ = 0, ne_ = 0, nspec_ = 0;
    double heat_cap_v_ = 0.0;
    mfem::DenseMatrix Tc_;
    mfem::Vector Tv_;
    std::vector<mfem::DenseMatrix> Qc_;
    std::vector<mfem::Vector> Qv_;
    mutable pbte_handle *dev_ = nullptr;
    bool basis_uploaded_ = false;
};
}  // namespace pbte
