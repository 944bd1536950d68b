#include "MacroscopicQuantities.hpp"

#include "CudaCheck.hpp"
#include "ProblemImage.hpp"

#include <cmath>
#include <filesystem>
#include <fstream>
#include <limits>

namespace pbte
{
MacroscopicQuantities::MacroscopicQuantities(const mfem::FiniteElementSpace &fes, const PhononProperties &props, const AngleQuadrature &quad)
    : fes_(fes), props_(props), quad_(quad)
{
    dim_ = fes_.GetMesh()->Dimension();
    ne_ = fes_.GetNE();
    ndof_ = fes_.GetFE(0)->GetDof();  // uniform-order L2 space, as in the reference (:80)
    nspec_ = props_.Frequency().empty() ? 0 : static_cast<int>(props_.Frequency(0).size());
    heat_cap_v_ = props_.avgHeatCapacity();
    Tc_.SetSize(ndof_, ne_);
    Tv_.SetSize(ne_);
    Qc_.assign(dim_, mfem::DenseMatrix(ndof_, ne_));
    Qv_.assign(dim_, mfem::Vector(ne_));
    // host mirrors start at zero; the device context is created on first use
    Tc_ = 0.0;
    Tv_ = 0.0;
    for (int d = 0; d < dim_; ++d)
    {
        Qc_[d] = 0.0;
        Qv_[d] = 0.0;
    }
}

MacroscopicQuantities::~MacroscopicQuantities()
{
    if (dev_) pbte_cuda_destroy(dev_);
}

void MacroscopicQuantities::EnsureDevice() const
{
    if (dev_) return;
    const ProblemImage image = ProblemImage::MomentsOnly(dim_, ne_, ndof_, quad_, props_);
    pbte_options_t opt;
    pbte_cuda_default_options(&opt);
    ThrowOnStatus(pbte_cuda_create(&image.View(), &opt, &dev_), "pbte_cuda_create (moments)");
}

void MacroscopicQuantities::PullFromDevice(pbte_handle *h)
{
    std::vector<double> qc(static_cast<size_t>(dim_) * ndof_ * ne_), qv(static_cast<size_t>(dim_) * ne_);
    ThrowOnStatus(pbte_cuda_download_moments(h, Tc_.Data(), qc.data(), Tv_.GetData(), qv.data()), "pbte_cuda_download_moments");
    for (int d = 0; d < dim_; ++d)
    {
        std::copy(qc.begin() + static_cast<size_t>(d) * ndof_ * ne_, qc.begin() + static_cast<size_t>(d + 1) * ndof_ * ne_, Qc_[d].Data());
        std::copy(qv.begin() + static_cast<size_t>(d) * ne_, qv.begin() + static_cast<size_t>(d + 1) * ne_, Qv_[d].GetData());
    }
}

void MacroscopicQuantities::Reset()
{
    Tc_ = 0.0;
    Tv_ = 0.0;
    for (int d = 0; d < dim_; ++d)
    {
        Qc_[d] = 0.0;
        Qv_[d] = 0.0;
    }
    if (dev_) ThrowOnStatus(pbte_cuda_macro_reset(dev_), "pbte_cuda_macro_reset");
}

void MacroscopicQuantities::AccumulateDirectionalCoeff(int dir_idx, int branch, int spec, const mfem::DenseMatrix &coeff)
{
    MFEM_VERIFY(coeff.Height() == ndof_ && coeff.Width() == ne_, "coefficient block shape mismatch");
    EnsureDevice();
    ThrowOnStatus(pbte_cuda_macro_accumulate(dev_, dir_idx, branch * nspec_ + spec, coeff.Data()), "pbte_cuda_macro_accumulate");
    // Tc/Qc mirrors are refreshed by Finalize (the reference's callers read them only after Finalize)
}

void MacroscopicQuantities::Finalize(const std::vector<ElementIntegralData> &elem_data)
{
    EnsureDevice();
    std::vector<double> basis(static_cast<size_t>(ne_) * ndof_);
    for (int e = 0; e < ne_; ++e)
    {
        MFEM_VERIFY(elem_data[e].basis_integrals.Size() == ndof_, "basis_integrals size mismatch");
        std::copy(elem_data[e].basis_integrals.begin(), elem_data[e].basis_integrals.end(), basis.begin() + static_cast<size_t>(e) * ndof_);
    }
    ThrowOnStatus(pbte_cuda_macro_finalize(dev_, basis.data()), "pbte_cuda_macro_finalize");
    PullFromDevice(dev_);
}

double MacroscopicQuantities::Residual(const mfem::Vector &prev_Tv) const
{
    MFEM_VERIFY(prev_Tv.Size() == ne_, "prev_Tv size mismatch");
    EnsureDevice();
    double res = 0.0;
    // the device context may be fresh (moments set by PBTESolver::Solve): push the current Tv first
    ThrowOnStatus(pbte_cuda_upload_moments(dev_, nullptr, Tv_.GetData()), "pbte_cuda_upload_moments");
    const int st = pbte_cuda_macro_residual(dev_, prev_Tv.GetData(), &res);
    if (st == PBTE_ERR_ZERO_NORM) MFEM_VERIFY(false, "Tv norm is zero; cannot compute residual.");
    ThrowOnStatus(st, "pbte_cuda_macro_residual");
    return res;
}

// Output helpers (post-processing, host side; SURVEY §2 row 4b marks them out of the hot path).
// Field names and the Q packing follow the reference (src/MacroscopicQuantities.cpp:168-270), including
// its byNODES interleaving so that the written arrays stay comparable.
void MacroscopicQuantities::WriteParaView(const std::string &prefix, bool high_order) const
{
    mfem::Mesh *mesh = fes_.GetMesh();
    const mfem::FiniteElementCollection *fec = fes_.FEColl();
    MFEM_VERIFY(fec != nullptr, "FEColl is null");
    mfem::GridFunction T(const_cast<mfem::FiniteElementSpace *>(&fes_));
    mfem::FiniteElementSpace vfes(mesh, fec, dim_, fes_.GetOrdering());
    mfem::GridFunction Q(&vfes);
    mfem::Array<int> vdofs;
    mfem::Vector loc(dim_ * ndof_);
    for (int e = 0; e < ne_; ++e)
    {
        fes_.GetElementVDofs(e, vdofs);
        for (int j = 0; j < vdofs.Size(); ++j) T(vdofs[j]) = Tc_(j, e);
        vfes.GetElementVDofs(e, vdofs);
        for (int i = 0; i < ndof_; ++i)
            for (int d = 0; d < dim_; ++d)
                loc[fes_.GetOrdering() == mfem::Ordering::byNODES ? i * dim_ + d : d * ndof_ + i] = Qc_[d](i, e);
        Q.SetSubVector(vdofs, loc);
    }
    mfem::ParaViewDataCollection dc(prefix, mesh);
    dc.SetPrefixPath("output/vis");
    dc.SetDataFormat(mfem::VTKFormat::BINARY);
    dc.SetHighOrderOutput(high_order);
    dc.SetLevelsOfDetail(1);
    dc.RegisterField("T", &T);
    dc.RegisterField("Q", &Q);
    dc.Save();
}

void MacroscopicQuantities::Write2DSliceTemperature(const std::string &out_path, int Nx, int Ny, double clamp_tol) const
{
    mfem::Mesh &mesh = *fes_.GetMesh();
    MFEM_VERIFY(mesh.SpaceDimension() == 2, "Write2DSliceTemperature currently supports 2D meshes only.");
    MFEM_VERIFY(Nx >= 2 && Ny >= 2, "Nx/Ny must be >= 2.");
    double lo[2] = {std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()};
    double hi[2] = {-lo[0], -lo[1]};
    for (int v = 0; v < mesh.GetNV(); ++v)
        for (int d = 0; d < 2; ++d)
        {
            lo[d] = std::min(lo[d], mesh.GetVertex(v)[d]);
            hi[d] = std::max(hi[d], mesh.GetVertex(v)[d]);
        }
    namespace fs = std::filesystem;
    const fs::path p(out_path);
    if (p.has_parent_path()) fs::create_directories(p.parent_path());
    std::ofstream os(p);
    MFEM_VERIFY(static_cast<bool>(os), "Failed to open output file for sampling: " + p.string());
    os.setf(std::ios::fixed);
    os.precision(16);
    os << "# nx " << Nx << " ny " << Ny << "\nx y T\n";
    mfem::Vector pt(2);
    for (int j = 0; j < Ny; ++j)
    {
        const double y = lo[1] + static_cast<double>(j) / (Ny - 1) * (hi[1] - lo[1]);
        for (int i = 0; i < Nx; ++i)
        {
            const double x = lo[0] + static_cast<double>(i) / (Nx - 1) * (hi[0] - lo[0]);
            pt[0] = i == 0 ? lo[0] + clamp_tol : (i == Nx - 1 ? hi[0] - clamp_tol : x);
            pt[1] = j == 0 ? lo[1] + clamp_tol : (j == Ny - 1 ? hi[1] - clamp_tol : y);
            double T = std::numeric_limits<double>::quiet_NaN();
            for (int e = 0; e < ne_; ++e)  // first containing element in ascending id, as the reference's brute-force locator
            {
                mfem::InverseElementTransformation inv(mesh.GetElementTransformation(e));
                inv.SetMaxIter(64);
                inv.SetReferenceTol(1e-10);
                inv.SetPhysicalRelTol(1e-10);
                mfem::IntegrationPoint ip;
                if (inv.Transform(pt, ip) != mfem::InverseElementTransformation::Inside) continue;
                const mfem::FiniteElement *fe = fes_.GetFE(e);
                mfem::Vector shape(fe->GetDof());
                fe->CalcShape(ip, shape);
                T = 0.0;
                for (int k = 0; k < fe->GetDof(); ++k) T += Tc_(k, e) * shape[k];
                break;
            }
            os << x << " " << y << " " << T << "\n";
        }
    }
    os << std::flush;
    std::cout << "2D temperature slice written to: " << p << std::endl;
}
}  // namespace pbte
