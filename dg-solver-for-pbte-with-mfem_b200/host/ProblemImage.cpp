#include "ProblemImage.hpp"

#include "Utils.hpp"

#include <algorithm>
#include <cmath>
#include <limits>

namespace pbte
{
ProblemImage &ProblemImage::operator=(const ProblemImage &o)
{
    view_ = o.view_;
    missing_attr_ = o.missing_attr_;
    basis_ = o.basis_; mass_ = o.mass_; stiff_ = o.stiff_; fmass_ = o.fmass_; fcoup_ = o.fcoup_; fint_ = o.fint_;
    fnormal_ = o.fnormal_; ftbc_ = o.ftbc_; dir_ = o.dir_; dirw_ = o.dirw_; invkn_ = o.invkn_; vg_ = o.vg_; hc_ = o.hc_; fw_ = o.fw_;
    fnbr_ = o.fnbr_;
    Rebind();
    return *this;
}

void ProblemImage::Rebind()
{
    auto ptr = [](const std::vector<double> &v) { return v.empty() ? nullptr : v.data(); };
    view_.abi_version = PBTE_ABI_VERSION;
    view_.basis_integrals = ptr(basis_);
    view_.mass = ptr(mass_);
    view_.stiff = ptr(stiff_);
    view_.face_mass = ptr(fmass_);
    view_.face_coupling = ptr(fcoup_);
    view_.face_integral = ptr(fint_);
    view_.face_normal = ptr(fnormal_);
    view_.face_tbc = ptr(ftbc_);
    view_.face_nbr = fnbr_.empty() ? nullptr : fnbr_.data();
    view_.dir = ptr(dir_);
    view_.dir_w = ptr(dirw_);
    view_.inv_kn = ptr(invkn_);
    view_.vg = ptr(vg_);
    view_.heat_cap = ptr(hc_);
    view_.freq_w = ptr(fw_);
}

void ProblemImage::AddDirections(const AngleQuadrature &quad)
{
    const auto &dirs = quad.Directions();
    view_.ndir = static_cast<int32_t>(dirs.size());
    view_.omega = quad.TotalWeight();
    dir_.resize(3 * dirs.size());
    dirw_.resize(dirs.size());
    for (size_t k = 0; k < dirs.size(); ++k)
    {
        for (int d = 0; d < 3; ++d) dir_[3 * k + d] = dirs[k].direction[d];
        dirw_[k] = dirs[k].weight;
    }
}

void ProblemImage::AddBands(const PhononProperties &props)
{
    const int nbranch = static_cast<int>(props.Frequency().size());
    const int nspec = nbranch > 0 ? static_cast<int>(props.Frequency(0).size()) : 0;
    view_.nband = nbranch * nspec;
    view_.heat_cap_v = props.avgHeatCapacity();
    for (int b = 0; b < nbranch; ++b)
        for (int s = 0; s < nspec; ++s)
        {
            invkn_.push_back(props.InvKn(b)[s]);
            vg_.push_back(props.GroupVelocity(b)[s]);
            hc_.push_back(props.HeatCapacity(b)[s]);
            fw_.push_back(props.FrequencyWeight(b)[s]);
        }
}

// Element-local face slots in ascending global face id — the order in which the face loop of the
// integrator appends to ElementIntegralData (and hence the order of the per-face matrices).
void ProblemImage::AddTopology(const mfem::Mesh &mesh)
{
    const int ne = mesh.GetNE(), nfaces = mesh.GetNumFaces();
    view_.dim = mesh.Dimension();
    view_.ne = ne;
    std::vector<int> count(ne, 0);
    for (int f = 0; f < nfaces; ++f)
    {
        int e1 = -1, e2 = -1;
        mesh.GetFaceElements(f, &e1, &e2);
        if (e1 >= 0) ++count[e1];
        if (e2 >= 0) ++count[e2];
    }
    const int F = ne > 0 ? *std::max_element(count.begin(), count.end()) : 0;
    MFEM_VERIFY(F <= PBTE_MAX_FACES, "element with more than PBTE_MAX_FACES faces");
    view_.nfaces = F;
    fnbr_.assign(static_cast<size_t>(ne) * F, PBTE_FACE_NONE);
    fnormal_.assign(static_cast<size_t>(ne) * F * 3, 0.0);
    std::fill(count.begin(), count.end(), 0);
    for (int f = 0; f < nfaces; ++f)
    {
        int el[2] = {-1, -1};
        mesh.GetFaceElements(f, &el[0], &el[1]);
        for (int side = 0; side < 2; ++side)
        {
            const int e = el[side];
            if (e < 0) continue;
            const size_t slot = static_cast<size_t>(e) * F + count[e]++;
            fnbr_[slot] = el[1 - side] >= 0 ? el[1 - side] : PBTE_FACE_BOUNDARY;
            const mfem::Vector n = utils::ComputeOutwardFaceNormal(mesh, f, e);
            for (int d = 0; d < n.Size(); ++d) fnormal_[3 * slot + d] = n[d];
        }
    }
}

ProblemImage ProblemImage::Topology(const mfem::Mesh &mesh, const AngleQuadrature &quad)
{
    ProblemImage p;
    p.AddTopology(mesh);
    p.AddDirections(quad);
    p.Rebind();
    return p;
}

ProblemImage ProblemImage::MomentsOnly(int dim, int ne, int ndof, const AngleQuadrature &quad, const PhononProperties &props)
{
    ProblemImage p;
    p.view_.dim = dim;
    p.view_.ne = ne;
    p.view_.ndof = ndof;
    p.AddDirections(quad);
    p.AddBands(props);
    p.Rebind();
    return p;
}

void ProblemImage::SetBasisIntegrals(const std::vector<ElementIntegralData> &elem_data)
{
    const int ne = view_.ne, n = view_.ndof;
    MFEM_VERIFY(static_cast<int>(elem_data.size()) == ne, "element data size mismatch");
    basis_.assign(static_cast<size_t>(ne) * n, 0.0);
    for (int e = 0; e < ne; ++e)
    {
        MFEM_VERIFY(elem_data[e].basis_integrals.Size() == n, "basis_integrals size mismatch");
        std::copy(elem_data[e].basis_integrals.begin(), elem_data[e].basis_integrals.end(), basis_.begin() + static_cast<size_t>(e) * n);
    }
    Rebind();
}

ProblemImage ProblemImage::Full(const mfem::Mesh &mesh, const AngleQuadrature &quad, const std::vector<ElementIntegralData> &elem_data,
                                const PhononProperties &props, const std::map<int, double> &bcs)
{
    ProblemImage p;
    p.AddTopology(mesh);
    p.AddDirections(quad);
    p.AddBands(props);
    const int ne = p.view_.ne, F = p.view_.nfaces, dim = p.view_.dim;
    MFEM_VERIFY(static_cast<int>(elem_data.size()) == ne, "element data size mismatch");
    const int n = ne > 0 ? elem_data[0].mass_matrix.Height() : 0;
    p.view_.ndof = n;
    const size_t nn = static_cast<size_t>(n) * n;
    p.basis_.assign(static_cast<size_t>(ne) * n, 0.0);
    p.mass_.assign(ne * nn, 0.0);
    p.stiff_.assign(ne * dim * nn, 0.0);
    p.fmass_.assign(ne * F * nn, 0.0);
    p.fcoup_.assign(ne * F * nn, 0.0);
    p.fint_.assign(static_cast<size_t>(ne) * F * n, 0.0);
    p.ftbc_.assign(static_cast<size_t>(ne) * F, 0.0);
    for (int e = 0; e < ne; ++e)
    {
        const ElementIntegralData &ed = elem_data[e];
        MFEM_VERIFY(ed.mass_matrix.Height() == n && ed.mass_matrix.Width() == n, "non-uniform ndof is not supported");
        MFEM_VERIFY(static_cast<int>(ed.stiffness_matrices.size()) == dim, "stiffness matrix count != mesh dimension");
        std::copy(ed.basis_integrals.begin(), ed.basis_integrals.end(), p.basis_.begin() + static_cast<size_t>(e) * n);
        std::copy(ed.mass_matrix.Data(), ed.mass_matrix.Data() + nn, p.mass_.begin() + e * nn);
        for (int d = 0; d < dim; ++d)
            std::copy(ed.stiffness_matrices[d].Data(), ed.stiffness_matrices[d].Data() + nn, p.stiff_.begin() + (static_cast<size_t>(e) * dim + d) * nn);
        const int nf = static_cast<int>(ed.face_couplings.size());
        MFEM_VERIFY(nf <= F, "more face couplings than mesh faces on an element");
        for (int f = 0; f < nf; ++f)
        {
            const size_t slot = static_cast<size_t>(e) * F + f;
            const auto &fc = ed.face_couplings[f];
            const int expect = p.fnbr_[slot];
            MFEM_VERIFY((fc.neighbor_elem >= 0 ? fc.neighbor_elem : PBTE_FACE_BOUNDARY) == expect,
                        "face_couplings order does not match ascending global face id");
            std::copy(ed.face_mass_matrices[f].Data(), ed.face_mass_matrices[f].Data() + nn, p.fmass_.begin() + slot * nn);
            if (fc.neighbor_elem >= 0)
            {
                MFEM_VERIFY(fc.coupling.Height() == n && fc.coupling.Width() == n, "coupling matrix shape mismatch");
                std::copy(fc.coupling.Data(), fc.coupling.Data() + nn, p.fcoup_.begin() + slot * nn);
            }
            else
            {
                std::copy(fc.isothermal_rhs.begin(), fc.isothermal_rhs.end(), p.fint_.begin() + slot * n);
                auto it = bcs.find(fc.boundary_attr);
                if (it != bcs.end()) p.ftbc_[slot] = it->second;
                else
                {
                    p.ftbc_[slot] = std::numeric_limits<double>::quiet_NaN();
                    if (p.missing_attr_ < 0) p.missing_attr_ = fc.boundary_attr;
                }
            }
        }
    }
    p.Rebind();
    return p;
}
}  // namespace pbte
