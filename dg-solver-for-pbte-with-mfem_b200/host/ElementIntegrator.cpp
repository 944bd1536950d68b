// Quadrature assembly of the local matrices.  Semantics follow the reference's
// src/ElementIntegrator.cpp: volume rule of order 2p+1 (:33-41,:79-80), accumulation
// w * phi_i * phi_j and w * dphi_i/dx_d * phi_j (:84-112); face loop over global faces in ascending
// id, both sides, face rule of order p1+p2+1 (:43-55,:118-257); boundary attribute per face from the
// boundary elements (:260-274).
#include "ElementIntegrator.hpp"

namespace pbte
{
DGElementIntegrator::DGElementIntegrator(const mfem::FiniteElementSpace &fes, int quadrature_order)
    : fes_(fes), mesh_(*fes.GetMesh()), quadrature_order_(quadrature_order), rules_(0, mfem::Quadrature1D::GaussLegendre)
{
}

void DGElementIntegrator::AssembleVolume(int e, ElementIntegralData &out)
{
    const mfem::FiniteElement &fe = *fes_.GetFE(e);
    mfem::ElementTransformation &tr = *mesh_.GetElementTransformation(e);
    const int n = fe.GetDof(), dim = mesh_.Dimension();
    out.basis_integrals.SetSize(n);
    out.basis_integrals = 0.0;
    out.mass_matrix.SetSize(n);
    out.mass_matrix = 0.0;
    out.stiffness_matrices.assign(dim, mfem::DenseMatrix(n));
    for (mfem::DenseMatrix &g : out.stiffness_matrices) g = 0.0;

    const int order = quadrature_order_ > 0 ? quadrature_order_ : 2 * fe.GetOrder() + 1;
    const mfem::IntegrationRule &ir = rules_.Get(fe.GetGeomType(), order);
    mfem::Vector phi(n);
    mfem::DenseMatrix grad(n, dim);
    for (int q = 0; q < ir.GetNPoints(); ++q)
    {
        const mfem::IntegrationPoint &ip = ir.IntPoint(q);
        tr.SetIntPoint(&ip);
        const double w = ip.weight * tr.Weight();
        fe.CalcShape(ip, phi);
        fe.CalcPhysDShape(tr, grad);
        for (int i = 0; i < n; ++i)
        {
            out.basis_integrals(i) += w * phi(i);
            for (int j = 0; j < n; ++j)
            {
                out.mass_matrix(i, j) += w * phi(i) * phi(j);
                for (int d = 0; d < dim; ++d) out.stiffness_matrices[d](i, j) += w * grad(i, d) * phi(j);
            }
        }
    }
}

void DGElementIntegrator::AssembleFaces(std::vector<ElementIntegralData> &all)
{
    const int nfaces = mesh_.GetNumFaces();
    std::vector<int> attr_of_face(nfaces, 0);
    for (int be = 0; be < mesh_.GetNBE(); ++be)
    {
        const int f = mesh_.GetBdrElementFaceIndex(be);
        if (f >= 0 && f < nfaces) attr_of_face[f] = mesh_.GetBdrAttribute(be);
    }
    const mfem::IntegrationPoint origin;  // (0,0,0): the point the reference's face loop evaluates at

    for (int f = 0; f < nfaces; ++f)
    {
        mfem::FaceElementTransformations *ftr = mesh_.GetFaceElementTransformations(f);
        if (!ftr) continue;
        const int elem[2] = {ftr->Elem1No, ftr->Elem2No};
        for (int side = 0; side < 2; ++side)
        {
            const int e = elem[side], nb = elem[1 - side];
            if (e < 0) continue;
            const mfem::FiniteElement &fe = *fes_.GetFE(e);
            const mfem::FiniteElement *fe_nb = nb >= 0 ? fes_.GetFE(nb) : nullptr;
            const int n = fe.GetDof(), n_nb = fe_nb ? fe_nb->GetDof() : 0;
            const int order = quadrature_order_ > 0 ? quadrature_order_ : fe.GetOrder() + (fe_nb ? fe_nb->GetOrder() : fe.GetOrder()) + 1;
            const mfem::IntegrationRule &ir = rules_.Get(ftr->GetGeometryType(), order);

            mfem::DenseMatrix F(n), K;
            mfem::Vector g(n), phi(n), phi_nb(n_nb);
            F = 0.0;
            g = 0.0;
            if (fe_nb)
            {
                K.SetSize(n, n_nb);
                K = 0.0;
            }
            for (int q = 0; q < ir.GetNPoints(); ++q)
            {
                const mfem::IntegrationPoint &ipf = ir.IntPoint(q);
                const mfem::IntegrationPoint *ip_own = &origin, *ip_nb = &origin;
                if (face_rule_ == FacePointRule::Consistent)
                {
                    ftr->SetAllIntPoints(&ipf);
                    ip_own = side == 0 ? &ftr->GetElement1IntPoint() : &ftr->GetElement2IntPoint();
                    ip_nb = side == 0 ? &ftr->GetElement2IntPoint() : &ftr->GetElement1IntPoint();
                }
                else ftr->Face->SetIntPoint(&ipf);
                fe.CalcShape(*ip_own, phi);
                if (fe_nb) fe_nb->CalcShape(*ip_nb, phi_nb);
                const double w = ipf.weight * ftr->Face->Weight();
                for (int i = 0; i < n; ++i)
                {
                    g(i) += w * phi(i);
                    for (int j = 0; j < n; ++j) F(i, j) += w * phi(i) * phi(j);
                    for (int j = 0; j < n_nb; ++j) K(i, j) += w * phi(i) * phi_nb(j);
                }
            }
            ElementIntegralData &ed = all[e];
            ElementIntegralData::FaceCoupling fc;
            fc.face_id = f;
            fc.boundary_attr = attr_of_face[f];
            fc.neighbor_elem = nb >= 0 ? nb : -1;
            if (nb >= 0) fc.coupling = std::move(K);
            else fc.isothermal_rhs = g;
            ed.face_mass_matrices.push_back(std::move(F));
            ed.face_integrals.push_back(std::move(g));
            ed.face_couplings.push_back(std::move(fc));
        }
    }
}

std::vector<ElementIntegralData> DGElementIntegrator::AssembleAll()
{
    std::vector<ElementIntegralData> all(mesh_.GetNE());
    for (int e = 0; e < mesh_.GetNE(); ++e) AssembleVolume(e, all[e]);
    AssembleFaces(all);
    return all;
}
}  // namespace pbte
