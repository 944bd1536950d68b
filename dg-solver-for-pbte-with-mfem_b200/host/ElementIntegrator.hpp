// Local DG matrices of every element: the data contract between the FE setup and the GPU sweep.
// ElementIntegralData has the same fields as the reference's include/ElementIntegrator.hpp:11-35;
// DGElementIntegrator keeps its constructor and AssembleAll() (:43-47).
#pragma once

#include "mfem.hpp"

#include <vector>

namespace pbte
{
struct ElementIntegralData
{
    mfem::Vector basis_integrals;                       // m_i   = int_K phi_i
    mfem::DenseMatrix mass_matrix;                      // M_ij  = int_K phi_i phi_j
    std::vector<mfem::DenseMatrix> stiffness_matrices;  // G_d,ij = int_K d_d(phi_i) phi_j, d < dim
    std::vector<mfem::DenseMatrix> face_mass_matrices;  // F_f,ij = int_f phi_i phi_j (own side twice)
    std::vector<mfem::Vector> face_integrals;           // g_f,i  = int_f phi_i

    struct FaceCoupling
    {
        int face_id = -1;
        int neighbor_elem = -1;  // -1 boundary, -2 neighbour on another rank (never produced here)
        int boundary_attr = 0;
        bool is_shared = false;
        mfem::DenseMatrix coupling;   // K_f,ij = int_f phi_i(this) phi_j(neighbour)
        mfem::Vector isothermal_rhs;  // = g_f on boundary faces
    };
    std::vector<FaceCoupling> face_couplings;  // ascending global face id, aligned with the two lists above
};

/// How the element-side evaluation point of a face quadrature node is obtained.
enum class FacePointRule
{
    /// What the reference computes (src/ElementIntegrator.cpp:185-197): only the face
    /// transformation receives the quadrature point, so both element shapes are evaluated at the
    /// reference-element origin for every node.  Default, required for parity with the reference.
    ReferenceCompatible,
    /// Mathematically consistent trace integrals (face point mapped into each element).
    Consistent
};

class DGElementIntegrator
{
public:
    explicit DGElementIntegrator(const mfem::FiniteElementSpace &fes, int quadrature_order = -1);
    std::vector<ElementIntegralData> AssembleAll();
    const mfem::FiniteElementSpace &FESpace() const { return fes_; }
    void SetFacePointRule(FacePointRule r) { face_rule_ = r; }

private:
    void AssembleVolume(int e, ElementIntegralData &out);
    void AssembleFaces(std::vector<ElementIntegralData> &all);

    const mfem::FiniteElementSpace &fes_;
    mfem::Mesh &mesh_;
    int quadrature_order_;
    FacePointRule face_rule_ = FacePointRule::ReferenceCompatible;
    mfem::IntegrationRules rules_;
};
}  // namespace pbte
