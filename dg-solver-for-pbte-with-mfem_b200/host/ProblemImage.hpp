// Flattening of the host objects (mesh topology, local matrices, direction set, band tables, wall
// temperatures) into the POD problem image of include/pbte_cuda.h.  This is the only place where
// MFEM types meet the C ABI.
#pragma once

#include "AngularQuadrature.hpp"
#include "ElementIntegrator.hpp"
#include "PhononProperties.hpp"
#include "pbte_cuda.h"

#include "mfem.hpp"

#include <map>
#include <vector>

namespace pbte
{
class ProblemImage
{
public:
    /// Topology + directions only (enough for pbte_cuda_build_schedule): used by AngularSweepOrder::Build.
    static ProblemImage Topology(const mfem::Mesh &mesh, const AngleQuadrature &quad);
    /// Directions + band tables + sizes only (enough for the MacroscopicQuantities entry points).
    static ProblemImage MomentsOnly(int dim, int ne, int ndof, const AngleQuadrature &quad, const PhononProperties &props);
    /// Everything PBTESolver needs.  Boundary faces whose attribute is missing from `isothermal_bcs`
    /// get a NaN wall temperature; MissingBoundaryAttr() reports the first such attribute.
    static ProblemImage Full(const mfem::Mesh &mesh, const AngleQuadrature &quad, const std::vector<ElementIntegralData> &elem_data,
                             const PhononProperties &props, const std::map<int, double> &isothermal_bcs);

    const pbte_problem_t &View() const { return view_; }
    int MissingBoundaryAttr() const { return missing_attr_; }
    void SetBasisIntegrals(const std::vector<ElementIntegralData> &elem_data);

    ProblemImage() = default;
    ProblemImage(const ProblemImage &o) { *this = o; }
    ProblemImage &operator=(const ProblemImage &o);

private:
    void AddDirections(const AngleQuadrature &quad);
    void AddBands(const PhononProperties &props);
    void AddTopology(const mfem::Mesh &mesh);
    void Rebind();

    pbte_problem_t view_{};
    int missing_attr_ = -1;
    std::vector<double> basis_, mass_, stiff_, fmass_, fcoup_, fint_, fnormal_, ftbc_, dir_, dirw_, invkn_, vg_, hc_, fw_;
    std::vector<int32_t> fnbr_;
};
}  // namespace pbte
