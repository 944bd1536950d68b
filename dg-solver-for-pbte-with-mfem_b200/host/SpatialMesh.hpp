// Mesh loading / refinement / scaling and the L2 (DG) space.  API-compatible with the reference's
// include/SpatialMesh.hpp:19-120 (serial part); MFEM is used for mesh and FE-space setup only.
#pragma once

#include "mfem.hpp"

#include <map>
#include <memory>
#include <string>

namespace pbte
{
class SpatialMesh
{
public:
    SpatialMesh() = default;

    /// File path, or one of the built-in names unit-square[-tri|-quad] (8x8) / unit-cube[-tet|-hex] (4x4x4).
    void LoadMesh(const std::string &path_or_builtin);
    /// Reads `mesh: path:` (or `mesh_path:`) and the `boundary_conditions:` list of {attr, temperature}.
    void LoadMeshFromConfig(const std::string &config_path);
    void BuildDGSpace(int order, mfem::Ordering::Type ordering = mfem::Ordering::byNODES, const std::string &log_path = "");
    void UniformRefine(int levels);
    void ScaleCoordinates(double factor);
    /// Programmatic alternative to the config list (attr -> wall temperature).
    void SetIsothermalBoundary(int attr, double temperature) { isothermal_bc_[attr] = temperature; }

    mfem::Mesh &Mesh() { return *mesh_; }
    const mfem::Mesh &Mesh() const { return *mesh_; }
    mfem::FiniteElementSpace &FESpace() { return *fes_; }
    const mfem::FiniteElementSpace &FESpace() const { return *fes_; }
    int Dimension() const { return mesh_ ? mesh_->Dimension() : 0; }
    const std::map<int, double> &IsothermalBoundaryTemps() const { return isothermal_bc_; }

private:
    std::string Summary() const;
    std::string source_;
    int order_ = -1;
    std::unique_ptr<mfem::Mesh> mesh_;
    std::unique_ptr<mfem::FiniteElementCollection> fec_;
    std::unique_ptr<mfem::FiniteElementSpace> fes_;
    std::map<int, double> isothermal_bc_;
};
}  // namespace pbte
