// Per-direction upwind element order.  Same public API as the reference's
// include/AngularSweepOrder.hpp:17-40; Build() runs on the GPU through the C ABI
// (pbte_cuda_build_schedule / pbte_cuda_get_order) and returns the reference's list order bit-exactly.
#pragma once

#include "AngularQuadrature.hpp"
#include "mfem.hpp"

#include <string>
#include <utility>
#include <vector>

namespace pbte
{
class AngularSweepOrder
{
public:
    /// Throws std::runtime_error("Angular sweep ordering stalled...") on a cyclic upwind dependency.
    static AngularSweepOrder Build(const mfem::Mesh &mesh, const AngleQuadrature &angle_quad);

    int NumDirections() const { return static_cast<int>(orders_.size()); }
    const std::vector<int> &Order(int k) const { return orders_[k]; }
    const std::vector<std::vector<int>> &Orders() const { return orders_; }
    void WriteToFile(const AngleQuadrature &angle_quad, const mfem::Mesh &mesh, const std::string &path) const;

    /// Wavefront level of every element for direction k (the GPU execution schedule derived from the
    /// same DAG) and the number of distinct DAGs among the directions.
    const std::vector<int> &Levels(int k) const { return levels_[k]; }
    int NumScheduleClasses() const { return n_classes_; }

private:
    std::vector<std::vector<int>> orders_, levels_;
    int n_classes_ = 0;
};
}  // namespace pbte
