#include "AngularSweepOrder.hpp"

#include "CudaCheck.hpp"
#include "ProblemImage.hpp"

#include <filesystem>
#include <fstream>

namespace pbte
{
AngularSweepOrder AngularSweepOrder::Build(const mfem::Mesh &mesh, const AngleQuadrature &angle_quad)
{
    AngularSweepOrder result;
    const int ne = mesh.GetNE();
    const int ndir = static_cast<int>(angle_quad.Directions().size());
    if (ne == 0 || ndir == 0) return result;  // reference: empty object (src/AngularSweepOrder.cpp:83-86)

    const ProblemImage image = ProblemImage::Topology(mesh, angle_quad);
    pbte_options_t opt;
    pbte_cuda_default_options(&opt);
    pbte_handle_t *h = nullptr;
    ThrowOnStatus(pbte_cuda_create(&image.View(), &opt, &h), "pbte_cuda_create");
    struct Guard
    {
        pbte_handle_t *h;
        ~Guard() { pbte_cuda_destroy(h); }
    } guard{h};
    ThrowOnStatus(pbte_cuda_build_schedule(h), "pbte_cuda_build_schedule");
    result.orders_.assign(ndir, std::vector<int>(ne, -1));
    result.levels_.assign(ndir, std::vector<int>(ne, -1));
    for (int k = 0; k < ndir; ++k)
    {
        ThrowOnStatus(pbte_cuda_get_order(h, k, result.orders_[k].data()), "pbte_cuda_get_order");
        ThrowOnStatus(pbte_cuda_get_levels(h, k, result.levels_[k].data()), "pbte_cuda_get_levels");
    }
    int32_t max_levels = 0;
    ThrowOnStatus(pbte_cuda_schedule_info(h, &result.n_classes_, &max_levels), "pbte_cuda_schedule_info");
    return result;
}

// Text format of the reference's log (src/AngularSweepOrder.cpp:149-181).
void AngularSweepOrder::WriteToFile(const AngleQuadrature &angle_quad, const mfem::Mesh &mesh, const std::string &path) const
{
    namespace fs = std::filesystem;
    const fs::path p(path);
    if (p.has_parent_path()) fs::create_directories(p.parent_path());
    std::ofstream os(p);
    if (!os) throw std::runtime_error("Failed to open sweep log: " + path);
    os << "Sweep order per direction\ndimension: " << mesh.Dimension() << "\nelements: " << mesh.GetNE()
       << "\ndirections: " << angle_quad.Directions().size() << "\n\n";
    for (int k = 0; k < NumDirections(); ++k)
    {
        const AngleDirection &d = angle_quad.Directions()[k];
        os << "dir " << k << " theta=" << d.polar << " phi=" << d.azimuth << " w=" << d.weight << " order:";
        for (int e : orders_[k]) os << " " << e;
        os << "\n";
    }
    os << std::flush;
}
}  // namespace pbte
