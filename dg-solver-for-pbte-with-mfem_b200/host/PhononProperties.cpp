// Band tables; formulas as stated in SURVEY Appendix A / the reference's src/PhononProperties.cpp:52-140.
#include "PhononProperties.hpp"

#include <yaml-cpp/yaml.h>

#include <cmath>
#include <filesystem>
#include <fstream>
#include <stdexcept>

namespace pbte
{
namespace
{
const double kPi = 3.14159265358979323846;
const double kHbar = 1.054571800e-34;
const double kBoltz = 1.38064852e-23;
}  // namespace

PhononMaterial PhononProperties::LoadMaterial(const std::string &path)
{
    const YAML::Node cfg = YAML::LoadFile(path);
    PhononMaterial m;
    m.C_LA = cfg["C_LA"].as<std::vector<double>>();
    m.C_TA = cfg["C_TA"].as<std::vector<double>>();
    m.lattice_dist = cfg["lattice_dist"].as<double>();
    m.num_branches = cfg["num_branches"].as<int>();
    if (m.num_branches < 2) throw std::invalid_argument("num_branches must be >= 2 (LA and TA)");
    m.K_range.assign(m.num_branches, 0.0);
    m.K_range[1] = 2.0 * kPi / m.lattice_dist;
    m.Ai = cfg["Ai"].as<double>();
    m.BL = cfg["BL"].as<double>();
    m.BT = cfg["BT"].as<double>();
    m.BU = cfg["BU"].as<double>();
    m.num_spectral = cfg["num_spectral"].as<int>();
    m.ref_temp = cfg["reference_temperature"].as<double>();
    m.ref_len = cfg["reference_length"].as<double>();
    return m;
}

PhononProperties PhononProperties::Build(const PhononMaterial &mat)
{
    PhononProperties ph;
    const int nb = mat.num_branches, ns = mat.num_spectral;
    const double kmax = mat.K_range[1], T0 = mat.ref_temp;
    ph.num_branches_ = nb;
    ph.num_spectral_ = ns;
    ph.k_max_ = kmax;
    ph.ref_temp_ = T0;
    ph.ref_len_ = mat.ref_len;
    for (Table *t : {&ph.k_, &ph.w_, &ph.dw_, &ph.vg_, &ph.inv_kn_, &ph.density_, &ph.heat_cap_})
        t->assign(nb, std::vector<double>(ns, 0.0));

    for (int branch = 0; branch < 2; ++branch)
    {
        const std::vector<double> &c = branch == 0 ? mat.C_LA : mat.C_TA;
        for (int j = 0; j < ns; ++j)
        {
            const double k = (2.0 * (j + 1) - 1.0) / (2.0 * ns) * kmax;  // midpoint grid in k
            const double w = c[0] * k + c[1] * k * k;
            const double vg = c[0] + 2.0 * c[1] * k;
            double nu = mat.Ai * std::pow(w, 4);
            if (branch == 0) nu += mat.BL * std::pow(T0, 3) * std::pow(w, 2);
            else if (k < kmax / 2.0) nu += mat.BT * w * std::pow(T0, 4);
            else nu += mat.BU * std::pow(w, 2) / std::sinh(kHbar * w / kBoltz / T0);
            ph.k_[branch][j] = k;
            ph.w_[branch][j] = w;
            ph.vg_[branch][j] = vg;
            ph.dw_[branch][j] = kmax * vg;
            ph.inv_kn_[branch][j] = nu;
            ph.density_[branch][j] = k * k / vg / 2.0 / std::pow(kPi, 2);
        }
    }
    ph.heat_cap_v_ = 0.0;
    for (int p = 0; p < nb; ++p)
        for (int j = 0; j < ns; ++j)
        {
            const double w = ph.w_[p][j];
            const double x = kHbar * w / kBoltz / T0;
            const double ex = std::exp(x), den = ex - 1.0;
            const double C = std::pow(kHbar, 2) * w * w * ph.density_[p][j] * ex / (den * den) / kBoltz / (T0 * T0);
            ph.heat_cap_[p][j] = C;
            ph.heat_cap_v_ += C * ph.inv_kn_[p][j] * (kmax * ph.vg_[p][j]);
        }
    return ph;
}

void PhononProperties::WriteToFile(const std::string &path) const
{
    namespace fs = std::filesystem;
    const fs::path p(path);
    if (p.has_parent_path()) fs::create_directories(p.parent_path());
    std::ofstream os(p);
    if (!os) throw std::runtime_error("Failed to open phonon log: " + path);
    os << "Phonon properties\nnum_branches: " << num_branches_ << "\nnum_spectral: " << num_spectral_ << "\nk_max: " << k_max_
       << "\nreference_temperature: " << ref_temp_ << "\nreference_length: " << ref_len_ << "\nHeatCapV: " << heat_cap_v_
       << "\n\nbranch idx k w dw vg invKn density heatCap\n";
    for (int p2 = 0; p2 < num_branches_; ++p2)
        for (int j = 0; j < num_spectral_; ++j)
            os << p2 << " " << j << " " << k_[p2][j] << " " << w_[p2][j] << " " << dw_[p2][j] << " " << vg_[p2][j] << " "
               << inv_kn_[p2][j] << " " << density_[p2][j] << " " << heat_cap_[p2][j] << "\n";
}
}  // namespace pbte
