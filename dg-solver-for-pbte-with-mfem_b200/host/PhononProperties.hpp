// Silicon LA/TA band tables (dispersion, group velocity, scattering rate, mode heat capacity).
// API-compatible with the reference's include/PhononProperties.hpp:8-84; the 2*num_spectral scalars
// per table are inputs to the GPU sweep (SURVEY Appendix A lists the formulas).
#pragma once

#include <string>
#include <vector>

namespace pbte
{
struct PhononMaterial
{
    int num_branches = 0;
    std::vector<double> C_LA, C_TA;  // omega = c0 k + c1 k^2
    double lattice_dist = 0.0;
    std::vector<double> K_range;     // [0, k_max]
    double Ai = 0.0, BL = 0.0, BT = 0.0, BU = 0.0;
    int num_spectral = 0;
    double ref_temp = 0.0;
    double ref_len = 0.0;
};

class PhononProperties
{
public:
    typedef std::vector<std::vector<double>> Table;  // [branch][spectral bin]

    static PhononMaterial LoadMaterial(const std::string &path);
    static PhononProperties Build(const PhononMaterial &mat);
    void WriteToFile(const std::string &path) const;

    const Table &WaveVector() const { return k_; }
    const std::vector<double> &WaveVector(int b) const { return k_[b]; }
    const Table &Frequency() const { return w_; }
    const std::vector<double> &Frequency(int b) const { return w_[b]; }
    const Table &FrequencyWeight() const { return dw_; }
    const std::vector<double> &FrequencyWeight(int b) const { return dw_[b]; }
    const Table &GroupVelocity() const { return vg_; }
    const std::vector<double> &GroupVelocity(int b) const { return vg_[b]; }
    const Table &InvKn() const { return inv_kn_; }
    const std::vector<double> &InvKn(int b) const { return inv_kn_[b]; }
    const Table &DensityOfStates() const { return density_; }
    const std::vector<double> &DensityOfStates(int b) const { return density_[b]; }
    const Table &HeatCapacity() const { return heat_cap_; }
    const std::vector<double> &HeatCapacity(int b) const { return heat_cap_[b]; }
    double avgHeatCapacity() const { return heat_cap_v_; }

private:
    int num_branches_ = 0, num_spectral_ = 0;
    double k_max_ = 0.0, ref_temp_ = 0.0, ref_len_ = 0.0, heat_cap_v_ = 0.0;
    Table k_, w_, dw_, vg_, inv_kn_, density_, heat_cap_;
};
}  // namespace pbte
