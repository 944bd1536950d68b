// Discrete-ordinates direction set (polar x azimuth product rule).
// API-compatible with the reference's include/AngularQuadrature.hpp:8-82 (same type and member names,
// same argument meaning and exceptions); the values are inputs to the GPU sweep, built on the host.
#pragma once

#include <array>
#include <string>
#include <vector>

namespace pbte
{
enum class AngleDiscretizationType { Uniform, GaussLegendre };

struct AngleDiscretizationOptions
{
    int dimension = 3;       // 2: in-plane directions only, 3: full sphere
    int polar_points = 8;    // nodes in mu = cos(theta); ignored for dimension 2
    int azimuth_points = 16; // nodes in phi
    AngleDiscretizationType polar_scheme = AngleDiscretizationType::GaussLegendre;
    AngleDiscretizationType azimuth_scheme = AngleDiscretizationType::Uniform;
};

struct AngleDirection
{
    double polar = 0.0;    // theta
    double azimuth = 0.0;  // phi
    double weight = 0.0;   // solid-angle weight (sums to 2*pi in 2-D, 4*pi in 3-D)
    std::array<double, 3> direction{0.0, 0.0, 0.0};
};

class AngleQuadrature
{
public:
    static AngleQuadrature Build(const AngleDiscretizationOptions &opts);
    static AngleDiscretizationOptions LoadOptionsFromConfig(const std::string &config_path);
    static AngleDiscretizationType ParseSchemeName(const std::string &name);
    static std::string SchemeName(AngleDiscretizationType type);
    void WriteToFile(const std::string &path) const;

    int Dimension() const { return dimension_; }
    double TotalWeight() const { return total_weight_; }
    const std::vector<double> &PolarAngles() const { return polar_angles_; }
    const std::vector<double> &PolarWeights() const { return polar_weights_; }
    const std::vector<double> &AzimuthAngles() const { return azimuth_angles_; }
    const std::vector<double> &AzimuthWeights() const { return azimuth_weights_; }
    const std::vector<AngleDirection> &Directions() const { return directions_; }

private:
    int dimension_ = 0;
    double total_weight_ = 0.0;
    std::vector<double> polar_angles_, polar_weights_, azimuth_angles_, azimuth_weights_;
    std::vector<AngleDirection> directions_;
};
}  // namespace pbte
