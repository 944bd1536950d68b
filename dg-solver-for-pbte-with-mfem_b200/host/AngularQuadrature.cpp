// Direction set construction; behaviour follows the reference's src/AngularQuadrature.cpp:224-327
// (polar index outer, azimuth inner; weights rescaled to the exact solid angle) and its config
// scanner :171-222 (keys under an `angles:` block).
#include "AngularQuadrature.hpp"

#include "mfem.hpp"

#include <cctype>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <stdexcept>

namespace pbte
{
namespace
{
const double kPi = 3.14159265358979323846;

struct Rule1D
{
    std::vector<double> x, w;
};

std::string Strip(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}

// midpoint rule: cells of width h, nodes at cell centres
Rule1D Midpoint(int n, double lo, double hi)
{
    if (n <= 0) throw std::invalid_argument("Uniform rule requires positive point count.");
    Rule1D r;
    const double h = (hi - lo) / static_cast<double>(n);
    for (int i = 0; i < n; ++i)
    {
        r.x.push_back(lo + (i + 0.5) * h);
        r.w.push_back(h);
    }
    return r;
}

// n-point Gauss-Legendre rule taken from MFEM's segment rule of order 2n-1 and mapped to [lo,hi]
Rule1D Gauss(int n, double lo, double hi)
{
    if (n <= 0) throw std::invalid_argument("Gauss-Legendre rule requires positive point count.");
    mfem::IntegrationRules table(0, mfem::Quadrature1D::GaussLegendre);
    const mfem::IntegrationRule &ir = table.Get(mfem::Geometry::SEGMENT, 2 * n - 1);
    double xmin = 1e300, xmax = -1e300;
    for (int i = 0; i < ir.GetNPoints(); ++i)
    {
        xmin = std::min(xmin, ir.IntPoint(i).x);
        xmax = std::max(xmax, ir.IntPoint(i).x);
    }
    // MFEM's reference segment is [0,1]; tolerate a [-1,1] convention as the reference does (:76-114)
    const bool unit = xmin >= -1e-12 && xmax <= 1.0 + 1e-12;
    Rule1D r;
    for (int i = 0; i < ir.GetNPoints(); ++i)
    {
        const mfem::IntegrationPoint &ip = ir.IntPoint(i);
        if (unit)
        {
            const double len = hi - lo;
            r.x.push_back(lo + len * ip.x);
            r.w.push_back(len * ip.weight);
        }
        else
        {
            const double half = 0.5 * (hi - lo), mid = 0.5 * (hi + lo);
            r.x.push_back(mid + half * ip.x);
            r.w.push_back(half * ip.weight);
        }
    }
    return r;
}

Rule1D MakeRule(AngleDiscretizationType t, int n, double lo, double hi)
{
    return t == AngleDiscretizationType::Uniform ? Midpoint(n, lo, hi) : Gauss(n, lo, hi);
}
}  // namespace

AngleDiscretizationType AngleQuadrature::ParseSchemeName(const std::string &name)
{
    std::string key = Strip(name);
    for (char &c : key) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    if (key == "uniform") return AngleDiscretizationType::Uniform;
    if (key == "gauss" || key == "gauss-legendre" || key == "legendre") return AngleDiscretizationType::GaussLegendre;
    throw std::invalid_argument("Unknown discretization scheme: " + name);
}

std::string AngleQuadrature::SchemeName(AngleDiscretizationType type)
{
    return type == AngleDiscretizationType::Uniform ? "uniform" : "gauss";
}

AngleDiscretizationOptions AngleQuadrature::LoadOptionsFromConfig(const std::string &config_path)
{
    std::ifstream in(config_path);
    if (!in) throw std::runtime_error("Failed to open config file: " + config_path);
    AngleDiscretizationOptions o;
    bool inside = false;
    std::string line;
    auto value_of = [](const std::string &l, const char *key, std::string &out) {
        const std::string k(key);
        if (l.compare(0, k.size(), k) != 0) return false;
        out = Strip(l.substr(k.size()));
        return true;
    };
    while (std::getline(in, line))
    {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        const std::string l = Strip(line);
        if (l.empty()) continue;
        if (l.compare(0, 7, "angles:") == 0)
        {
            inside = true;
            continue;
        }
        if (!inside) continue;
        std::string v;
        if (value_of(l, "dimension:", v)) o.dimension = std::stoi(v);
        else if (value_of(l, "polar_points:", v)) o.polar_points = std::stoi(v);
        else if (value_of(l, "azimuth_points:", v)) o.azimuth_points = std::stoi(v);
        else if (value_of(l, "polar_scheme:", v)) o.polar_scheme = ParseSchemeName(v);
        else if (value_of(l, "azimuth_scheme:", v)) o.azimuth_scheme = ParseSchemeName(v);
    }
    return o;
}

AngleQuadrature AngleQuadrature::Build(const AngleDiscretizationOptions &opts)
{
    if (opts.dimension != 2 && opts.dimension != 3) throw std::invalid_argument("Angular quadrature dimension must be 2 or 3.");
    if (opts.azimuth_points <= 0) throw std::invalid_argument("azimuth_points must be positive.");
    AngleQuadrature q;
    q.dimension_ = opts.dimension;

    Rule1D mu;  // nodes in cos(theta)
    if (opts.dimension == 2)
    {
        mu.x.push_back(0.0);
        mu.w.push_back(1.0);
    }
    else
    {
        if (opts.polar_points <= 0) throw std::invalid_argument("polar_points must be positive for 3D.");
        mu = MakeRule(opts.polar_scheme, opts.polar_points, -1.0, 1.0);
    }
    const Rule1D phi = MakeRule(opts.azimuth_scheme, opts.azimuth_points, 0.0, 2.0 * kPi);

    for (size_t i = 0; i < mu.x.size(); ++i)
    {
        q.polar_angles_.push_back(std::acos(std::max(-1.0, std::min(1.0, mu.x[i]))));
        q.polar_weights_.push_back(mu.w[i]);
    }
    q.azimuth_angles_ = phi.x;
    q.azimuth_weights_ = phi.w;

    double total = 0.0;
    for (size_t it = 0; it < q.polar_angles_.size(); ++it)
    {
        const double theta = q.polar_angles_[it];
        const double st = std::sin(theta), ct = std::cos(theta);
        for (size_t ip = 0; ip < phi.x.size(); ++ip)
        {
            AngleDirection d;
            d.polar = theta;
            d.azimuth = phi.x[ip];
            d.weight = q.polar_weights_[it] * phi.w[ip];
            d.direction = {st * std::cos(d.azimuth), st * std::sin(d.azimuth), opts.dimension == 3 ? ct : 0.0};
            total += d.weight;
            q.directions_.push_back(d);
        }
    }
    const double exact = opts.dimension == 3 ? 4.0 * kPi : 2.0 * kPi;
    q.total_weight_ = total;
    if (total > 0.0)
    {
        const double scale = exact / total;
        for (AngleDirection &d : q.directions_) d.weight *= scale;
        q.total_weight_ = exact;
    }
    return q;
}

void AngleQuadrature::WriteToFile(const std::string &path) const
{
    namespace fs = std::filesystem;
    const fs::path p(path);
    if (p.has_parent_path()) fs::create_directories(p.parent_path());
    std::ofstream os(p);
    if (!os) throw std::runtime_error("Failed to open angle log file: " + path);
    os << "Angular quadrature summary\n"
       << "  dimension        : " << dimension_ << "\n"
       << "  polar points     : " << polar_angles_.size() << "\n"
       << "  azimuth points   : " << azimuth_angles_.size() << "\n"
       << "  directions       : " << directions_.size() << "\n"
       << "  total weight     : " << total_weight_ << "\n\n"
       << "Directions (idx, theta, phi, weight, dir_x, dir_y, dir_z)\n";
    for (size_t i = 0; i < directions_.size(); ++i)
    {
        const AngleDirection &d = directions_[i];
        os << i << " " << d.polar << " " << d.azimuth << " " << d.weight << " " << d.direction[0] << " " << d.direction[1]
           << " " << d.direction[2] << "\n";
    }
}
}  // namespace pbte
