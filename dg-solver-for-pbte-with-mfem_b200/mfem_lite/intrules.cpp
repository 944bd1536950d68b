// mfem_lite: quadrature rules.
//   SEGMENT  : n-point Gauss–Legendre on [0,1], n = order/2 + 1 (MFEM convention: real order 2n-1)
//   SQUARE/CUBE : tensor products of the segment rule (x fastest ... z slowest is irrelevant to sums)
//   TRIANGLE/TETRAHEDRON : collapsed-coordinate (Duffy) Gauss–Jacobi product rules, exact for total
//     degree `order`.  MFEM ships tabulated simplex rules instead; both are exact for the polynomial
//     integrands of affine elements, so the element matrices agree to round-off.
#include "mfem.hpp"

namespace mfem
{
namespace
{
// Jacobi polynomial P_n^{(a,b)}(x) by the three-term recurrence.
double JacobiP(int n, double a, double b, double x)
{
    if (n == 0) return 1.0;
    double p0 = 1.0;
    double p1 = 0.5 * ((a - b) + (a + b + 2.0) * x);
    for (int k = 2; k <= n; ++k)
    {
        const double c = 2.0 * k + a + b;
        const double a1 = 2.0 * k * (k + a + b) * (c - 2.0);
        const double a2 = (c - 1.0) * (a * a - b * b);
        const double a3 = (c - 2.0) * (c - 1.0) * c;
        const double a4 = 2.0 * (k + a - 1.0) * (k + b - 1.0) * c;
        const double p2 = ((a2 + a3 * x) * p1 - a4 * p0) / a1;
        p0 = p1;
        p1 = p2;
    }
    return p1;
}

double JacobiDP(int n, double a, double b, double x)
{
    if (n == 0) return 0.0;
    return 0.5 * (n + a + b + 1.0) * JacobiP(n - 1, a + 1.0, b + 1.0, x);
}

// n-point Gauss–Jacobi rule for weight (1-t)^alpha on [-1,1] (beta = 0); nodes ascending.
void GaussJacobi(int n, int alpha, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    const double a = alpha, b = 0.0;
    // initial guesses: Chebyshev-like points, refined by Newton with deflation of found roots
    for (int i = 0; i < n; ++i)
    {
        double r = -std::cos((2.0 * i + 1.0) * M_PI / (2.0 * n));
        if (i > 0) r = 0.5 * (r + x[i - 1]);
        for (int it = 0; it < 200; ++it)
        {
            double s = 0.0;
            for (int j = 0; j < i; ++j) s += 1.0 / (r - x[j]);
            const double p = JacobiP(n, a, b, r);
            const double dp = JacobiDP(n, a, b, r);
            const double dr = -p / (dp - s * p);
            r += dr;
            if (std::fabs(dr) < 1e-16 * std::max(1.0, std::fabs(r))) break;
        }
        x[i] = r;
    }
    std::sort(x.begin(), x.end());
    // polish without deflation
    for (int i = 0; i < n; ++i)
    {
        for (int it = 0; it < 4; ++it) x[i] -= JacobiP(n, a, b, x[i]) / JacobiDP(n, a, b, x[i]);
    }
    const double c = std::pow(2.0, a + 1.0);
    for (int i = 0; i < n; ++i)
    {
        const double dp = JacobiDP(n, a, b, x[i]);
        w[i] = c / ((1.0 - x[i] * x[i]) * dp * dp);
    }
}
}  // namespace

void GaussLegendre01(int n, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    if (n == 1)
    {
        x[0] = 0.5;
        w[0] = 1.0;
        return;
    }
    const int m = (n + 1) / 2;
    for (int i = 1; i <= m; ++i)
    {
        // Newton on the Legendre polynomial P_n(z), z in (-1,1), starting from the classic guess
        double z = std::cos(M_PI * (i - 0.25) / (n + 0.5));
        double pp = 0.0;
        for (int it = 0; it < 100; ++it)
        {
            double p2 = 1.0, p1 = z;
            for (int j = 2; j <= n; ++j)
            {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
            }
            // p1 = P_n(z), p2 = P_{n-1}(z)
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        {
            double p2 = 1.0, p1 = z;
            for (int j = 2; j <= n; ++j)
            {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
        }
        const double xi = 0.5 * (1.0 - z);  // small root on [0,1]
        const double wi = 1.0 / ((1.0 - z * z) * pp * pp);  // = (2/((1-z^2) pp^2)) / 2
        x[i - 1] = xi;
        x[n - i] = 1.0 - xi;
        w[i - 1] = wi;
        w[n - i] = wi;
    }
}

IntegrationRules::IntegrationRules(int, int type)
{
    MFEM_VERIFY(type == Quadrature1D::GaussLegendre, "mfem_lite: only Gauss-Legendre rules are available");
}

const IntegrationRule &IntegrationRules::Get(int geom, int order)
{
    if (order < 0) order = 0;
    const int n = order / 2 + 1;
    const auto key = std::make_pair(geom, n);
    auto it = cache_.find(key);
    if (it != cache_.end()) return *it->second;

    std::vector<double> gx, gw;
    GaussLegendre01(n, gx, gw);
    std::unique_ptr<IntegrationRule> ir;
    switch (geom)
    {
        case Geometry::POINT:
        {
            ir.reset(new IntegrationRule(1));
            ir->IntPoint(0).weight = 1.0;
            break;
        }
        case Geometry::SEGMENT:
        {
            ir.reset(new IntegrationRule(n));
            for (int i = 0; i < n; ++i)
            {
                ir->IntPoint(i).x = gx[i];
                ir->IntPoint(i).weight = gw[i];
            }
            break;
        }
        case Geometry::SQUARE:
        {
            ir.reset(new IntegrationRule(n * n));
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i)
                {
                    IntegrationPoint &ip = ir->IntPoint(j * n + i);
                    ip.x = gx[i];
                    ip.y = gx[j];
                    ip.weight = gw[i] * gw[j];
                }
            break;
        }
        case Geometry::CUBE:
        {
            ir.reset(new IntegrationRule(n * n * n));
            for (int k = 0; k < n; ++k)
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < n; ++i)
                    {
                        IntegrationPoint &ip = ir->IntPoint((k * n + j) * n + i);
                        ip.x = gx[i];
                        ip.y = gx[j];
                        ip.z = gx[k];
                        ip.weight = gw[i] * gw[j] * gw[k];
                    }
            break;
        }
        case Geometry::TRIANGLE:
        {
            // x = u, y = v (1-u); dx dy = (1-u) du dv
            std::vector<double> jx, jw;
            GaussJacobi(n, 1, jx, jw);
            ir.reset(new IntegrationRule(n * n));
            for (int i = 0; i < n; ++i)
            {
                const double u = 0.5 * (1.0 + jx[i]);
                const double wu = 0.25 * jw[i];
                for (int j = 0; j < n; ++j)
                {
                    IntegrationPoint &ip = ir->IntPoint(i * n + j);
                    ip.x = u;
                    ip.y = gx[j] * (1.0 - u);
                    ip.weight = wu * gw[j];
                }
            }
            break;
        }
        case Geometry::TETRAHEDRON:
        {
            // x = u, y = v (1-u), z = w (1-u)(1-v); dV = (1-u)^2 (1-v) du dv dw
            std::vector<double> ax, aw, bx, bw;
            GaussJacobi(n, 2, ax, aw);
            GaussJacobi(n, 1, bx, bw);
            ir.reset(new IntegrationRule(n * n * n));
            for (int i = 0; i < n; ++i)
            {
                const double u = 0.5 * (1.0 + ax[i]);
                const double wu = 0.125 * aw[i];
                for (int j = 0; j < n; ++j)
                {
                    const double v = 0.5 * (1.0 + bx[j]);
                    const double wv = 0.25 * bw[j];
                    for (int k = 0; k < n; ++k)
                    {
                        IntegrationPoint &ip = ir->IntPoint((i * n + j) * n + k);
                        ip.x = u;
                        ip.y = v * (1.0 - u);
                        ip.z = gx[k] * (1.0 - u) * (1.0 - v);
                        ip.weight = wu * wv * gw[k];
                    }
                }
            }
            break;
        }
        default: mfem_error("IntegrationRules::Get: unsupported geometry");
    }
    ir->SetOrder(2 * n - 1);
    for (int i = 0; i < ir->GetNPoints(); ++i) ir->IntPoint(i).index = i;
    const IntegrationRule &ref = *ir;
    cache_[key] = std::move(ir);
    return ref;
}

}  // namespace mfem
