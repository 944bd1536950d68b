// mfem_lite: geometry tables, isoparametric transformations, L2 Gauss–Legendre nodal elements,
// finite-element space and grid function.
//
// Conventions [log-confirmed = pinned by the reference's shipped output/log/integrals_all.txt]:
//   * L2_FECollection(p, dim) uses the GaussLegendre ("open") nodal basis          [log-confirmed]
//   * triangle nodes (o_i/w, o_j/w), w = o_i + o_j + o_{p-i-j}, j outer / i inner   [log-confirmed]
//   * tetrahedron analogue with k outermost; quad/hex tensor nodes with i fastest   [from memory]
#include "mfem.hpp"

namespace mfem
{
const char *Geometry::Name[NUM_GEOMETRIES] = {"Point", "Segment", "Triangle", "Square", "Tetrahedron", "Cube", "Prism", "Pyramid"};
const int Geometry::NumVerts[NUM_GEOMETRIES] = {1, 2, 3, 4, 4, 8, 6, 5};
const int Geometry::Dimension[NUM_GEOMETRIES] = {0, 1, 2, 2, 3, 3, 3, 3};

// ------------------------------------------------------------------ ElementTransformation
void ElementTransformation::Setup(Geometry::Type geom, int sdim, const double *const *verts)
{
    geom_ = geom;
    sdim_ = sdim;
    const int nv = Geometry::NumVerts[geom];
    pm_.SetSize(sdim, nv);
    for (int j = 0; j < nv; ++j)
        for (int d = 0; d < sdim; ++d) pm_(d, j) = verts[j][d];
    eval_ = 0;
    ip_ = nullptr;
}

void ElementTransformation::GeomShape(const IntegrationPoint &ip, Vector &s) const
{
    const double x = ip.x, y = ip.y, z = ip.z;
    s.SetSize(Geometry::NumVerts[geom_]);
    switch (geom_)
    {
        case Geometry::SEGMENT:
            s[0] = 1.0 - x;
            s[1] = x;
            break;
        case Geometry::TRIANGLE:
            s[0] = 1.0 - x - y;
            s[1] = x;
            s[2] = y;
            break;
        case Geometry::SQUARE:
            s[0] = (1.0 - x) * (1.0 - y);
            s[1] = x * (1.0 - y);
            s[2] = x * y;
            s[3] = (1.0 - x) * y;
            break;
        case Geometry::TETRAHEDRON:
            s[0] = 1.0 - x - y - z;
            s[1] = x;
            s[2] = y;
            s[3] = z;
            break;
        case Geometry::CUBE:
        {
            const double ox = 1.0 - x, oy = 1.0 - y, oz = 1.0 - z;
            s[0] = ox * oy * oz;
            s[1] = x * oy * oz;
            s[2] = x * y * oz;
            s[3] = ox * y * oz;
            s[4] = ox * oy * z;
            s[5] = x * oy * z;
            s[6] = x * y * z;
            s[7] = ox * y * z;
            break;
        }
        default: mfem_error("ElementTransformation: unsupported geometry");
    }
}

void ElementTransformation::GeomDShape(const IntegrationPoint &ip, DenseMatrix &ds) const
{
    const double x = ip.x, y = ip.y, z = ip.z;
    ds.SetSize(Geometry::NumVerts[geom_], Geometry::Dimension[geom_]);
    switch (geom_)
    {
        case Geometry::SEGMENT:
            ds(0, 0) = -1.0;
            ds(1, 0) = 1.0;
            break;
        case Geometry::TRIANGLE:
            ds(0, 0) = -1.0; ds(0, 1) = -1.0;
            ds(1, 0) = 1.0;  ds(1, 1) = 0.0;
            ds(2, 0) = 0.0;  ds(2, 1) = 1.0;
            break;
        case Geometry::SQUARE:
            ds(0, 0) = -1.0 + y; ds(0, 1) = -1.0 + x;
            ds(1, 0) = 1.0 - y;  ds(1, 1) = -x;
            ds(2, 0) = y;        ds(2, 1) = x;
            ds(3, 0) = -y;       ds(3, 1) = 1.0 - x;
            break;
        case Geometry::TETRAHEDRON:
            ds(0, 0) = -1.0; ds(0, 1) = -1.0; ds(0, 2) = -1.0;
            ds(1, 0) = 1.0;  ds(1, 1) = 0.0;  ds(1, 2) = 0.0;
            ds(2, 0) = 0.0;  ds(2, 1) = 1.0;  ds(2, 2) = 0.0;
            ds(3, 0) = 0.0;  ds(3, 1) = 0.0;  ds(3, 2) = 1.0;
            break;
        case Geometry::CUBE:
        {
            const double ox = 1.0 - x, oy = 1.0 - y, oz = 1.0 - z;
            ds(0, 0) = -oy * oz; ds(0, 1) = -ox * oz; ds(0, 2) = -ox * oy;
            ds(1, 0) = oy * oz;  ds(1, 1) = -x * oz;  ds(1, 2) = -x * oy;
            ds(2, 0) = y * oz;   ds(2, 1) = x * oz;   ds(2, 2) = -x * y;
            ds(3, 0) = -y * oz;  ds(3, 1) = ox * oz;  ds(3, 2) = -ox * y;
            ds(4, 0) = -oy * z;  ds(4, 1) = -ox * z;  ds(4, 2) = ox * oy;
            ds(5, 0) = oy * z;   ds(5, 1) = -x * z;   ds(5, 2) = x * oy;
            ds(6, 0) = y * z;    ds(6, 1) = x * z;    ds(6, 2) = x * y;
            ds(7, 0) = -y * z;   ds(7, 1) = ox * z;   ds(7, 2) = ox * y;
            break;
        }
        default: mfem_error("ElementTransformation: unsupported geometry");
    }
}

const DenseMatrix &ElementTransformation::Jacobian()
{
    if (!(eval_ & 1))
    {
        MFEM_VERIFY(ip_ != nullptr, "ElementTransformation: integration point not set");
        GeomDShape(*ip_, ds_);
        // J = PointMat * dshape  (sdim x dim), summed over the geometric nodes in order
        const int nv = pm_.Width(), dim = ds_.Width();
        J_.SetSize(sdim_, dim);
        J_ = 0.0;
        for (int k = 0; k < nv; ++k)
            for (int j = 0; j < dim; ++j)
                for (int i = 0; i < sdim_; ++i) J_(i, j) += pm_(i, k) * ds_(k, j);
        eval_ |= 1;
    }
    return J_;
}

double ElementTransformation::Weight()
{
    if (!(eval_ & 2))
    {
        w_ = Jacobian().Weight();
        eval_ |= 2;
    }
    return w_;
}

const DenseMatrix &ElementTransformation::InverseJacobian()
{
    if (!(eval_ & 4))
    {
        CalcInverse(Jacobian(), Jinv_);
        eval_ |= 4;
    }
    return Jinv_;
}

void ElementTransformation::Transform(const IntegrationPoint &ip, Vector &x)
{
    Vector s;
    GeomShape(ip, s);
    x.SetSize(sdim_);
    pm_.Mult(s, x);
}

void FaceElementTransformations::SetAllIntPoints(const IntegrationPoint *face_ip)
{
    Face->SetIntPoint(face_ip);
    // reference coordinates of the vertices of each geometry
    static const double ref_seg[2][3] = {{0, 0, 0}, {1, 0, 0}};
    static const double ref_tri[3][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    static const double ref_sq[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
    static const double ref_tet[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static const double ref_cube[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    auto ref_of = [&](Geometry::Type g) -> const double(*)[3] {
        switch (g)
        {
            case Geometry::SEGMENT: return ref_seg;
            case Geometry::TRIANGLE: return ref_tri;
            case Geometry::SQUARE: return ref_sq;
            case Geometry::TETRAHEDRON: return ref_tet;
            case Geometry::CUBE: return ref_cube;
            default: mfem_error("SetAllIntPoints: unsupported geometry");
        }
    };
    // face shape functions at the face point
    const Geometry::Type fg = Face->GetGeometryType();
    const int nfv = Geometry::NumVerts[fg];
    double s[4];
    const double x = face_ip->x, y = face_ip->y;
    switch (fg)
    {
        case Geometry::SEGMENT: s[0] = 1.0 - x; s[1] = x; break;
        case Geometry::TRIANGLE: s[0] = 1.0 - x - y; s[1] = x; s[2] = y; break;
        case Geometry::SQUARE: s[0] = (1 - x) * (1 - y); s[1] = x * (1 - y); s[2] = x * y; s[3] = (1 - x) * y; break;
        default: mfem_error("SetAllIntPoints: unsupported face geometry");
    }
    auto map_to = [&](Geometry::Type eg, const int *ev, IntegrationPoint &out) {
        const double(*ref)[3] = ref_of(eg);
        const int nev = Geometry::NumVerts[eg];
        double p[3] = {0, 0, 0};
        for (int i = 0; i < nfv; ++i)
        {
            int l = -1;
            for (int k = 0; k < nev; ++k)
                if (ev[k] == fv[i]) l = k;
            MFEM_VERIFY(l >= 0, "face vertex not found in element");
            for (int d = 0; d < 3; ++d) p[d] += s[i] * ref[l][d];
        }
        out.x = p[0];
        out.y = p[1];
        out.z = p[2];
        out.weight = face_ip->weight;
    };
    if (Elem1No >= 0)
    {
        map_to(egeom1, ev1, eip1_);
        Elem1->SetIntPoint(&eip1_);
    }
    if (Elem2No >= 0)
    {
        map_to(egeom2, ev2, eip2_);
        Elem2->SetIntPoint(&eip2_);
    }
}

int InverseElementTransformation::Transform(const Vector &pt, IntegrationPoint &ip)
{
    MFEM_VERIFY(T_ != nullptr, "InverseElementTransformation: no transformation");
    const Geometry::Type g = T_->GetGeometryType();
    const int dim = T_->GetDimension();
    MFEM_VERIFY(dim == T_->GetSpaceDim(), "InverseElementTransformation: square Jacobian required");
    // start from the reference centre
    IntegrationPoint xi;
    switch (g)
    {
        case Geometry::SEGMENT: xi.x = 0.5; break;
        case Geometry::TRIANGLE: xi.x = xi.y = 1.0 / 3.0; break;
        case Geometry::SQUARE: xi.x = xi.y = 0.5; break;
        case Geometry::TETRAHEDRON: xi.x = xi.y = xi.z = 0.25; break;
        case Geometry::CUBE: xi.x = xi.y = xi.z = 0.5; break;
        default: mfem_error("InverseElementTransformation: unsupported geometry");
    }
    Vector x(dim), dx(dim), dxi(dim);
    for (int it = 0; it < std::max(max_iter_, 1); ++it)
    {
        T_->SetIntPoint(&xi);
        T_->Transform(xi, x);
        for (int d = 0; d < dim; ++d) dx[d] = pt[d] - x[d];
        T_->InverseJacobian().Mult(dx, dxi);
        xi.x += dxi[0];
        if (dim > 1) xi.y += dxi[1];
        if (dim > 2) xi.z += dxi[2];
        if (dxi.Norml2() <= ref_tol_) break;
    }
    ip = xi;
    const double t = ip_tol_ > ref_tol_ ? ip_tol_ : ref_tol_;
    bool inside = true;
    switch (g)
    {
        case Geometry::SEGMENT: inside = ip.x >= -t && ip.x <= 1.0 + t; break;
        case Geometry::TRIANGLE: inside = ip.x >= -t && ip.y >= -t && ip.x + ip.y <= 1.0 + t; break;
        case Geometry::SQUARE: inside = ip.x >= -t && ip.x <= 1.0 + t && ip.y >= -t && ip.y <= 1.0 + t; break;
        case Geometry::TETRAHEDRON: inside = ip.x >= -t && ip.y >= -t && ip.z >= -t && ip.x + ip.y + ip.z <= 1.0 + t; break;
        case Geometry::CUBE:
            inside = ip.x >= -t && ip.x <= 1.0 + t && ip.y >= -t && ip.y <= 1.0 + t && ip.z >= -t && ip.z <= 1.0 + t;
            break;
        default: break;
    }
    return inside ? Inside : Outside;
}

// ------------------------------------------------------------------ 1-D Lagrange basis helpers
namespace
{
struct Lagrange1D
{
    std::vector<double> x;  // nodes
    explicit Lagrange1D(int p)
    {
        std::vector<double> w;
        GaussLegendre01(p + 1, x, w);
    }
    void Eval(double y, double *u) const
    {
        const int n = static_cast<int>(x.size());
        for (int i = 0; i < n; ++i)
        {
            double v = 1.0;
            for (int j = 0; j < n; ++j)
                if (j != i) v *= (y - x[j]) / (x[i] - x[j]);
            u[i] = v;
        }
    }
    void Eval(double y, double *u, double *d) const
    {
        Eval(y, u);
        const int n = static_cast<int>(x.size());
        for (int i = 0; i < n; ++i)
        {
            double s = 0.0;
            for (int k = 0; k < n; ++k)
            {
                if (k == i) continue;
                double v = 1.0 / (x[i] - x[k]);
                for (int j = 0; j < n; ++j)
                    if (j != i && j != k) v *= (y - x[j]) / (x[i] - x[j]);
                s += v;
            }
            d[i] = s;
        }
    }
};

// Chebyshev polynomials T_n(2x-1), n = 0..p, and their x-derivatives
void Cheb(int p, double x, double *u, double *d)
{
    const double z = 2.0 * x - 1.0;
    u[0] = 1.0;
    d[0] = 0.0;
    if (p == 0) return;
    u[1] = z;
    d[1] = 2.0;
    for (int n = 1; n < p; ++n)
    {
        u[n + 1] = 2.0 * z * u[n] - u[n - 1];
        d[n + 1] = 4.0 * u[n] + 2.0 * z * d[n] - d[n - 1];
    }
}

class L2_TensorElement : public FiniteElement
{
public:
    L2_TensorElement(int dim, int p) : b_(p)
    {
        dim_ = dim;
        order_ = p;
        geom_ = dim == 1 ? Geometry::SEGMENT : (dim == 2 ? Geometry::SQUARE : Geometry::CUBE);
        n1_ = p + 1;
        dof_ = 1;
        for (int d = 0; d < dim; ++d) dof_ *= n1_;
        nodes_.SetSize(dof_, dim);
        for (int o = 0; o < dof_; ++o)
        {
            int r = o;
            for (int d = 0; d < dim; ++d)
            {
                nodes_(o, d) = b_.x[r % n1_];
                r /= n1_;
            }
        }
    }
    void CalcShape(const IntegrationPoint &ip, Vector &shape) const override
    {
        double ux[16], uy[16], uz[16];
        b_.Eval(ip.x, ux);
        if (dim_ > 1) b_.Eval(ip.y, uy);
        if (dim_ > 2) b_.Eval(ip.z, uz);
        shape.SetSize(dof_);
        int o = 0;
        if (dim_ == 1)
            for (int i = 0; i < n1_; ++i) shape[o++] = ux[i];
        else if (dim_ == 2)
            for (int j = 0; j < n1_; ++j)
                for (int i = 0; i < n1_; ++i) shape[o++] = ux[i] * uy[j];
        else
            for (int k = 0; k < n1_; ++k)
                for (int j = 0; j < n1_; ++j)
                    for (int i = 0; i < n1_; ++i) shape[o++] = ux[i] * uy[j] * uz[k];
    }
    void CalcDShape(const IntegrationPoint &ip, DenseMatrix &ds) const override
    {
        double ux[16], uy[16], uz[16], dx[16], dy[16], dz[16];
        b_.Eval(ip.x, ux, dx);
        if (dim_ > 1) b_.Eval(ip.y, uy, dy);
        if (dim_ > 2) b_.Eval(ip.z, uz, dz);
        ds.SetSize(dof_, dim_);
        int o = 0;
        if (dim_ == 1)
            for (int i = 0; i < n1_; ++i) ds(o++, 0) = dx[i];
        else if (dim_ == 2)
            for (int j = 0; j < n1_; ++j)
                for (int i = 0; i < n1_; ++i, ++o)
                {
                    ds(o, 0) = dx[i] * uy[j];
                    ds(o, 1) = ux[i] * dy[j];
                }
        else
            for (int k = 0; k < n1_; ++k)
                for (int j = 0; j < n1_; ++j)
                    for (int i = 0; i < n1_; ++i, ++o)
                    {
                        ds(o, 0) = dx[i] * uy[j] * uz[k];
                        ds(o, 1) = ux[i] * dy[j] * uz[k];
                        ds(o, 2) = ux[i] * uy[j] * dz[k];
                    }
    }

private:
    Lagrange1D b_;
    int n1_;
};

class L2_SimplexElement : public FiniteElement
{
public:
    L2_SimplexElement(int dim, int p)
    {
        MFEM_VERIFY(p < 16, "mfem_lite: simplex order too high");
        dim_ = dim;
        order_ = p;
        geom_ = dim == 2 ? Geometry::TRIANGLE : Geometry::TETRAHEDRON;
        dof_ = dim == 2 ? (p + 1) * (p + 2) / 2 : (p + 1) * (p + 2) * (p + 3) / 6;
        std::vector<double> op, w;
        GaussLegendre01(p + 1, op, w);
        nodes_.SetSize(dof_, dim);
        int o = 0;
        if (dim == 2)
        {
            for (int j = 0; j <= p; ++j)
                for (int i = 0; i + j <= p; ++i, ++o)
                {
                    const double ww = op[i] + op[j] + op[p - i - j];
                    nodes_(o, 0) = op[i] / ww;
                    nodes_(o, 1) = op[j] / ww;
                }
        }
        else
        {
            for (int k = 0; k <= p; ++k)
                for (int j = 0; j + k <= p; ++j)
                    for (int i = 0; i + j + k <= p; ++i, ++o)
                    {
                        const double ww = op[i] + op[j] + op[k] + op[p - i - j - k];
                        nodes_(o, 0) = op[i] / ww;
                        nodes_(o, 1) = op[j] / ww;
                        nodes_(o, 2) = op[k] / ww;
                    }
        }
        // T(o, k) = u_o(node_k); nodal shape = T^{-1} u
        DenseMatrix T(dof_, dof_);
        Vector u(dof_);
        DenseMatrix du;
        for (int k = 0; k < dof_; ++k)
        {
            IntegrationPoint ip;
            ip.x = nodes_(k, 0);
            ip.y = nodes_(k, 1);
            ip.z = dim == 3 ? nodes_(k, 2) : 0.0;
            Modal(ip, u, nullptr);
            for (int oo = 0; oo < dof_; ++oo) T(oo, k) = u[oo];
        }
        T_ = T;
        Ti_.Factor(T_);
    }
    void CalcShape(const IntegrationPoint &ip, Vector &shape) const override
    {
        Vector u(dof_);
        Modal(ip, u, nullptr);
        shape.SetSize(dof_);
        Ti_.Mult(u, shape);
    }
    void CalcDShape(const IntegrationPoint &ip, DenseMatrix &ds) const override
    {
        Vector u(dof_), col(dof_), out(dof_);
        DenseMatrix du(dof_, dim_);
        Modal(ip, u, &du);
        ds.SetSize(dof_, dim_);
        for (int d = 0; d < dim_; ++d)
        {
            du.GetColumn(d, col);
            Ti_.Mult(col, out);
            ds.SetCol(d, out);
        }
    }

private:
    void Modal(const IntegrationPoint &ip, Vector &u, DenseMatrix *du) const
    {
        const int p = order_;
        double sx[16], sy[16], sz[16], sl[16], dx[16], dy[16], dz[16], dl[16];
        Cheb(p, ip.x, sx, dx);
        Cheb(p, ip.y, sy, dy);
        int o = 0;
        if (dim_ == 2)
        {
            Cheb(p, 1.0 - ip.x - ip.y, sl, dl);
            for (int j = 0; j <= p; ++j)
                for (int i = 0; i + j <= p; ++i, ++o)
                {
                    const int k = p - i - j;
                    u[o] = sx[i] * sy[j] * sl[k];
                    if (du)
                    {
                        (*du)(o, 0) = (dx[i] * sl[k] - sx[i] * dl[k]) * sy[j];
                        (*du)(o, 1) = (dy[j] * sl[k] - sy[j] * dl[k]) * sx[i];
                    }
                }
        }
        else
        {
            Cheb(p, ip.z, sz, dz);
            Cheb(p, 1.0 - ip.x - ip.y - ip.z, sl, dl);
            for (int k = 0; k <= p; ++k)
                for (int j = 0; j + k <= p; ++j)
                    for (int i = 0; i + j + k <= p; ++i, ++o)
                    {
                        const int l = p - i - j - k;
                        u[o] = sx[i] * sy[j] * sz[k] * sl[l];
                        if (du)
                        {
                            (*du)(o, 0) = (dx[i] * sl[l] - sx[i] * dl[l]) * sy[j] * sz[k];
                            (*du)(o, 1) = (dy[j] * sl[l] - sy[j] * dl[l]) * sx[i] * sz[k];
                            (*du)(o, 2) = (dz[k] * sl[l] - sz[k] * dl[l]) * sx[i] * sy[j];
                        }
                    }
        }
    }
    DenseMatrix T_;
    DenseMatrixInverse Ti_;
};
}  // namespace

void FiniteElement::CalcPhysDShape(ElementTransformation &Trans, DenseMatrix &dshape) const
{
    DenseMatrix ref;
    CalcDShape(Trans.GetIntPoint(), ref);
    Mult(ref, Trans.InverseJacobian(), dshape);
}

L2_FECollection::L2_FECollection(int p, int dim, int btype) : dim_(dim)
{
    MFEM_VERIFY(btype == BasisType::GaussLegendre, "mfem_lite: only the Gauss-Legendre L2 basis is available");
    MFEM_VERIFY(p >= 0 && p < 15, "mfem_lite: unsupported order");
    order_ = p;
    std::ostringstream os;
    os << "L2_" << dim << "D_P" << p;
    name_ = os.str();
}

const FiniteElement *L2_FECollection::FiniteElementForGeometry(Geometry::Type g) const
{
    auto it = fes_.find(g);
    if (it != fes_.end()) return it->second.get();
    FiniteElement *fe = nullptr;
    switch (g)
    {
        case Geometry::SEGMENT: fe = new L2_TensorElement(1, order_); break;
        case Geometry::SQUARE: fe = new L2_TensorElement(2, order_); break;
        case Geometry::CUBE: fe = new L2_TensorElement(3, order_); break;
        case Geometry::TRIANGLE: fe = new L2_SimplexElement(2, order_); break;
        case Geometry::TETRAHEDRON: fe = new L2_SimplexElement(3, order_); break;
        default: mfem_error("L2_FECollection: unsupported geometry");
    }
    fes_[g].reset(fe);
    return fe;
}

// ------------------------------------------------------------------ FiniteElementSpace
}  // namespace mfem

#include "mfem.hpp"
namespace mfem
{
FiniteElementSpace::FiniteElementSpace(Mesh *mesh, const FiniteElementCollection *fec, int vdim, int ordering)
    : mesh_(mesh), fec_(fec), vdim_(vdim), ordering_(static_cast<Ordering::Type>(ordering))
{
    const int ne = mesh_->GetNE();
    offsets_.resize(ne + 1);
    offsets_[0] = 0;
    for (int e = 0; e < ne; ++e) offsets_[e + 1] = offsets_[e] + GetFE(e)->GetDof();
    ndofs_ = offsets_[ne];
}

const FiniteElement *FiniteElementSpace::GetFE(int i) const
{
    return fec_->FiniteElementForGeometry(mesh_->GetElementBaseGeometry(i));
}

int FiniteElementSpace::GetNE() const { return mesh_->GetNE(); }

void FiniteElementSpace::GetElementDofs(int e, Array<int> &dofs) const
{
    const int n = offsets_[e + 1] - offsets_[e];
    dofs.SetSize(n);
    for (int j = 0; j < n; ++j) dofs[j] = offsets_[e] + j;
}

void FiniteElementSpace::GetElementVDofs(int e, Array<int> &vdofs) const
{
    const int n = offsets_[e + 1] - offsets_[e];
    vdofs.SetSize(n * vdim_);
    for (int vd = 0; vd < vdim_; ++vd)
        for (int j = 0; j < n; ++j)
        {
            const int dof = offsets_[e] + j;
            vdofs[vd * n + j] = ordering_ == Ordering::byNODES ? dof + vd * ndofs_ : dof * vdim_ + vd;
        }
}

double GridFunction::GetValue(int e, const IntegrationPoint &ip, int vdim) const
{
    const FiniteElement *fe = fes_->GetFE(e);
    Vector shape(fe->GetDof());
    fe->CalcShape(ip, shape);
    Array<int> vdofs;
    fes_->GetElementVDofs(e, vdofs);
    const int n = fe->GetDof();
    double v = 0.0;
    for (int j = 0; j < n; ++j) v += (*this)(vdofs[(vdim - 1) * n + j]) * shape[j];
    return v;
}

}  // namespace mfem
