// mfem_lite — a minimal, self-written stand-in for the subset of the MFEM 4.8 API that the PBTE
// host layer touches (mesh, L2 finite-element space, quadrature, small dense linear algebra).
//
// Why it exists: the deployment target links real MFEM ("MFEM used only for mesh and FE-space
// setup"), but MFEM is not installed in the build container and cannot be fetched.  This header
// keeps the host classes and the reference's own sources compiling here.  It is NOT a copy of
// MFEM: it re-implements documented behaviour (names, argument meaning, numbering conventions)
// from the published API; conventions that the reference's shipped logs confirm are marked
// [log-confirmed] in the .cpp files, the rest [from memory of MFEM 4.x].
//
// Include order compatibility: `#include "mfem.hpp"` is all a client needs.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <fstream>
#include <functional>
#include <limits>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mfem
{
typedef double real_t;

// ----------------------------------------------------------------------------------------------
// error handling
// ----------------------------------------------------------------------------------------------
class ErrorException : public std::runtime_error
{
public:
    explicit ErrorException(const std::string &m) : std::runtime_error(m) {}
};

enum ErrorAction { MFEM_ERROR_ABORT = 0, MFEM_ERROR_THROW };
void set_error_action(ErrorAction action);
ErrorAction get_error_action();
[[noreturn]] void mfem_error(const char *msg = nullptr);
inline void mfem_error(const std::string &msg) { mfem_error(msg.c_str()); }

#define MFEM_VERIFY(cond, msg)                                                        \
    do {                                                                              \
        if (!(cond)) {                                                                \
            std::ostringstream mfem_lite_oss_;                                        \
            mfem_lite_oss_ << "Verification failed: (" #cond ") is false:\n --> " << msg \
                           << "\n ... in file: " << __FILE__ << ':' << __LINE__;      \
            ::mfem::mfem_error(mfem_lite_oss_.str().c_str());                         \
        }                                                                             \
    } while (0)
#define MFEM_ABORT(msg)                                   \
    do {                                                  \
        std::ostringstream mfem_lite_oss_;                \
        mfem_lite_oss_ << "MFEM abort: " << msg;          \
        ::mfem::mfem_error(mfem_lite_oss_.str().c_str()); \
    } while (0)
#define MFEM_ASSERT(cond, msg) ((void)0)

template <class T>
inline void Swap(T &a, T &b)
{
    T c = a;
    a = b;
    b = c;
}

// ----------------------------------------------------------------------------------------------
// Array<T>
// ----------------------------------------------------------------------------------------------
template <class T>
class Array
{
public:
    Array() = default;
    explicit Array(int n) : d_(n) {}
    int Size() const { return static_cast<int>(d_.size()); }
    void SetSize(int n) { d_.resize(n); }
    void SetSize(int n, const T &v) { d_.resize(n, v); }
    T &operator[](int i) { return d_[i]; }
    const T &operator[](int i) const { return d_[i]; }
    int Append(const T &v)
    {
        d_.push_back(v);
        return Size();
    }
    T *GetData() { return d_.data(); }
    const T *GetData() const { return d_.data(); }
    void operator=(const T &v) { std::fill(d_.begin(), d_.end(), v); }
    T *begin() { return d_.data(); }
    T *end() { return d_.data() + d_.size(); }
    const T *begin() const { return d_.data(); }
    const T *end() const { return d_.data() + d_.size(); }
    void DeleteAll() { d_.clear(); }
    T &Last() { return d_.back(); }

private:
    std::vector<T> d_;
};

// ----------------------------------------------------------------------------------------------
// Vector
// ----------------------------------------------------------------------------------------------
class Vector
{
public:
    Vector() = default;
    explicit Vector(int n) { SetSize(n); }
    /// Non-owning view on external storage (as in MFEM).
    Vector(double *data, int n) : data_(data), size_(n), cap_(0) {}
    Vector(const Vector &o);
    Vector(Vector &&o) noexcept;
    ~Vector();

    Vector &operator=(const Vector &o);
    Vector &operator=(Vector &&o) noexcept;
    Vector &operator=(double v);

    void SetSize(int n);
    int Size() const { return size_; }
    double *GetData() { return data_; }
    const double *GetData() const { return data_; }

    double &operator[](int i) { return data_[i]; }
    const double &operator[](int i) const { return data_[i]; }
    double &operator()(int i) { return data_[i]; }
    const double &operator()(int i) const { return data_[i]; }

    /// Dot product, accumulated sequentially i = 0..n-1.
    double operator*(const Vector &v) const;
    double operator*(const double *v) const;
    Vector &operator*=(double c);
    Vector &operator/=(double c);
    Vector &operator-=(const Vector &v);
    Vector &operator+=(const Vector &v);
    /// this += a * x
    Vector &Add(double a, const Vector &x);
    double Norml2() const;
    double Max() const;
    double Min() const;
    void Swap(Vector &o);
    void SetSubVector(const Array<int> &dofs, const Vector &elemvect);
    void GetSubVector(const Array<int> &dofs, Vector &elemvect) const;

    double *begin() { return data_; }
    double *end() { return data_ + size_; }
    const double *begin() const { return data_; }
    const double *end() const { return data_ + size_; }

private:
    double *data_ = nullptr;
    int size_ = 0;
    int cap_ = 0;  // >0: owning with this capacity; 0: empty or non-owning view
};

// ----------------------------------------------------------------------------------------------
// DenseMatrix (column-major: entry (i,j) at data[i + j*height])
// ----------------------------------------------------------------------------------------------
class DenseMatrix
{
public:
    DenseMatrix() = default;
    explicit DenseMatrix(int s) { SetSize(s, s); }
    DenseMatrix(int h, int w) { SetSize(h, w); }
    /// View of caller-owned storage (never freed here), as in MFEM.
    DenseMatrix(double *d, int h, int w) { UseExternalData(d, h, w); }
    DenseMatrix(const DenseMatrix &o);  // deep copy into owned storage
    DenseMatrix(DenseMatrix &&o) noexcept;
    DenseMatrix &operator=(const DenseMatrix &o);
    DenseMatrix &operator=(DenseMatrix &&o) noexcept;
    DenseMatrix &operator=(double v);

    void SetSize(int s) { SetSize(s, s); }
    void SetSize(int h, int w);
    void UseExternalData(double *d, int h, int w);
    bool OwnsData() const { return ext_ == nullptr; }
    int Height() const { return h_; }
    int Width() const { return w_; }
    int NumRows() const { return h_; }
    int NumCols() const { return w_; }
    double *Data() { return ext_ ? ext_ : d_.data(); }
    const double *Data() const { return ext_ ? ext_ : d_.data(); }
    double *GetData() { return Data(); }
    const double *GetData() const { return Data(); }

    double &operator()(int i, int j) { return Data()[i + static_cast<size_t>(j) * h_]; }
    const double &operator()(int i, int j) const { return Data()[i + static_cast<size_t>(j) * h_]; }

    DenseMatrix &operator*=(double c);
    DenseMatrix &operator+=(const DenseMatrix &A);
    /// this += c * A
    void Add(double c, const DenseMatrix &A);
    /// y = this * x, accumulated column by column.
    void Mult(const Vector &x, Vector &y) const;
    void Mult(const double *x, double *y) const;
    void MultTranspose(const Vector &x, Vector &y) const;
    void GetColumn(int c, Vector &col) const;
    void SetCol(int c, const Vector &col);
    void Transpose();
    void Swap(DenseMatrix &o);
    double FNorm() const;
    double Det() const;
    /// sqrt(det(J^T J)) for non-square Jacobians, det for square ones.
    double Weight() const;

private:
    size_t Count() const { return static_cast<size_t>(h_) * w_; }
    std::vector<double> d_;
    double *ext_ = nullptr;  // non-null: view of external storage
    int h_ = 0, w_ = 0;
};

/// C = A * B
void Mult(const DenseMatrix &A, const DenseMatrix &B, DenseMatrix &C);
/// inverse (square, <= 3) or left pseudo-inverse of a 3x2 / 2x1 / 3x1 Jacobian
void CalcInverse(const DenseMatrix &A, DenseMatrix &Ainv);

/// Row-pivoted in-place LU (first maximal pivot), MFEM LUFactors semantics.
class DenseMatrixInverse
{
public:
    DenseMatrixInverse() = default;
    explicit DenseMatrixInverse(const DenseMatrix &mat);
    explicit DenseMatrixInverse(const DenseMatrix *mat);
    void Factor();
    void Factor(const DenseMatrix &mat);
    void Mult(const Vector &x, Vector &y) const;
    void Mult(const double *x, double *y) const;
    int Size() const { return n_; }
    const std::vector<double> &LUData() const { return lu_; }
    const std::vector<int> &Pivots() const { return ipiv_; }

private:
    const DenseMatrix *a_ = nullptr;
    std::vector<double> lu_;
    std::vector<int> ipiv_;
    int n_ = 0;
};

// ----------------------------------------------------------------------------------------------
// Geometry / quadrature
// ----------------------------------------------------------------------------------------------
class Geometry
{
public:
    enum Type { INVALID = -1, POINT = 0, SEGMENT, TRIANGLE, SQUARE, TETRAHEDRON, CUBE, PRISM, PYRAMID, NUM_GEOMETRIES };
    static const char *Name[NUM_GEOMETRIES];
    static const int NumVerts[NUM_GEOMETRIES];
    static const int Dimension[NUM_GEOMETRIES];
};

class Element
{
public:
    enum Type { POINT, SEGMENT, TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON, WEDGE, PYRAMID };
};

class IntegrationPoint
{
public:
    real_t x = 0.0, y = 0.0, z = 0.0, weight = 0.0;
    int index = 0;
    void Init(int i)
    {
        x = y = z = weight = 0.0;
        index = i;
    }
    void Set(const real_t *p, int dim)
    {
        x = p[0];
        y = dim > 1 ? p[1] : 0.0;
        z = dim > 2 ? p[2] : 0.0;
    }
    void Set2(real_t x1, real_t x2) { x = x1, y = x2; }
    void Set3(real_t x1, real_t x2, real_t x3) { x = x1, y = x2, z = x3; }
    void Get(real_t *p, int dim) const
    {
        p[0] = x;
        if (dim > 1) p[1] = y;
        if (dim > 2) p[2] = z;
    }
};

class IntegrationRule
{
public:
    IntegrationRule() = default;
    explicit IntegrationRule(int n) : pts_(n)
    {
        for (int i = 0; i < n; ++i) pts_[i].Init(i);
    }
    int GetNPoints() const { return static_cast<int>(pts_.size()); }
    int Size() const { return GetNPoints(); }
    IntegrationPoint &IntPoint(int i) { return pts_[i]; }
    const IntegrationPoint &IntPoint(int i) const { return pts_[i]; }
    int GetOrder() const { return order_; }
    void SetOrder(int o) { order_ = o; }

private:
    std::vector<IntegrationPoint> pts_;
    int order_ = 0;
};

class Quadrature1D
{
public:
    enum { Invalid = -1, GaussLegendre = 0, GaussLobatto = 1, OpenUniform = 2, ClosedUniform = 3, OpenHalfUniform = 4, ClosedGL = 5 };
};

/// Gauss–Legendre points on [0,1], ascending; weights sum to 1.
void GaussLegendre01(int n, std::vector<double> &x, std::vector<double> &w);

class IntegrationRules
{
public:
    explicit IntegrationRules(int ref = 0, int type = Quadrature1D::GaussLegendre);
    /// Rule exact for polynomials of total degree `order` (simplices) or per-direction degree
    /// `order` (tensor geometries).
    const IntegrationRule &Get(int geom_type, int order);

private:
    std::map<std::pair<int, int>, std::unique_ptr<IntegrationRule>> cache_;
};

// ----------------------------------------------------------------------------------------------
// transformations
// ----------------------------------------------------------------------------------------------
class Mesh;

/// Isoparametric map of one (multi)linear geometric cell; also used for faces (sdim > dim).
class ElementTransformation
{
public:
    int Attribute = 0, ElementNo = -1;

    void Setup(Geometry::Type geom, int sdim, const double *const *verts);
    Geometry::Type GetGeometryType() const { return geom_; }
    int GetDimension() const { return Geometry::Dimension[geom_]; }
    int GetSpaceDim() const { return sdim_; }
    void SetIntPoint(const IntegrationPoint *ip)
    {
        ip_ = ip;
        eval_ = 0;
    }
    const IntegrationPoint &GetIntPoint() const { return *ip_; }
    const DenseMatrix &Jacobian();
    const DenseMatrix &InverseJacobian();
    double Weight();
    void Transform(const IntegrationPoint &ip, Vector &x);
    const DenseMatrix &GetPointMat() const { return pm_; }
    int Order() const { return 1; }

private:
    void GeomDShape(const IntegrationPoint &ip, DenseMatrix &ds) const;
    void GeomShape(const IntegrationPoint &ip, Vector &s) const;
    Geometry::Type geom_ = Geometry::INVALID;
    int sdim_ = 0;
    DenseMatrix pm_;  // sdim x nverts
    const IntegrationPoint *ip_ = nullptr;
    int eval_ = 0;  // bit0 J, bit1 W, bit2 Jinv
    DenseMatrix J_, Jinv_, ds_;
    double w_ = 0.0;
};
typedef ElementTransformation IsoparametricTransformation;

class FaceElementTransformations
{
public:
    int Elem1No = -1, Elem2No = -1;
    ElementTransformation *Elem1 = nullptr, *Elem2 = nullptr;
    ElementTransformation *Face = nullptr;
    Geometry::Type GetGeometryType() const { return Face->GetGeometryType(); }
    /// NOTE: only FaceElementTransformations::SetIntPoint/SetAllIntPoints fill these; a caller that
    /// sets the point on `Face` alone reads the value-initialised (0,0,0) point.  The reference's
    /// face loop does exactly that, so the stand-in keeps the same contract.
    const IntegrationPoint &GetElement1IntPoint() const { return eip1_; }
    const IntegrationPoint &GetElement2IntPoint() const { return eip2_; }
    /// Proper mapping face point -> element points (used by the "consistent" integrator option).
    void SetAllIntPoints(const IntegrationPoint *face_ip);

    // local-face bookkeeping filled by Mesh (needed by SetAllIntPoints)
    int lf1 = -1, lf2 = -1;
    Geometry::Type egeom1 = Geometry::INVALID, egeom2 = Geometry::INVALID;
    int fv[4] = {-1, -1, -1, -1};
    int ev1[8], ev2[8];

private:
    IntegrationPoint eip1_, eip2_;
};

class InverseElementTransformation
{
public:
    enum InitGuessType { Center = 0, ClosestPhysNode = 1, ClosestRefNode = 2, GivenPoint = 3 };
    enum SolverType { Newton = 0, NewtonSegmentProject = 1, NewtonElementProject = 2 };
    enum TransformResult { Inside = 0, Outside = 1, Unknown = 2 };
    explicit InverseElementTransformation(ElementTransformation *T = nullptr) : T_(T) {}
    void SetTransformation(ElementTransformation &T) { T_ = &T; }
    void SetPrintLevel(int) {}
    void SetMaxIter(int it) { max_iter_ = it; }
    void SetReferenceTol(double t) { ref_tol_ = t; }
    void SetPhysicalRelTol(double t) { phys_rtol_ = t; }
    void SetElementTol(double t) { ip_tol_ = t; }
    void SetInitialGuessType(InitGuessType) {}
    void SetSolverType(SolverType) {}
    int Transform(const Vector &pt, IntegrationPoint &ip);

private:
    ElementTransformation *T_;
    int max_iter_ = 16;
    double ref_tol_ = 1e-15, phys_rtol_ = 1e-15, ip_tol_ = 1e-8;
};

// ----------------------------------------------------------------------------------------------
// finite elements (L2, Gauss–Legendre nodal basis)
// ----------------------------------------------------------------------------------------------
class FiniteElement
{
public:
    virtual ~FiniteElement() = default;
    int GetDof() const { return dof_; }
    int GetOrder() const { return order_; }
    int GetDim() const { return dim_; }
    Geometry::Type GetGeomType() const { return geom_; }
    virtual void CalcShape(const IntegrationPoint &ip, Vector &shape) const = 0;
    /// dshape is dof x dim (reference derivatives)
    virtual void CalcDShape(const IntegrationPoint &ip, DenseMatrix &dshape) const = 0;
    /// dshape = CalcDShape(Trans.GetIntPoint()) * Trans.InverseJacobian()
    void CalcPhysDShape(ElementTransformation &Trans, DenseMatrix &dshape) const;
    /// reference coordinates of the nodes (dof x dim)
    const DenseMatrix &Nodes() const { return nodes_; }

protected:
    Geometry::Type geom_ = Geometry::INVALID;
    int dim_ = 0, order_ = 0, dof_ = 0;
    DenseMatrix nodes_;
};

class BasisType
{
public:
    enum { Invalid = -1, GaussLegendre = 0, GaussLobatto = 1, Positive = 2 };
};

class FiniteElementCollection
{
public:
    virtual ~FiniteElementCollection() = default;
    virtual const FiniteElement *FiniteElementForGeometry(Geometry::Type g) const = 0;
    virtual const char *Name() const = 0;
    int GetOrder() const { return order_; }

protected:
    int order_ = 0;
};

class L2_FECollection : public FiniteElementCollection
{
public:
    L2_FECollection(int p, int dim, int btype = BasisType::GaussLegendre);
    const FiniteElement *FiniteElementForGeometry(Geometry::Type g) const override;
    const char *Name() const override { return name_.c_str(); }

private:
    int dim_;
    std::string name_;
    mutable std::map<int, std::unique_ptr<FiniteElement>> fes_;
};

class Ordering
{
public:
    enum Type { byNODES, byVDIM };
};

class GridFunction;

class FiniteElementSpace
{
public:
    FiniteElementSpace(Mesh *mesh, const FiniteElementCollection *fec, int vdim = 1,
                       int ordering = Ordering::byNODES);
    virtual ~FiniteElementSpace() = default;
    const FiniteElement *GetFE(int i) const;
    int GetNE() const;
    int GetNDofs() const { return ndofs_; }
    int GetVSize() const { return ndofs_ * vdim_; }
    int GetVDim() const { return vdim_; }
    Ordering::Type GetOrdering() const { return ordering_; }
    Mesh *GetMesh() const { return mesh_; }
    const FiniteElementCollection *FEColl() const { return fec_; }
    void GetElementDofs(int e, Array<int> &dofs) const;
    /// component-major list: all dofs of component 0, then component 1, ...
    void GetElementVDofs(int e, Array<int> &vdofs) const;

private:
    Mesh *mesh_;
    const FiniteElementCollection *fec_;
    int vdim_;
    Ordering::Type ordering_;
    int ndofs_ = 0;
    std::vector<int> offsets_;
};

class GridFunction : public Vector
{
public:
    GridFunction() = default;
    explicit GridFunction(FiniteElementSpace *fes) : Vector(fes->GetVSize()), fes_(fes) { Vector::operator=(0.0); }
    GridFunction &operator=(double v)
    {
        Vector::operator=(v);
        return *this;
    }
    FiniteElementSpace *FESpace() { return fes_; }
    const FiniteElementSpace *FESpace() const { return fes_; }
    /// value of component `vdim` (1-based as in MFEM) at a reference point of element e
    double GetValue(int e, const IntegrationPoint &ip, int vdim = 1) const;

private:
    FiniteElementSpace *fes_ = nullptr;
};

// ----------------------------------------------------------------------------------------------
// Mesh
// ----------------------------------------------------------------------------------------------
class Mesh
{
public:
    Mesh() = default;
    /// MFEM mesh v1.0 reader (straight-sided tri/quad/tet/hex).
    explicit Mesh(const std::string &filename, int generate_edges = 0, int refine = 1, bool fix_orientation = true);
    Mesh(const char *filename, int generate_edges, int refine, bool fix_orientation = true)
        : Mesh(std::string(filename), generate_edges, refine, fix_orientation) {}
    Mesh(const Mesh &o);
    Mesh(Mesh &&o) noexcept;
    Mesh &operator=(Mesh &&o) noexcept;

    static Mesh MakeCartesian2D(int nx, int ny, Element::Type type, bool generate_edges = false,
                                real_t sx = 1.0, real_t sy = 1.0, bool sfc_ordering = true);
    static Mesh MakeCartesian3D(int nx, int ny, int nz, Element::Type type, real_t sx = 1.0,
                                real_t sy = 1.0, real_t sz = 1.0, bool sfc_ordering = true);

    void UniformRefinement(int ref_algo = 0);

    int Dimension() const { return dim_; }
    int SpaceDimension() const { return sdim_; }
    int GetNE() const { return static_cast<int>(elems_.size()); }
    int GetNV() const { return static_cast<int>(verts_.size() / 3); }
    int GetNBE() const { return static_cast<int>(bdr_.size()); }
    int GetNumFaces() const { return static_cast<int>(faces_.size()); }
    int GetNEdges() const { return static_cast<int>(edges_.size()); }
    const real_t *GetVertex(int i) const { return &verts_[3 * static_cast<size_t>(i)]; }
    real_t *GetVertex(int i) { return &verts_[3 * static_cast<size_t>(i)]; }
    GridFunction *GetNodes() { return nullptr; }
    const GridFunction *GetNodes() const { return nullptr; }

    void GetElementVertices(int e, Array<int> &v) const;
    void GetFaceVertices(int f, Array<int> &v) const;
    void GetFaceElements(int f, int *e1, int *e2) const;
    int GetBdrElementFaceIndex(int be) const { return bdr_face_[be]; }
    int GetBdrAttribute(int be) const { return bdr_[be].attr; }
    int GetAttribute(int e) const { return elems_[e].attr; }
    Geometry::Type GetElementBaseGeometry(int e) const { return static_cast<Geometry::Type>(elems_[e].geom); }
    Geometry::Type GetElementGeometry(int e) const { return GetElementBaseGeometry(e); }
    Geometry::Type GetFaceGeometry(int f) const { return static_cast<Geometry::Type>(faces_[f].geom); }
    ElementTransformation *GetElementTransformation(int e);
    void GetElementTransformation(int e, ElementTransformation *tr) const;
    FaceElementTransformations *GetFaceElementTransformations(int f, int mask = 31);

    struct Cell
    {
        int geom = -1, attr = 1, nv = 0;
        int v[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    };
    struct Face
    {
        int geom = -1, nv = 0;
        int v[4] = {-1, -1, -1, -1};
        int e1 = -1, e2 = -1, lf1 = -1, lf2 = -1;
    };
    const Cell &GetCell(int e) const { return elems_[e]; }
    const Face &GetFaceRec(int f) const { return faces_[f]; }

    /// local face -> local vertex tables
    static int NumLocalFaces(int geom);
    static int LocalFaceVerts(int geom, int lf, int *lv);
    static int NumLocalEdges(int geom);
    static void LocalEdgeVerts(int geom, int le, int *lv);

private:
    void Finalize(bool refine, bool fix_orientation);
    void FixOrientation();
    void MarkForRefinement();
    void BuildEdges();
    void BuildFaces();
    void AddVertex(double x, double y, double z = 0.0);
    int AddCell(int geom, int attr, const int *v);
    void AddBdr(int geom, int attr, const int *v);
    void GenerateBoundary(double sx, double sy, double sz);

    int dim_ = 0, sdim_ = 0;
    std::vector<double> verts_;  // always 3 per vertex (z = 0 in 2-D)
    std::vector<Cell> elems_, bdr_;
    std::vector<Face> faces_;
    std::vector<std::pair<int, int>> edges_;       // (v0, v1) in first-appearance orientation
    std::vector<std::vector<int>> el_to_edge_;     // per element, local edge -> edge id
    std::vector<std::vector<int>> el_to_face_;     // per element, local face -> face id
    std::vector<int> bdr_face_;
    ElementTransformation tr1_, tr2_, trf_;
    FaceElementTransformations ftr_;
};

// ----------------------------------------------------------------------------------------------
// output / CLI helpers
// ----------------------------------------------------------------------------------------------
enum class VTKFormat { ASCII, BINARY, BINARY32 };

/// Stand-in writer: instead of ParaView XML it writes, per registered field, the values at the
/// vertices of every element (what a level-of-detail-1 ParaView dump contains) as plain text to
/// <prefix_path>/<name>/vertex_values.txt.
class ParaViewDataCollection
{
public:
    ParaViewDataCollection(const std::string &name, Mesh *mesh) : name_(name), mesh_(mesh) {}
    void SetPrefixPath(const std::string &p) { prefix_ = p; }
    void SetDataFormat(VTKFormat) {}
    void SetHighOrderOutput(bool) {}
    void SetLevelsOfDetail(int) {}
    void SetCycle(int) {}
    void SetTime(double) {}
    void RegisterField(const std::string &field_name, GridFunction *gf) { fields_.emplace_back(field_name, gf); }
    void Save();

private:
    std::string name_, prefix_;
    Mesh *mesh_;
    std::vector<std::pair<std::string, GridFunction *>> fields_;
};

class OptionsParser
{
public:
    OptionsParser(int argc, char *argv[]) : argc_(argc), argv_(argv) {}
    void AddOption(std::string *var, const char *s, const char *l, const char *descr, bool req = false);
    void AddOption(const char **var, const char *s, const char *l, const char *descr, bool req = false);
    void AddOption(int *var, const char *s, const char *l, const char *descr, bool req = false);
    void AddOption(double *var, const char *s, const char *l, const char *descr, bool req = false);
    void AddOption(bool *var, const char *en_s, const char *en_l, const char *dis_s, const char *dis_l,
                   const char *descr, bool req = false);
    void Parse();
    bool Good() const { return good_; }
    void PrintUsage(std::ostream &os) const;
    void PrintOptions(std::ostream &os) const;

private:
    struct Opt
    {
        int kind;  // 0 string, 1 cstr, 2 int, 3 double, 4 bool
        void *var;
        std::string s, l, ds, dl, descr;
    };
    int argc_;
    char **argv_;
    std::vector<Opt> opts_;
    bool good_ = true;
};

}  // namespace mfem
