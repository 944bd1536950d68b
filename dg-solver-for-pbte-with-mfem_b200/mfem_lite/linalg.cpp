// mfem_lite: Vector / DenseMatrix / DenseMatrixInverse.
// Operation ORDER follows MFEM 4.x where results feed the PBTE hot path [from memory of MFEM 4.x]:
//   * Vector::operator*      sequential sum i = 0..n-1
//   * Vector::Norml2         scaled (dnrm2-style) accumulation
//   * DenseMatrix::Mult      column-by-column accumulation
//   * DenseMatrixInverse     in-place row-pivoted LU, first maximal pivot; Mult = permute,
//                            unit-lower forward sweep, upper backward sweep (column oriented)
#include "mfem.hpp"

namespace mfem
{
static ErrorAction g_error_action = MFEM_ERROR_ABORT;
void set_error_action(ErrorAction a) { g_error_action = a; }
ErrorAction get_error_action() { return g_error_action; }

void mfem_error(const char *msg)
{
    if (g_error_action == MFEM_ERROR_THROW)
    {
        throw ErrorException(msg ? msg : "mfem_error");
    }
    if (msg) std::cerr << "\n\n" << msg << "\n";
    std::abort();
}

// ------------------------------------------------------------------ Vector
Vector::Vector(const Vector &o)
{
    SetSize(o.size_);
    if (size_) std::memcpy(data_, o.data_, sizeof(double) * size_);
}

Vector::Vector(Vector &&o) noexcept : data_(o.data_), size_(o.size_), cap_(o.cap_)
{
    o.data_ = nullptr;
    o.size_ = o.cap_ = 0;
}

Vector::~Vector()
{
    if (cap_ > 0) delete[] data_;
}

Vector &Vector::operator=(const Vector &o)
{
    if (this == &o) return *this;
    SetSize(o.size_);
    if (size_) std::memcpy(data_, o.data_, sizeof(double) * size_);
    return *this;
}

Vector &Vector::operator=(Vector &&o) noexcept
{
    if (this == &o) return *this;
    if (cap_ > 0) delete[] data_;
    data_ = o.data_;
    size_ = o.size_;
    cap_ = o.cap_;
    o.data_ = nullptr;
    o.size_ = o.cap_ = 0;
    return *this;
}

Vector &Vector::operator=(double v)
{
    for (int i = 0; i < size_; ++i) data_[i] = v;
    return *this;
}

void Vector::SetSize(int n)
{
    if (n <= cap_ || (cap_ == 0 && data_ != nullptr && n <= size_))
    {
        size_ = n;
        return;
    }
    double *nd = n > 0 ? new double[n] : nullptr;
    if (cap_ > 0) delete[] data_;
    data_ = nd;
    size_ = n;
    cap_ = n;
}

double Vector::operator*(const Vector &v) const { return (*this) * v.data_; }

double Vector::operator*(const double *v) const
{
    double prod = 0.0;
    for (int i = 0; i < size_; ++i) prod += data_[i] * v[i];
    return prod;
}

Vector &Vector::operator*=(double c)
{
    for (int i = 0; i < size_; ++i) data_[i] *= c;
    return *this;
}

Vector &Vector::operator/=(double c)
{
    const double m = 1.0 / c;
    for (int i = 0; i < size_; ++i) data_[i] *= m;
    return *this;
}

Vector &Vector::operator-=(const Vector &v)
{
    for (int i = 0; i < size_; ++i) data_[i] -= v.data_[i];
    return *this;
}

Vector &Vector::operator+=(const Vector &v)
{
    for (int i = 0; i < size_; ++i) data_[i] += v.data_[i];
    return *this;
}

Vector &Vector::Add(double a, const Vector &x)
{
    if (a != 0.0)
    {
        for (int i = 0; i < size_; ++i) data_[i] += a * x.data_[i];
    }
    return *this;
}

double Vector::Norml2() const
{
    if (size_ == 0) return 0.0;
    if (size_ == 1) return std::abs(data_[0]);
    double scale = 0.0, sum = 0.0;
    for (int i = 0; i < size_; ++i)
    {
        if (data_[i] != 0.0)
        {
            const double a = std::fabs(data_[i]);
            if (scale <= a)
            {
                const double r = scale / a;
                sum = 1.0 + sum * (r * r);
                scale = a;
                continue;
            }
            const double r = a / scale;
            sum += r * r;
        }
    }
    return scale * std::sqrt(sum);
}

double Vector::Max() const
{
    double m = data_[0];
    for (int i = 1; i < size_; ++i) m = std::max(m, data_[i]);
    return m;
}

double Vector::Min() const
{
    double m = data_[0];
    for (int i = 1; i < size_; ++i) m = std::min(m, data_[i]);
    return m;
}

void Vector::Swap(Vector &o)
{
    std::swap(data_, o.data_);
    std::swap(size_, o.size_);
    std::swap(cap_, o.cap_);
}

void Vector::SetSubVector(const Array<int> &dofs, const Vector &ev)
{
    for (int i = 0; i < dofs.Size(); ++i)
    {
        const int j = dofs[i];
        if (j >= 0) data_[j] = ev[i];
        else data_[-1 - j] = -ev[i];
    }
}

void Vector::GetSubVector(const Array<int> &dofs, Vector &ev) const
{
    ev.SetSize(dofs.Size());
    for (int i = 0; i < dofs.Size(); ++i)
    {
        const int j = dofs[i];
        ev[i] = j >= 0 ? data_[j] : -data_[-1 - j];
    }
}

// ------------------------------------------------------------------ DenseMatrix
DenseMatrix::DenseMatrix(const DenseMatrix &o) : d_(o.Data(), o.Data() + o.Count()), h_(o.h_), w_(o.w_) {}

DenseMatrix::DenseMatrix(DenseMatrix &&o) noexcept : d_(std::move(o.d_)), ext_(o.ext_), h_(o.h_), w_(o.w_)
{
    o.ext_ = nullptr;
    o.h_ = o.w_ = 0;
}

DenseMatrix &DenseMatrix::operator=(const DenseMatrix &o)
{
    if (this == &o) return *this;
    if (h_ != o.h_ || w_ != o.w_)
    {
        ext_ = nullptr;
        h_ = o.h_;
        w_ = o.w_;
        d_.assign(o.Data(), o.Data() + o.Count());
    }
    else if (Count()) std::memcpy(Data(), o.Data(), sizeof(double) * Count());
    return *this;
}

DenseMatrix &DenseMatrix::operator=(DenseMatrix &&o) noexcept
{
    d_ = std::move(o.d_);
    ext_ = o.ext_;
    h_ = o.h_;
    w_ = o.w_;
    o.ext_ = nullptr;
    o.h_ = o.w_ = 0;
    return *this;
}

DenseMatrix &DenseMatrix::operator=(double v)
{
    std::fill(Data(), Data() + Count(), v);
    return *this;
}

void DenseMatrix::SetSize(int h, int w)
{
    if (h == h_ && w == w_) return;
    ext_ = nullptr;
    h_ = h;
    w_ = w;
    d_.assign(static_cast<size_t>(h) * w, 0.0);
}

void DenseMatrix::UseExternalData(double *d, int h, int w)
{
    std::vector<double>().swap(d_);
    ext_ = d;
    h_ = h;
    w_ = w;
}

DenseMatrix &DenseMatrix::operator*=(double c)
{
    double *p = Data();
    for (size_t i = 0, n = Count(); i < n; ++i) p[i] *= c;
    return *this;
}

DenseMatrix &DenseMatrix::operator+=(const DenseMatrix &A)
{
    double *p = Data();
    const double *q = A.Data();
    for (size_t i = 0, n = Count(); i < n; ++i) p[i] += q[i];
    return *this;
}

void DenseMatrix::Add(double c, const DenseMatrix &A)
{
    for (int j = 0; j < w_; ++j)
        for (int i = 0; i < h_; ++i) (*this)(i, j) += c * A(i, j);
}

void DenseMatrix::Mult(const double *x, double *y) const
{
    if (w_ == 0)
    {
        for (int r = 0; r < h_; ++r) y[r] = 0.0;
        return;
    }
    const double *col = Data();
    double xc = x[0];
    for (int r = 0; r < h_; ++r) y[r] = xc * col[r];
    col += h_;
    for (int c = 1; c < w_; ++c)
    {
        xc = x[c];
        for (int r = 0; r < h_; ++r) y[r] += xc * col[r];
        col += h_;
    }
}

void DenseMatrix::Mult(const Vector &x, Vector &y) const { Mult(x.GetData(), y.GetData()); }

void DenseMatrix::MultTranspose(const Vector &x, Vector &y) const
{
    const double *col = Data();
    for (int c = 0; c < w_; ++c)
    {
        double s = 0.0;
        for (int r = 0; r < h_; ++r) s += x[r] * col[r];
        y[c] = s;
        col += h_;
    }
}

void DenseMatrix::GetColumn(int c, Vector &col) const
{
    col.SetSize(h_);
    std::memcpy(col.GetData(), Data() + static_cast<size_t>(c) * h_, sizeof(double) * h_);
}

void DenseMatrix::SetCol(int c, const Vector &col)
{
    std::memcpy(Data() + static_cast<size_t>(c) * h_, col.GetData(), sizeof(double) * h_);
}

void DenseMatrix::Transpose()
{
    if (h_ == w_)
    {
        for (int i = 0; i < h_; ++i)
            for (int j = 0; j < i; ++j) std::swap((*this)(i, j), (*this)(j, i));
        return;
    }
    DenseMatrix t(w_, h_);
    for (int i = 0; i < h_; ++i)
        for (int j = 0; j < w_; ++j) t(j, i) = (*this)(i, j);
    *this = std::move(t);
}

void DenseMatrix::Swap(DenseMatrix &o)
{
    d_.swap(o.d_);
    std::swap(ext_, o.ext_);
    std::swap(h_, o.h_);
    std::swap(w_, o.w_);
}

double DenseMatrix::FNorm() const
{
    double mx = 0.0;
    const double *p = Data();
    const size_t cnt = Count();
    for (size_t i = 0; i < cnt; ++i) mx = std::max(mx, std::fabs(p[i]));
    if (mx == 0.0) return 0.0;
    double s = 0.0;
    for (size_t i = 0; i < cnt; ++i)
    {
        const double e = p[i] / mx;
        s += e * e;
    }
    return mx * std::sqrt(s);
}

double DenseMatrix::Det() const
{
    const double *d = Data();
    switch (h_)
    {
        case 0: return 1.0;
        case 1: return d[0];
        case 2: return d[0] * d[3] - d[1] * d[2];
        case 3:
            return d[0] * (d[4] * d[8] - d[5] * d[7]) + d[3] * (d[2] * d[7] - d[1] * d[8]) +
                   d[6] * (d[1] * d[5] - d[2] * d[4]);
        default: break;
    }
    DenseMatrixInverse lu(*this);
    double det = 1.0;
    for (int i = 0; i < h_; ++i)
    {
        det *= lu.LUData()[i + static_cast<size_t>(i) * h_];
        if (lu.Pivots()[i] != i) det = -det;
    }
    return det;
}

double DenseMatrix::Weight() const
{
    if (h_ == w_) return Det();
    const double *d = Data();
    if (h_ == 2 && w_ == 1) return std::sqrt(d[0] * d[0] + d[1] * d[1]);
    if (h_ == 3 && w_ == 1) return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (h_ == 3 && w_ == 2)
    {
        const double E = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double G = d[3] * d[3] + d[4] * d[4] + d[5] * d[5];
        const double F = d[0] * d[3] + d[1] * d[4] + d[2] * d[5];
        return std::sqrt(E * G - F * F);
    }
    mfem_error("DenseMatrix::Weight: unsupported shape");
}

void Mult(const DenseMatrix &A, const DenseMatrix &B, DenseMatrix &C)
{
    C.SetSize(A.Height(), B.Width());
    C = 0.0;
    for (int j = 0; j < B.Width(); ++j)
        for (int k = 0; k < A.Width(); ++k)
        {
            const double b = B(k, j);
            for (int i = 0; i < A.Height(); ++i) C(i, j) += A(i, k) * b;
        }
}

void CalcInverse(const DenseMatrix &A, DenseMatrix &Ainv)
{
    const int h = A.Height(), w = A.Width();
    Ainv.SetSize(w, h);
    const double *a = A.Data();
    double *inv = Ainv.Data();
    if (h == w)
    {
        const double t = 1.0 / A.Det();
        switch (h)
        {
            case 1: inv[0] = t; return;
            case 2:
                inv[0] = a[3] * t;
                inv[1] = -a[1] * t;
                inv[2] = -a[2] * t;
                inv[3] = a[0] * t;
                return;
            case 3:
                inv[0] = (a[4] * a[8] - a[5] * a[7]) * t;
                inv[3] = (a[5] * a[6] - a[3] * a[8]) * t;
                inv[6] = (a[3] * a[7] - a[4] * a[6]) * t;
                inv[1] = (a[2] * a[7] - a[1] * a[8]) * t;
                inv[4] = (a[0] * a[8] - a[2] * a[6]) * t;
                inv[7] = (a[1] * a[6] - a[0] * a[7]) * t;
                inv[2] = (a[1] * a[5] - a[2] * a[4]) * t;
                inv[5] = (a[2] * a[3] - a[0] * a[5]) * t;
                inv[8] = (a[0] * a[4] - a[1] * a[3]) * t;
                return;
            default: mfem_error("CalcInverse: size > 3 not supported");
        }
    }
    // left pseudo-inverse (J^T J)^{-1} J^T
    if (w == 1)
    {
        double s = 0.0;
        for (int i = 0; i < h; ++i) s += a[i] * a[i];
        for (int i = 0; i < h; ++i) inv[i] = a[i] / s;
        return;
    }
    if (h == 3 && w == 2)
    {
        const double e = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
        const double g = a[3] * a[3] + a[4] * a[4] + a[5] * a[5];
        const double f = a[0] * a[3] + a[1] * a[4] + a[2] * a[5];
        const double t = 1.0 / (e * g - f * f);
        for (int i = 0; i < 3; ++i)
        {
            Ainv(0, i) = (g * a[i] - f * a[3 + i]) * t;
            Ainv(1, i) = (e * a[3 + i] - f * a[i]) * t;
        }
        return;
    }
    mfem_error("CalcInverse: unsupported shape");
}

// ------------------------------------------------------------------ DenseMatrixInverse
DenseMatrixInverse::DenseMatrixInverse(const DenseMatrix &mat) : a_(&mat) { Factor(); }
DenseMatrixInverse::DenseMatrixInverse(const DenseMatrix *mat) : a_(mat) { Factor(); }

void DenseMatrixInverse::Factor(const DenseMatrix &mat)
{
    a_ = &mat;
    Factor();
}

void DenseMatrixInverse::Factor()
{
    MFEM_VERIFY(a_ != nullptr && a_->Height() == a_->Width(), "DenseMatrixInverse: square matrix required");
    const int m = n_ = a_->Height();
    lu_.assign(a_->Data(), a_->Data() + static_cast<size_t>(m) * m);
    ipiv_.assign(m, 0);
    double *data = lu_.data();
    for (int i = 0; i < m; ++i)
    {
        int piv = i;
        double a = std::abs(data[piv + static_cast<size_t>(i) * m]);
        for (int j = i + 1; j < m; ++j)
        {
            const double b = std::abs(data[j + static_cast<size_t>(i) * m]);
            if (b > a)
            {
                a = b;
                piv = j;
            }
        }
        ipiv_[i] = piv;
        if (piv != i)
        {
            for (int j = 0; j < m; ++j) std::swap(data[i + static_cast<size_t>(j) * m], data[piv + static_cast<size_t>(j) * m]);
        }
        MFEM_VERIFY(data[i + static_cast<size_t>(i) * m] != 0.0, "DenseMatrixInverse::Factor: singular matrix");
        const double a_ii_inv = 1.0 / data[i + static_cast<size_t>(i) * m];
        for (int j = i + 1; j < m; ++j) data[j + static_cast<size_t>(i) * m] *= a_ii_inv;
        for (int k = i + 1; k < m; ++k)
        {
            const double a_ik = data[i + static_cast<size_t>(k) * m];
            for (int j = i + 1; j < m; ++j) data[j + static_cast<size_t>(k) * m] -= a_ik * data[j + static_cast<size_t>(i) * m];
        }
    }
}

void DenseMatrixInverse::Mult(const double *xin, double *x) const
{
    const int m = n_;
    const double *data = lu_.data();
    if (x != xin) std::memcpy(x, xin, sizeof(double) * m);
    for (int i = 0; i < m; ++i) std::swap(x[i], x[ipiv_[i]]);
    for (int j = 0; j < m; ++j)
    {
        const double x_j = x[j];
        for (int i = j + 1; i < m; ++i) x[i] -= data[i + static_cast<size_t>(j) * m] * x_j;
    }
    for (int j = m - 1; j >= 0; --j)
    {
        const double x_j = (x[j] /= data[j + static_cast<size_t>(j) * m]);
        for (int i = 0; i < j; ++i) x[i] -= data[i + static_cast<size_t>(j) * m] * x_j;
    }
}

void DenseMatrixInverse::Mult(const Vector &x, Vector &y) const
{
    y.SetSize(n_);
    Mult(x.GetData(), y.GetData());
}

}  // namespace mfem
