// mfem_lite: straight-sided conforming meshes of triangles / quadrilaterals / tetrahedra / hexahedra.
//
// Numbering conventions, and how each one is known:
//   [log-confirmed]  pinned by the reference's shipped output/log/mesh_unit-square-iso_p{1,2}_dim2.txt
//     * file reader with refine=1 rotates every triangle so that its longest edge is edge 0-1
//     * 2-D faces (= edges) are numbered by first appearance, elements ascending, local edges
//       (0-1, 1-2, 2-0)
//     * uniform triangle refinement: new vertex = NV_old + edge id; children of element i are
//       4i..4i+3 = (v0,m01,m20) (m12,m20,m01) (m01,v1,m12) (m20,m12,v2); edges renumbered by the
//       same first-appearance rule; no re-rotation
//   [from memory of MFEM 4.x]  everything else (Cartesian generators, space-filling-curve ordering of
//     quads/hexes, tetrahedron marking, hex/tet/quad refinement tables, 3-D face tables).  None of
//     these affects solver parity on identical arrays — only whether a mesh generated here carries
//     the numbering real MFEM would produce.
#include "mfem.hpp"

#include <array>
#include <fstream>
#include <numeric>
#include <unordered_map>

namespace mfem
{
namespace
{
const int kTriEdges[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int kQuadEdges[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
const int kTetEdges[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
const int kHexEdges[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
const int kTetFaces[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
const int kHexFaces[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};

struct KeyHash
{
    size_t operator()(const std::array<int, 4> &k) const
    {
        size_t h = 1469598103934665603ull;
        for (int x : k)
        {
            h ^= static_cast<size_t>(static_cast<unsigned>(x));
            h *= 1099511628211ull;
        }
        return h;
    }
};

std::array<int, 4> SortedKey(const int *v, int n)
{
    std::array<int, 4> k = {-1, -1, -1, -1};
    for (int i = 0; i < n; ++i)
    {
        int j = i;
        for (; j > 0 && k[j - 1] > v[i]; --j) k[j] = k[j - 1];
        k[j] = v[i];
    }
    return k;
}

inline int sgn(int x) { return (x > 0) - (x < 0); }

// Generalised Hilbert ("gilbert") curve of J. Cerveny, 2-D and 3-D; visits every cell of a w x h (x d)
// grid once with unit steps whenever the sizes allow it.
void Gilbert2D(int x, int y, int ax, int ay, int bx, int by, std::vector<int> &out)
{
    const int w = std::abs(ax + ay), h = std::abs(bx + by);
    const int dax = sgn(ax), day = sgn(ay), dbx = sgn(bx), dby = sgn(by);
    if (h == 1)
    {
        for (int i = 0; i < w; ++i, x += dax, y += day) { out.push_back(x); out.push_back(y); }
        return;
    }
    if (w == 1)
    {
        for (int i = 0; i < h; ++i, x += dbx, y += dby) { out.push_back(x); out.push_back(y); }
        return;
    }
    int ax2 = ax / 2, ay2 = ay / 2, bx2 = bx / 2, by2 = by / 2;
    const int w2 = std::abs(ax2 + ay2), h2 = std::abs(bx2 + by2);
    if (2 * w > 3 * h)
    {
        if ((w2 & 1) && w > 2) { ax2 += dax; ay2 += day; }
        Gilbert2D(x, y, ax2, ay2, bx, by, out);
        Gilbert2D(x + ax2, y + ay2, ax - ax2, ay - ay2, bx, by, out);
    }
    else
    {
        if ((h2 & 1) && h > 2) { bx2 += dbx; by2 += dby; }
        Gilbert2D(x, y, bx2, by2, ax2, ay2, out);
        Gilbert2D(x + bx2, y + by2, ax, ay, bx - bx2, by - by2, out);
        Gilbert2D(x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), -bx2, -by2, -(ax - ax2), -(ay - ay2), out);
    }
}

void Gilbert3D(int x, int y, int z, int ax, int ay, int az, int bx, int by, int bz, int cx, int cy, int cz,
               std::vector<int> &out)
{
    const int w = std::abs(ax + ay + az), h = std::abs(bx + by + bz), d = std::abs(cx + cy + cz);
    const int dax = sgn(ax), day = sgn(ay), daz = sgn(az);
    const int dbx = sgn(bx), dby = sgn(by), dbz = sgn(bz);
    const int dcx = sgn(cx), dcy = sgn(cy), dcz = sgn(cz);
    auto emit = [&](int n, int sx, int sy, int sz) {
        for (int i = 0; i < n; ++i, x += sx, y += sy, z += sz) { out.push_back(x); out.push_back(y); out.push_back(z); }
    };
    if (h == 1 && d == 1) { emit(w, dax, day, daz); return; }
    if (w == 1 && d == 1) { emit(h, dbx, dby, dbz); return; }
    if (w == 1 && h == 1) { emit(d, dcx, dcy, dcz); return; }
    int ax2 = ax / 2, ay2 = ay / 2, az2 = az / 2;
    int bx2 = bx / 2, by2 = by / 2, bz2 = bz / 2;
    int cx2 = cx / 2, cy2 = cy / 2, cz2 = cz / 2;
    const int w2 = std::abs(ax2 + ay2 + az2), h2 = std::abs(bx2 + by2 + bz2), d2 = std::abs(cx2 + cy2 + cz2);
    if ((w2 & 1) && w > 2) { ax2 += dax; ay2 += day; az2 += daz; }
    if ((h2 & 1) && h > 2) { bx2 += dbx; by2 += dby; bz2 += dbz; }
    if ((d2 & 1) && d > 2) { cx2 += dcx; cy2 += dcy; cz2 += dcz; }
    if (2 * w > 3 * h && 2 * w > 3 * d)
    {
        Gilbert3D(x, y, z, ax2, ay2, az2, bx, by, bz, cx, cy, cz, out);
        Gilbert3D(x + ax2, y + ay2, z + az2, ax - ax2, ay - ay2, az - az2, bx, by, bz, cx, cy, cz, out);
    }
    else if (3 * h > 4 * d)
    {
        Gilbert3D(x, y, z, bx2, by2, bz2, cx, cy, cz, ax2, ay2, az2, out);
        Gilbert3D(x + bx2, y + by2, z + bz2, ax, ay, az, bx - bx2, by - by2, bz - bz2, cx, cy, cz, out);
        Gilbert3D(x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), z + (az - daz) + (bz2 - dbz),
                  -bx2, -by2, -bz2, cx, cy, cz, -(ax - ax2), -(ay - ay2), -(az - az2), out);
    }
    else if (3 * d > 4 * h)
    {
        Gilbert3D(x, y, z, cx2, cy2, cz2, ax2, ay2, az2, bx, by, bz, out);
        Gilbert3D(x + cx2, y + cy2, z + cz2, ax, ay, az, bx, by, bz, cx - cx2, cy - cy2, cz - cz2, out);
        Gilbert3D(x + (ax - dax) + (cx2 - dcx), y + (ay - day) + (cy2 - dcy), z + (az - daz) + (cz2 - dcz),
                  -cx2, -cy2, -cz2, -(ax - ax2), -(ay - ay2), -(az - az2), bx, by, bz, out);
    }
    else
    {
        Gilbert3D(x, y, z, bx2, by2, bz2, cx2, cy2, cz2, ax2, ay2, az2, out);
        Gilbert3D(x + bx2, y + by2, z + bz2, cx, cy, cz, ax2, ay2, az2, bx - bx2, by - by2, bz - bz2, out);
        Gilbert3D(x + (bx2 - dbx) + (cx - dcx), y + (by2 - dby) + (cy - dcy), z + (bz2 - dbz) + (cz - dcz),
                  ax, ay, az, -bx2, -by2, -bz2, -(cx - cx2), -(cy - cy2), -(cz - cz2), out);
        Gilbert3D(x + (ax - dax) + bx2 + (cx - dcx), y + (ay - day) + by2 + (cy - dcy), z + (az - daz) + bz2 + (cz - dcz),
                  -cx, -cy, -cz, -(ax - ax2), -(ay - ay2), -(az - az2), bx - bx2, by - by2, bz - bz2, out);
        Gilbert3D(x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), z + (az - daz) + (bz2 - dbz),
                  -bx2, -by2, -bz2, cx2, cy2, cz2, -(ax - ax2), -(ay - ay2), -(az - az2), out);
    }
}
}  // namespace

// ------------------------------------------------------------------ static tables
int Mesh::NumLocalFaces(int geom)
{
    switch (geom)
    {
        case Geometry::TRIANGLE: return 3;
        case Geometry::SQUARE: return 4;
        case Geometry::TETRAHEDRON: return 4;
        case Geometry::CUBE: return 6;
        default: return 0;
    }
}

int Mesh::LocalFaceVerts(int geom, int lf, int *lv)
{
    switch (geom)
    {
        case Geometry::TRIANGLE: lv[0] = kTriEdges[lf][0]; lv[1] = kTriEdges[lf][1]; return 2;
        case Geometry::SQUARE: lv[0] = kQuadEdges[lf][0]; lv[1] = kQuadEdges[lf][1]; return 2;
        case Geometry::TETRAHEDRON: for (int i = 0; i < 3; ++i) lv[i] = kTetFaces[lf][i]; return 3;
        case Geometry::CUBE: for (int i = 0; i < 4; ++i) lv[i] = kHexFaces[lf][i]; return 4;
        default: return 0;
    }
}

int Mesh::NumLocalEdges(int geom)
{
    switch (geom)
    {
        case Geometry::SEGMENT: return 1;
        case Geometry::TRIANGLE: return 3;
        case Geometry::SQUARE: return 4;
        case Geometry::TETRAHEDRON: return 6;
        case Geometry::CUBE: return 12;
        default: return 0;
    }
}

void Mesh::LocalEdgeVerts(int geom, int le, int *lv)
{
    const int(*t)[2] = nullptr;
    switch (geom)
    {
        case Geometry::TRIANGLE: t = kTriEdges; break;
        case Geometry::SQUARE: t = kQuadEdges; break;
        case Geometry::TETRAHEDRON: t = kTetEdges; break;
        case Geometry::CUBE: t = kHexEdges; break;
        default: mfem_error("LocalEdgeVerts: unsupported geometry");
    }
    lv[0] = t[le][0];
    lv[1] = t[le][1];
}

// ------------------------------------------------------------------ construction helpers
Mesh::Mesh(const Mesh &o)
    : dim_(o.dim_), sdim_(o.sdim_), verts_(o.verts_), elems_(o.elems_), bdr_(o.bdr_), faces_(o.faces_),
      edges_(o.edges_), el_to_edge_(o.el_to_edge_), el_to_face_(o.el_to_face_), bdr_face_(o.bdr_face_)
{
}

Mesh::Mesh(Mesh &&o) noexcept { *this = std::move(o); }

Mesh &Mesh::operator=(Mesh &&o) noexcept
{
    dim_ = o.dim_;
    sdim_ = o.sdim_;
    verts_ = std::move(o.verts_);
    elems_ = std::move(o.elems_);
    bdr_ = std::move(o.bdr_);
    faces_ = std::move(o.faces_);
    edges_ = std::move(o.edges_);
    el_to_edge_ = std::move(o.el_to_edge_);
    el_to_face_ = std::move(o.el_to_face_);
    bdr_face_ = std::move(o.bdr_face_);
    return *this;
}

void Mesh::AddVertex(double x, double y, double z)
{
    verts_.push_back(x);
    verts_.push_back(y);
    verts_.push_back(z);
}

int Mesh::AddCell(int geom, int attr, const int *v)
{
    Cell c;
    c.geom = geom;
    c.attr = attr;
    c.nv = Geometry::NumVerts[geom];
    for (int i = 0; i < c.nv; ++i) c.v[i] = v[i];
    elems_.push_back(c);
    return static_cast<int>(elems_.size()) - 1;
}

void Mesh::AddBdr(int geom, int attr, const int *v)
{
    Cell c;
    c.geom = geom;
    c.attr = attr;
    c.nv = Geometry::NumVerts[geom];
    for (int i = 0; i < c.nv; ++i) c.v[i] = v[i];
    bdr_.push_back(c);
}

// ------------------------------------------------------------------ file reader
Mesh::Mesh(const std::string &filename, int, int refine, bool fix_orientation)
{
    std::ifstream in(filename);
    MFEM_VERIFY(static_cast<bool>(in), "Mesh file not found: " << filename);
    // tokenise, dropping comments
    std::vector<std::string> tok;
    std::string line;
    bool first = true;
    while (std::getline(in, line))
    {
        const auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (first)
        {
            first = false;
            MFEM_VERIFY(line.find("MFEM mesh v1.0") != std::string::npos, "unsupported mesh format (need 'MFEM mesh v1.0')");
            continue;
        }
        std::istringstream ls(line);
        std::string t;
        while (ls >> t) tok.push_back(t);
    }
    size_t p = 0;
    auto next = [&]() -> const std::string & {
        MFEM_VERIFY(p < tok.size(), "unexpected end of mesh file");
        return tok[p++];
    };
    auto expect = [&](const char *kw) { MFEM_VERIFY(next() == kw, "mesh file: expected '" << kw << "'"); };
    expect("dimension");
    dim_ = std::stoi(next());
    expect("elements");
    const int ne = std::stoi(next());
    for (int i = 0; i < ne; ++i)
    {
        const int attr = std::stoi(next());
        const int geom = std::stoi(next());
        MFEM_VERIFY(geom >= Geometry::TRIANGLE && geom <= Geometry::CUBE, "unsupported element geometry " << geom);
        int v[8];
        for (int k = 0; k < Geometry::NumVerts[geom]; ++k) v[k] = std::stoi(next());
        AddCell(geom, attr, v);
    }
    expect("boundary");
    const int nb = std::stoi(next());
    for (int i = 0; i < nb; ++i)
    {
        const int attr = std::stoi(next());
        const int geom = std::stoi(next());
        int v[8];
        for (int k = 0; k < Geometry::NumVerts[geom]; ++k) v[k] = std::stoi(next());
        AddBdr(geom, attr, v);
    }
    expect("vertices");
    const int nv = std::stoi(next());
    sdim_ = std::stoi(next());
    for (int i = 0; i < nv; ++i)
    {
        double c[3] = {0, 0, 0};
        for (int d = 0; d < sdim_; ++d) c[d] = std::stod(next());
        AddVertex(c[0], c[1], c[2]);
    }
    Finalize(refine != 0, fix_orientation);
}

// ------------------------------------------------------------------ topology
void Mesh::FixOrientation()
{
    for (Cell &c : elems_)
    {
        const double *v0 = GetVertex(c.v[0]), *v1 = GetVertex(c.v[1]), *v2 = GetVertex(c.v[2]);
        if (c.geom == Geometry::TRIANGLE)
        {
            const double det = (v1[0] - v0[0]) * (v2[1] - v0[1]) - (v1[1] - v0[1]) * (v2[0] - v0[0]);
            if (det < 0.0) std::swap(c.v[0], c.v[1]);
        }
        else if (c.geom == Geometry::SQUARE)
        {
            const double det = (v1[0] - v0[0]) * (v2[1] - v0[1]) - (v1[1] - v0[1]) * (v2[0] - v0[0]);
            if (det < 0.0) std::swap(c.v[1], c.v[3]);
        }
        else if (c.geom == Geometry::TETRAHEDRON)
        {
            const double *v3 = GetVertex(c.v[3]);
            const double a[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
            const double b[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
            const double d[3] = {v3[0] - v0[0], v3[1] - v0[1], v3[2] - v0[2]};
            const double det = a[0] * (b[1] * d[2] - b[2] * d[1]) - a[1] * (b[0] * d[2] - b[2] * d[0]) + a[2] * (b[0] * d[1] - b[1] * d[0]);
            if (det < 0.0) std::swap(c.v[0], c.v[1]);
        }
    }
}

void Mesh::BuildEdges()
{
    edges_.clear();
    el_to_edge_.assign(elems_.size(), {});
    std::unordered_map<unsigned long long, int> emap;
    emap.reserve(elems_.size() * 4);
    for (size_t e = 0; e < elems_.size(); ++e)
    {
        const Cell &c = elems_[e];
        const int nle = NumLocalEdges(c.geom);
        el_to_edge_[e].resize(nle);
        for (int le = 0; le < nle; ++le)
        {
            int lv[2];
            LocalEdgeVerts(c.geom, le, lv);
            const int a = c.v[lv[0]], b = c.v[lv[1]];
            const unsigned long long key = (static_cast<unsigned long long>(std::min(a, b)) << 32) | static_cast<unsigned>(std::max(a, b));
            auto it = emap.find(key);
            int id;
            if (it == emap.end())
            {
                id = static_cast<int>(edges_.size());
                emap.emplace(key, id);
                edges_.emplace_back(a, b);
            }
            else id = it->second;
            el_to_edge_[e][le] = id;
        }
    }
}

void Mesh::MarkForRefinement()
{
    bool any = false;
    for (const Cell &c : elems_) any = any || c.geom == Geometry::TRIANGLE || c.geom == Geometry::TETRAHEDRON;
    if (!any) return;
    BuildEdges();
    const int nedges = static_cast<int>(edges_.size());
    std::vector<double> len(nedges);
    for (int i = 0; i < nedges; ++i)
    {
        const double *a = GetVertex(edges_[i].first), *b = GetVertex(edges_[i].second);
        double s = 0.0;
        for (int d = 0; d < sdim_; ++d) s += (a[d] - b[d]) * (a[d] - b[d]);
        len[i] = std::sqrt(s);
    }
    std::vector<int> idx(nedges), rank(nedges);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return len[a] < len[b] || (len[a] == len[b] && a < b); });
    for (int i = 0; i < nedges; ++i) rank[idx[i]] = i;

    for (size_t e = 0; e < elems_.size(); ++e)
    {
        Cell &c = elems_[e];
        const std::vector<int> &ed = el_to_edge_[e];
        if (c.geom == Geometry::TRIANGLE)
        {
            int L = rank[ed[0]], j = 0;
            if (rank[ed[1]] > L) { L = rank[ed[1]]; j = 1; }
            if (rank[ed[2]] > L) { j = 2; }
            const int v[3] = {c.v[0], c.v[1], c.v[2]};
            if (j == 1) { c.v[0] = v[1]; c.v[1] = v[2]; c.v[2] = v[0]; }
            else if (j == 2) { c.v[0] = v[2]; c.v[1] = v[0]; c.v[2] = v[1]; }
        }
        else if (c.geom == Geometry::TETRAHEDRON)
        {
            // local edges of kTetEdges: 0:(0,1) 1:(0,2) 2:(0,3) 3:(1,2) 4:(1,3) 5:(2,3)
            auto r = [&](int a, int b) {
                for (int le = 0; le < 6; ++le)
                {
                    const int x = kTetEdges[le][0], y = kTetEdges[le][1];
                    if ((x == a && y == b) || (x == b && y == a)) return rank[ed[le]];
                }
                return -1;
            };
            int L = r(0, 1), j = 0, l;
            if ((l = r(1, 2)) > L) { L = l; j = 1; }
            if ((l = r(2, 0)) > L) { L = l; j = 2; }
            if ((l = r(0, 3)) > L) { L = l; j = 3; }
            if ((l = r(1, 3)) > L) { L = l; j = 4; }
            if ((l = r(2, 3)) > L) { j = 5; }
            const int v[4] = {c.v[0], c.v[1], c.v[2], c.v[3]};
            switch (j)
            {
                case 1: c.v[0] = v[1]; c.v[1] = v[2]; c.v[2] = v[0]; break;
                case 2: c.v[0] = v[2]; c.v[1] = v[0]; c.v[2] = v[1]; break;
                case 3: c.v[0] = v[3]; c.v[1] = v[0]; c.v[2] = v[2]; c.v[3] = v[1]; break;
                case 4: c.v[0] = v[1]; c.v[1] = v[3]; c.v[2] = v[2]; c.v[3] = v[0]; break;
                case 5: c.v[0] = v[2]; c.v[1] = v[3]; c.v[2] = v[0]; c.v[3] = v[1]; break;
                default: break;
            }
            // second stage: compare the longest edges of the two faces opposite to the marked edge
            auto rv = [&](int a, int b) {
                const int ga = c.v[a], gb = c.v[b];
                for (int le = 0; le < 6; ++le)
                {
                    const int x = v[kTetEdges[le][0]], y = v[kTetEdges[le][1]];
                    if ((x == ga && y == gb) || (x == gb && y == ga)) return rank[ed[le]];
                }
                return -1;
            };
            int i0 = 2, i1 = 1;
            L = rv(0, 2);
            if ((l = rv(0, 3)) > L) { L = l; i0 = 3; }
            if ((l = rv(2, 3)) > L) { i0 = 5; }
            L = rv(1, 2);
            if ((l = rv(1, 3)) > L) { L = l; i1 = 4; }
            if ((l = rv(2, 3)) > L) { i1 = 5; }
            const bool swap = (i0 == 3 && (i1 == 4 || i1 == 5)) || (i0 == 5 && i1 == 4);
            if (swap)
            {
                std::swap(c.v[0], c.v[1]);
                std::swap(c.v[2], c.v[3]);
            }
        }
    }
}

void Mesh::BuildFaces()
{
    faces_.clear();
    el_to_face_.assign(elems_.size(), {});
    if (dim_ == 2)
    {
        BuildEdges();
        faces_.resize(edges_.size());
        for (size_t f = 0; f < edges_.size(); ++f)
        {
            faces_[f].geom = Geometry::SEGMENT;
            faces_[f].nv = 2;
            faces_[f].v[0] = edges_[f].first;
            faces_[f].v[1] = edges_[f].second;
        }
        for (size_t e = 0; e < elems_.size(); ++e)
        {
            el_to_face_[e] = el_to_edge_[e];
            for (size_t lf = 0; lf < el_to_face_[e].size(); ++lf)
            {
                Face &F = faces_[el_to_face_[e][lf]];
                if (F.e1 < 0) { F.e1 = static_cast<int>(e); F.lf1 = static_cast<int>(lf); }
                else
                {
                    MFEM_VERIFY(F.e2 < 0, "non-manifold mesh: edge shared by more than two elements");
                    F.e2 = static_cast<int>(e);
                    F.lf2 = static_cast<int>(lf);
                }
            }
        }
    }
    else
    {
        BuildEdges();
        std::unordered_map<std::array<int, 4>, int, KeyHash> fmap;
        fmap.reserve(elems_.size() * 4);
        for (size_t e = 0; e < elems_.size(); ++e)
        {
            const Cell &c = elems_[e];
            const int nlf = NumLocalFaces(c.geom);
            el_to_face_[e].resize(nlf);
            for (int lf = 0; lf < nlf; ++lf)
            {
                int lv[4], gv[4];
                const int n = LocalFaceVerts(c.geom, lf, lv);
                for (int i = 0; i < n; ++i) gv[i] = c.v[lv[i]];
                const auto key = SortedKey(gv, n);
                auto it = fmap.find(key);
                if (it == fmap.end())
                {
                    Face F;
                    F.geom = n == 3 ? Geometry::TRIANGLE : Geometry::SQUARE;
                    F.nv = n;
                    for (int i = 0; i < n; ++i) F.v[i] = gv[i];
                    F.e1 = static_cast<int>(e);
                    F.lf1 = lf;
                    const int id = static_cast<int>(faces_.size());
                    faces_.push_back(F);
                    fmap.emplace(key, id);
                    el_to_face_[e][lf] = id;
                }
                else
                {
                    Face &F = faces_[it->second];
                    MFEM_VERIFY(F.e2 < 0, "non-manifold mesh: face shared by more than two elements");
                    F.e2 = static_cast<int>(e);
                    F.lf2 = lf;
                    el_to_face_[e][lf] = it->second;
                }
            }
        }
    }
    // boundary element -> face
    std::unordered_map<std::array<int, 4>, int, KeyHash> by_verts;
    by_verts.reserve(faces_.size());
    for (size_t f = 0; f < faces_.size(); ++f)
        if (faces_[f].e2 < 0) by_verts.emplace(SortedKey(faces_[f].v, faces_[f].nv), static_cast<int>(f));
    bdr_face_.assign(bdr_.size(), -1);
    for (size_t b = 0; b < bdr_.size(); ++b)
    {
        auto it = by_verts.find(SortedKey(bdr_[b].v, bdr_[b].nv));
        MFEM_VERIFY(it != by_verts.end(), "boundary element " << b << " does not match a boundary face");
        bdr_face_[b] = it->second;
    }
}

void Mesh::Finalize(bool refine, bool fix_orientation)
{
    if (sdim_ == 0) sdim_ = dim_;
    if (fix_orientation) FixOrientation();
    if (refine) MarkForRefinement();
    BuildFaces();
}

void Mesh::GenerateBoundary(double sx, double sy, double sz)
{
    bdr_.clear();
    for (size_t f = 0; f < faces_.size(); ++f)
    {
        const Face &F = faces_[f];
        if (F.e2 >= 0) continue;
        double c[3] = {0, 0, 0};
        for (int i = 0; i < F.nv; ++i)
            for (int d = 0; d < 3; ++d) c[d] += GetVertex(F.v[i])[d];
        for (int d = 0; d < 3; ++d) c[d] /= F.nv;
        int attr = 0;
        const double tol = 1e-9;
        if (dim_ == 2)
        {
            if (std::fabs(c[1]) < tol * sy) attr = 1;            // bottom
            else if (std::fabs(c[0] - sx) < tol * sx) attr = 2;  // right
            else if (std::fabs(c[1] - sy) < tol * sy) attr = 3;  // top
            else if (std::fabs(c[0]) < tol * sx) attr = 4;       // left
        }
        else
        {
            if (std::fabs(c[2]) < tol * sz) attr = 1;            // bottom  z = 0
            else if (std::fabs(c[1]) < tol * sy) attr = 2;       // front   y = 0
            else if (std::fabs(c[0] - sx) < tol * sx) attr = 3;  // right   x = sx
            else if (std::fabs(c[1] - sy) < tol * sy) attr = 4;  // back    y = sy
            else if (std::fabs(c[0]) < tol * sx) attr = 5;       // left    x = 0
            else if (std::fabs(c[2] - sz) < tol * sz) attr = 6;  // top     z = sz
        }
        MFEM_VERIFY(attr > 0, "GenerateBoundary: face not on the box boundary");
        AddBdr(F.geom, attr, F.v);
    }
    bdr_face_.assign(bdr_.size(), -1);
    size_t b = 0;
    for (size_t f = 0; f < faces_.size(); ++f)
        if (faces_[f].e2 < 0) bdr_face_[b++] = static_cast<int>(f);
}

// ------------------------------------------------------------------ generators
Mesh Mesh::MakeCartesian2D(int nx, int ny, Element::Type type, bool, real_t sx, real_t sy, bool sfc_ordering)
{
    Mesh m;
    m.dim_ = m.sdim_ = 2;
    for (int j = 0; j <= ny; ++j)
    {
        const double cy = (static_cast<double>(j) / ny) * sy;
        for (int i = 0; i <= nx; ++i)
        {
            const double cx = (static_cast<double>(i) / nx) * sx;
            m.AddVertex(cx, cy);
        }
    }
    const int mm = nx + 1;
    if (type == Element::QUADRILATERAL)
    {
        std::vector<int> sfc;
        if (sfc_ordering)
        {
            if (nx >= ny) Gilbert2D(0, 0, nx, 0, 0, ny, sfc);
            else Gilbert2D(0, 0, 0, ny, nx, 0, sfc);
            MFEM_VERIFY(static_cast<int>(sfc.size()) == 2 * nx * ny, "space-filling curve size mismatch");
        }
        else
        {
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) { sfc.push_back(i); sfc.push_back(j); }
        }
        for (int k = 0; k < nx * ny; ++k)
        {
            const int i = sfc[2 * k], j = sfc[2 * k + 1];
            const int v[4] = {i + j * mm, i + 1 + j * mm, i + 1 + (j + 1) * mm, i + (j + 1) * mm};
            m.AddCell(Geometry::SQUARE, 1, v);
        }
    }
    else if (type == Element::TRIANGLE)
    {
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i)
            {
                const int q[4] = {i + j * mm, i + 1 + j * mm, i + 1 + (j + 1) * mm, i + (j + 1) * mm};
                const int t1[3] = {q[0], q[1], q[2]}, t2[3] = {q[2], q[3], q[0]};
                m.AddCell(Geometry::TRIANGLE, 1, t1);
                m.AddCell(Geometry::TRIANGLE, 1, t2);
            }
    }
    else mfem_error("MakeCartesian2D: element type must be TRIANGLE or QUADRILATERAL");
    m.Finalize(true, true);
    m.GenerateBoundary(sx, sy, 1.0);
    return m;
}

Mesh Mesh::MakeCartesian3D(int nx, int ny, int nz, Element::Type type, real_t sx, real_t sy, real_t sz, bool sfc_ordering)
{
    Mesh m;
    m.dim_ = m.sdim_ = 3;
    for (int k = 0; k <= nz; ++k)
    {
        const double cz = (static_cast<double>(k) / nz) * sz;
        for (int j = 0; j <= ny; ++j)
        {
            const double cy = (static_cast<double>(j) / ny) * sy;
            for (int i = 0; i <= nx; ++i)
            {
                const double cx = (static_cast<double>(i) / nx) * sx;
                m.AddVertex(cx, cy, cz);
            }
        }
    }
    auto VTX = [&](int i, int j, int k) { return i + (j + k * (ny + 1)) * (nx + 1); };
    std::vector<int> sfc;
    if (type == Element::HEXAHEDRON && sfc_ordering)
    {
        if (nx >= ny && nx >= nz) Gilbert3D(0, 0, 0, nx, 0, 0, 0, ny, 0, 0, 0, nz, sfc);
        else if (ny >= nx && ny >= nz) Gilbert3D(0, 0, 0, 0, ny, 0, nx, 0, 0, 0, 0, nz, sfc);
        else Gilbert3D(0, 0, 0, 0, 0, nz, nx, 0, 0, 0, ny, 0, sfc);
        MFEM_VERIFY(static_cast<int>(sfc.size()) == 3 * nx * ny * nz, "space-filling curve size mismatch");
    }
    else
    {
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) { sfc.push_back(i); sfc.push_back(j); sfc.push_back(k); }
    }
    static const int hex_to_tet[6][4] = {{0, 1, 2, 6}, {0, 5, 1, 6}, {0, 4, 5, 6}, {0, 2, 3, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}};
    for (int c = 0; c < nx * ny * nz; ++c)
    {
        const int i = sfc[3 * c], j = sfc[3 * c + 1], k = sfc[3 * c + 2];
        const int h[8] = {VTX(i, j, k),         VTX(i + 1, j, k),         VTX(i + 1, j + 1, k),         VTX(i, j + 1, k),
                          VTX(i, j, k + 1),     VTX(i + 1, j, k + 1),     VTX(i + 1, j + 1, k + 1),     VTX(i, j + 1, k + 1)};
        if (type == Element::HEXAHEDRON) m.AddCell(Geometry::CUBE, 1, h);
        else if (type == Element::TETRAHEDRON)
        {
            for (int t = 0; t < 6; ++t)
            {
                const int v[4] = {h[hex_to_tet[t][0]], h[hex_to_tet[t][1]], h[hex_to_tet[t][2]], h[hex_to_tet[t][3]]};
                m.AddCell(Geometry::TETRAHEDRON, 1, v);
            }
        }
        else mfem_error("MakeCartesian3D: element type must be TETRAHEDRON or HEXAHEDRON");
    }
    m.Finalize(true, true);
    m.GenerateBoundary(sx, sy, sz);
    return m;
}

// ------------------------------------------------------------------ uniform refinement
void Mesh::UniformRefinement(int)
{
    const int nv0 = GetNV();
    const int nedges = static_cast<int>(edges_.size());
    MFEM_VERIFY(!el_to_edge_.empty() || elems_.empty(), "UniformRefinement: topology not built");

    std::unordered_map<unsigned long long, int> emap;
    emap.reserve(edges_.size() * 2);
    for (int i = 0; i < nedges; ++i)
    {
        const int a = edges_[i].first, b = edges_[i].second;
        emap.emplace((static_cast<unsigned long long>(std::min(a, b)) << 32) | static_cast<unsigned>(std::max(a, b)), i);
    }
    auto mid = [&](int a, int b) {
        return nv0 + emap.at((static_cast<unsigned long long>(std::min(a, b)) << 32) | static_cast<unsigned>(std::max(a, b)));
    };
    auto average = [&](const int *idx, int n) {
        double c[3];
        for (int d = 0; d < 3; ++d) c[d] = GetVertex(idx[0])[d];
        for (int j = 1; j < n; ++j)
            for (int d = 0; d < 3; ++d) c[d] += GetVertex(idx[j])[d];
        for (int d = 0; d < 3; ++d) c[d] *= (1.0 / n);
        AddVertex(c[0], c[1], c[2]);
    };
    // edge midpoints
    for (int i = 0; i < nedges; ++i)
    {
        const int idx[2] = {edges_[i].first, edges_[i].second};
        average(idx, 2);
    }
    std::vector<Cell> old_elems;
    old_elems.swap(elems_);
    std::vector<Cell> old_bdr;
    old_bdr.swap(bdr_);

    if (dim_ == 2)
    {
        // quad centres follow the edge midpoints
        const int oelem = nv0 + nedges;
        int nquad = 0;
        for (const Cell &c : old_elems)
            if (c.geom == Geometry::SQUARE)
            {
                average(c.v, 4);
                ++nquad;
            }
        int qi = 0;
        for (const Cell &c : old_elems)
        {
            const int *v = c.v;
            if (c.geom == Geometry::TRIANGLE)
            {
                const int a = mid(v[0], v[1]), b = mid(v[1], v[2]), cc = mid(v[2], v[0]);
                const int t0[3] = {v[0], a, cc}, t1[3] = {b, cc, a}, t2[3] = {a, v[1], b}, t3[3] = {cc, b, v[2]};
                AddCell(Geometry::TRIANGLE, c.attr, t0);
                AddCell(Geometry::TRIANGLE, c.attr, t1);
                AddCell(Geometry::TRIANGLE, c.attr, t2);
                AddCell(Geometry::TRIANGLE, c.attr, t3);
            }
            else
            {
                const int a = mid(v[0], v[1]), b = mid(v[1], v[2]), cc = mid(v[2], v[3]), d = mid(v[3], v[0]);
                const int q = oelem + qi++;
                const int q0[4] = {v[0], a, q, d}, q1[4] = {a, v[1], b, q}, q2[4] = {q, b, v[2], cc}, q3[4] = {d, q, cc, v[3]};
                AddCell(Geometry::SQUARE, c.attr, q0);
                AddCell(Geometry::SQUARE, c.attr, q1);
                AddCell(Geometry::SQUARE, c.attr, q2);
                AddCell(Geometry::SQUARE, c.attr, q3);
            }
        }
        for (const Cell &b : old_bdr)
        {
            const int a = mid(b.v[0], b.v[1]);
            const int s0[2] = {b.v[0], a}, s1[2] = {a, b.v[1]};
            AddBdr(Geometry::SEGMENT, b.attr, s0);
            AddBdr(Geometry::SEGMENT, b.attr, s1);
        }
    }
    else
    {
        // quad-face centres, then hex centres
        const int oface = nv0 + nedges;
        std::vector<int> f2qf(faces_.size(), -1);
        int nqf = 0;
        for (size_t f = 0; f < faces_.size(); ++f)
            if (faces_[f].geom == Geometry::SQUARE)
            {
                f2qf[f] = nqf++;
                average(faces_[f].v, 4);
            }
        const int oelem = oface + nqf;
        for (const Cell &c : old_elems)
            if (c.geom == Geometry::CUBE) average(c.v, 8);
        std::unordered_map<std::array<int, 4>, int, KeyHash> fmap;
        for (size_t f = 0; f < faces_.size(); ++f)
            if (faces_[f].geom == Geometry::SQUARE) fmap.emplace(SortedKey(faces_[f].v, 4), oface + f2qf[f]);

        int hi = 0;
        for (size_t e = 0; e < old_elems.size(); ++e)
        {
            const Cell &c = old_elems[e];
            const int *v = c.v;
            if (c.geom == Geometry::TETRAHEDRON)
            {
                int ev[6];
                for (int k = 0; k < 6; ++k) ev[k] = mid(v[kTetEdges[k][0]], v[kTetEdges[k][1]]);
                // inner octahedron: choose the shortest of the three diagonals (ev0-ev5, ev1-ev4, ev2-ev3)
                auto d2 = [&](int a, int b) {
                    const double *p = GetVertex(a), *q = GetVertex(b);
                    return (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]);
                };
                int rt = 0;
                double best = d2(ev[0], ev[5]);
                if (d2(ev[1], ev[4]) < best) { best = d2(ev[1], ev[4]); rt = 1; }
                if (d2(ev[2], ev[3]) < best) { rt = 2; }
                static const int mv_all[3][4][4] = {{{0, 5, 1, 2}, {0, 5, 2, 4}, {0, 5, 4, 3}, {0, 5, 3, 1}},
                                                    {{1, 4, 3, 5}, {1, 4, 5, 2}, {1, 4, 2, 0}, {1, 4, 0, 3}},
                                                    {{2, 3, 0, 1}, {2, 3, 1, 5}, {2, 3, 5, 4}, {2, 3, 4, 0}}};
                const int t0[4] = {v[0], ev[0], ev[1], ev[2]}, t1[4] = {ev[0], v[1], ev[3], ev[4]};
                const int t2[4] = {ev[1], ev[3], v[2], ev[5]}, t3[4] = {ev[2], ev[4], ev[5], v[3]};
                AddCell(Geometry::TETRAHEDRON, c.attr, t0);
                AddCell(Geometry::TETRAHEDRON, c.attr, t1);
                AddCell(Geometry::TETRAHEDRON, c.attr, t2);
                AddCell(Geometry::TETRAHEDRON, c.attr, t3);
                for (int k = 0; k < 4; ++k)
                {
                    const int t[4] = {ev[mv_all[rt][k][0]], ev[mv_all[rt][k][1]], ev[mv_all[rt][k][2]], ev[mv_all[rt][k][3]]};
                    AddCell(Geometry::TETRAHEDRON, c.attr, t);
                }
            }
            else
            {
                int ev[12], qf[6];
                for (int k = 0; k < 12; ++k) ev[k] = mid(v[kHexEdges[k][0]], v[kHexEdges[k][1]]);
                for (int k = 0; k < 6; ++k)
                {
                    int gv[4];
                    for (int i = 0; i < 4; ++i) gv[i] = v[kHexFaces[k][i]];
                    qf[k] = fmap.at(SortedKey(gv, 4));
                }
                const int he = oelem + hi++;
                const int ch[8][8] = {{v[0], ev[0], qf[0], ev[3], ev[8], qf[1], he, qf[4]},
                                      {ev[0], v[1], ev[1], qf[0], qf[1], ev[9], qf[2], he},
                                      {qf[0], ev[1], v[2], ev[2], he, qf[2], ev[10], qf[3]},
                                      {ev[3], qf[0], ev[2], v[3], qf[4], he, qf[3], ev[11]},
                                      {ev[8], qf[1], he, qf[4], v[4], ev[4], qf[5], ev[7]},
                                      {qf[1], ev[9], qf[2], he, ev[4], v[5], ev[5], qf[5]},
                                      {he, qf[2], ev[10], qf[3], qf[5], ev[5], v[6], ev[6]},
                                      {qf[4], he, qf[3], ev[11], ev[7], qf[5], ev[6], v[7]}};
                for (int k = 0; k < 8; ++k) AddCell(Geometry::CUBE, c.attr, ch[k]);
            }
        }
        for (const Cell &b : old_bdr)
        {
            const int *v = b.v;
            if (b.geom == Geometry::TRIANGLE)
            {
                const int a = mid(v[0], v[1]), bb = mid(v[1], v[2]), cc = mid(v[2], v[0]);
                const int t0[3] = {v[0], a, cc}, t1[3] = {bb, cc, a}, t2[3] = {a, v[1], bb}, t3[3] = {cc, bb, v[2]};
                AddBdr(Geometry::TRIANGLE, b.attr, t0);
                AddBdr(Geometry::TRIANGLE, b.attr, t1);
                AddBdr(Geometry::TRIANGLE, b.attr, t2);
                AddBdr(Geometry::TRIANGLE, b.attr, t3);
            }
            else
            {
                const int a = mid(v[0], v[1]), bb = mid(v[1], v[2]), cc = mid(v[2], v[3]), d = mid(v[3], v[0]);
                const int q = fmap.at(SortedKey(v, 4));
                const int q0[4] = {v[0], a, q, d}, q1[4] = {a, v[1], bb, q}, q2[4] = {q, bb, v[2], cc}, q3[4] = {d, q, cc, v[3]};
                AddBdr(Geometry::SQUARE, b.attr, q0);
                AddBdr(Geometry::SQUARE, b.attr, q1);
                AddBdr(Geometry::SQUARE, b.attr, q2);
                AddBdr(Geometry::SQUARE, b.attr, q3);
            }
        }
    }
    // children inherit a valid orientation on affine parents; fix the rare negative inner tetrahedron
    FixOrientation();
    BuildFaces();
}

// ------------------------------------------------------------------ queries
void Mesh::GetElementVertices(int e, Array<int> &v) const
{
    const Cell &c = elems_[e];
    v.SetSize(c.nv);
    for (int i = 0; i < c.nv; ++i) v[i] = c.v[i];
}

void Mesh::GetFaceVertices(int f, Array<int> &v) const
{
    const Face &F = faces_[f];
    v.SetSize(F.nv);
    for (int i = 0; i < F.nv; ++i) v[i] = F.v[i];
}

void Mesh::GetFaceElements(int f, int *e1, int *e2) const
{
    *e1 = faces_[f].e1;
    *e2 = faces_[f].e2;
}

void Mesh::GetElementTransformation(int e, ElementTransformation *tr) const
{
    const Cell &c = elems_[e];
    const double *pv[8];
    for (int i = 0; i < c.nv; ++i) pv[i] = GetVertex(c.v[i]);
    tr->Setup(static_cast<Geometry::Type>(c.geom), sdim_, pv);
    tr->Attribute = c.attr;
    tr->ElementNo = e;
}

ElementTransformation *Mesh::GetElementTransformation(int e)
{
    GetElementTransformation(e, &tr1_);
    return &tr1_;
}

FaceElementTransformations *Mesh::GetFaceElementTransformations(int f, int)
{
    const Face &F = faces_[f];
    ftr_ = FaceElementTransformations();
    ftr_.Elem1No = F.e1;
    ftr_.Elem2No = F.e2;
    GetElementTransformation(F.e1, &tr1_);
    ftr_.Elem1 = &tr1_;
    ftr_.lf1 = F.lf1;
    ftr_.egeom1 = static_cast<Geometry::Type>(elems_[F.e1].geom);
    for (int i = 0; i < 8; ++i) ftr_.ev1[i] = elems_[F.e1].v[i];
    if (F.e2 >= 0)
    {
        GetElementTransformation(F.e2, &tr2_);
        ftr_.Elem2 = &tr2_;
        ftr_.lf2 = F.lf2;
        ftr_.egeom2 = static_cast<Geometry::Type>(elems_[F.e2].geom);
        for (int i = 0; i < 8; ++i) ftr_.ev2[i] = elems_[F.e2].v[i];
    }
    else ftr_.Elem2 = nullptr;
    const double *pv[4];
    for (int i = 0; i < F.nv; ++i)
    {
        pv[i] = GetVertex(F.v[i]);
        ftr_.fv[i] = F.v[i];
    }
    trf_.Setup(static_cast<Geometry::Type>(F.geom), sdim_, pv);
    trf_.ElementNo = f;
    ftr_.Face = &trf_;
    return &ftr_;
}

}  // namespace mfem
