// Unit face normals; operation order as in the reference's src/Utils.cpp:262-354:
//   2-D: (dy, -dx) of the first two face vertices; 3-D: (v1-v0) x (v2-v0) of the first three;
//   divided by Vector::Norml2; flipped when normal . (face centroid - element vertex centroid) < 0.
#include "Utils.hpp"

namespace utils
{
namespace
{
// mean of the listed vertices, accumulated in list order then scaled by 1/count (Vector::operator/=)
void Centroid(const mfem::Mesh &mesh, const mfem::Array<int> &ids, mfem::Vector &c)
{
    c = 0.0;
    for (int v : ids)
    {
        const double *x = mesh.GetVertex(v);
        for (int d = 0; d < c.Size(); ++d) c[d] += x[d];
    }
    if (ids.Size() > 0) c /= static_cast<double>(ids.Size());
}
}  // namespace

mfem::Vector ComputeFaceNormal(const mfem::Mesh &mesh, int face_id)
{
    mfem::Array<int> fv;
    mesh.GetFaceVertices(face_id, fv);
    const int sdim = mesh.SpaceDimension();
    mfem::Vector n(sdim);
    n = 0.0;
    if (sdim == 2 && fv.Size() >= 2)
    {
        const double *a = mesh.GetVertex(fv[0]), *b = mesh.GetVertex(fv[1]);
        n[0] = b[1] - a[1];
        n[1] = -(b[0] - a[0]);
    }
    else if (sdim == 3 && fv.Size() >= 3)
    {
        const double *a = mesh.GetVertex(fv[0]), *b = mesh.GetVertex(fv[1]), *c = mesh.GetVertex(fv[2]);
        const double p[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        const double q[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        n[0] = p[1] * q[2] - p[2] * q[1];
        n[1] = p[2] * q[0] - p[0] * q[2];
        n[2] = p[0] * q[1] - p[1] * q[0];
    }
    const double len = n.Norml2();
    if (len > 0.0) n /= len;
    return n;
}

mfem::Vector ComputeOutwardFaceNormal(const mfem::Mesh &mesh, int face_id, int elem_id)
{
    mfem::Vector n = ComputeFaceNormal(mesh, face_id);
    const int sdim = mesh.SpaceDimension();
    mfem::Vector ec(sdim), fc(sdim);
    mfem::Array<int> ids;
    mesh.GetElementVertices(elem_id, ids);
    Centroid(mesh, ids, ec);
    mesh.GetFaceVertices(face_id, ids);
    Centroid(mesh, ids, fc);
    fc -= ec;
    if (n * fc < 0.0) n *= -1.0;
    return n;
}
}  // namespace utils
