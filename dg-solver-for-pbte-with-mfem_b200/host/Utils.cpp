// Unit face normals; operation order as in the reference's src/Utils.cpp:262-354:
//   2-D: (dy, -dx) of the first two face vertices; 3-D: (v1-v0) x (v2-v0) of the first three;
//   divided by Vector::Norml2; flipped when normal . (face centroid - element vertex centroid) < 0.
#include "Utils.hpp"

#include "ElementIntegrator.hpp"

#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>

namespace utils
{
namespace
{
// opens output/log/<name> (directories created on the way); reports a failure the way the reference does
bool OpenLog(const char *name, std::ofstream &ofs, std::filesystem::path &path)
{
    path = std::filesystem::path("output/log") / name;
    std::filesystem::create_directories(path.parent_path());
    ofs.open(path);
    if (!ofs) std::cerr << "Failed to open " << path << " for writing.\n";
    return static_cast<bool>(ofs);
}

// "elem e" header line, then the ndof values of column e separated by blanks
void WriteColumns(std::ostream &os, const mfem::DenseMatrix &m)
{
    for (int e = 0; e < m.Width(); ++e)
    {
        os << "elem " << e << "\n";
        for (int i = 0; i < m.Height(); ++i) os << (i ? " " : "") << m(i, e);
        os << "\n";
    }
}

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

// ---- text dumps (formats of src/Utils.cpp:17-260; the shipped output/log files are the fixtures) -----------------------------
void WriteVector(std::ostream &os, const mfem::Vector &v, const std::string &label)
{
    os << label << " [size=" << v.Size() << "]:";
    for (int i = 0; i < v.Size(); ++i) os << " " << v[i];
    os << "\n";
}

void WriteMatrix(std::ostream &os, const mfem::DenseMatrix &m, const std::string &label)
{
    os << label << " [shape=" << m.Height() << "x" << m.Width() << "]\n";
    for (int i = 0; i < m.Height(); ++i)
    {
        os << "  ";
        for (int j = 0; j < m.Width(); ++j) os << (j ? " " : "") << m(i, j);
        os << "\n";
    }
}

bool DumpElementIntegrals(const std::vector<pbte::ElementIntegralData> &element_data, bool is_root)
{
    if (!is_root) return true;
    std::ofstream ofs;
    std::filesystem::path path;
    if (!OpenLog("integrals_all.txt", ofs, path)) return false;
    ofs << "DG integral dump (local rank block)\nrank: 0/1\nelements: " << element_data.size() << "\n";
    for (size_t e = 0; e < element_data.size(); ++e)
    {
        const pbte::ElementIntegralData &ed = element_data[e];
        ofs << "\n=== Element " << e << " (rank 0) ===\n";
        WriteVector(ofs, ed.basis_integrals, "basis_integrals");
        WriteMatrix(ofs, ed.mass_matrix, "mass_matrix");
        for (size_t d = 0; d < ed.stiffness_matrices.size(); ++d) WriteMatrix(ofs, ed.stiffness_matrices[d], "stiffness_matrix_dim" + std::to_string(d));
        for (size_t f = 0; f < ed.face_mass_matrices.size(); ++f)
        {
            WriteMatrix(ofs, ed.face_mass_matrices[f], "face_mass_matrix[" + std::to_string(f) + "]");
            WriteVector(ofs, ed.face_integrals[f], "face_integral[" + std::to_string(f) + "]");
        }
        for (size_t k = 0; k < ed.face_couplings.size(); ++k)
        {
            const pbte::ElementIntegralData::FaceCoupling &fc = ed.face_couplings[k];
            ofs << "face_coupling[" << k << "]: face_id=" << fc.face_id << ", neighbor=" << fc.neighbor_elem << ", attr=" << fc.boundary_attr
                << ", shared=" << (fc.is_shared ? 1 : 0) << "\n";
            if (fc.neighbor_elem >= 0 || fc.is_shared) WriteMatrix(ofs, fc.coupling, "  coupling");
            else WriteVector(ofs, fc.isothermal_rhs, "  isothermal_rhs");
        }
    }
    std::cout << "Integral dump written to: " << path << std::endl;
    return true;
}

bool DumpCoefficients(const std::vector<std::vector<std::vector<mfem::DenseMatrix>>> &coeff, const pbte::AngleQuadrature &quad, bool is_root)
{
    if (!is_root) return true;
    // ~13 characters per value: refuse what would be an accident at benchmark sizes
    double values = 0.0;
    for (const auto &per_dir : coeff)
        for (const auto &per_branch : per_dir)
            for (const mfem::DenseMatrix &m : per_branch) values += static_cast<double>(m.Height()) * m.Width();
    double limit_mb = 4096.0;
    if (const char *ev = std::getenv("PBTE_DUMP_LIMIT_MB")) limit_mb = std::atof(ev);
    if (values * 13.0 > limit_mb * 1048576.0)
    {
        std::cerr << "DumpCoefficients: about " << values * 13.0 / 1048576.0 << " MB of text exceed PBTE_DUMP_LIMIT_MB=" << limit_mb
                  << "; not written.\n";
        return false;
    }
    std::ofstream ofs;
    std::filesystem::path path;
    if (!OpenLog("coeff_all.txt", ofs, path)) return false;
    const auto &dirs = quad.Directions();
    for (size_t d = 0; d < coeff.size(); ++d)
        for (size_t b = 0; b < coeff[d].size(); ++b)
            for (size_t s = 0; s < coeff[d][b].size(); ++s)
            {
                const mfem::DenseMatrix &m = coeff[d][b][s];
                ofs << "# dir " << d << " branch " << b << " spec " << s << "\n# ndof " << m.Height() << " ne " << m.Width() << "\n# direction: ";
                for (size_t k = 0; k < dirs[d].direction.size(); ++k) ofs << (k ? " " : "") << dirs[d].direction[k];
                ofs << " weight " << dirs[d].weight << "\n";
                WriteColumns(ofs, m);
                ofs << "\n";
            }
    ofs << std::flush;
    std::cout << "Coefficient blocks written to: " << path << std::endl;
    return true;
}

bool DumpTemperature(const mfem::DenseMatrix &Tc, bool is_root)
{
    if (!is_root) return true;
    std::ofstream ofs;
    std::filesystem::path path;
    if (!OpenLog("Tc_all.txt", ofs, path)) return false;
    ofs << "# Tc matrix\n# ndof " << Tc.Height() << " ne " << Tc.Width() << "\n";
    WriteColumns(ofs, Tc);
    ofs << std::flush;
    std::cout << "Macroscopic temperature Tc written to: " << path << std::endl;
    return true;
}
}  // namespace utils
