// Behaviour follows the reference's src/SpatialMesh.cpp (LoadMesh :66-81, UniformRefine :83-103,
// config scanner :105-245, BuildDGSpace :247-268, built-ins :305-340, summary log :386-534).
#include "SpatialMesh.hpp"

#include <cctype>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace fs = std::filesystem;

namespace pbte
{
namespace
{
std::string Strip(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}
bool StartsWith(const std::string &s, const char *p) { return s.rfind(p, 0) == 0; }
std::string After(const std::string &s, const char *key) { return Strip(s.substr(std::string(key).size())); }
}  // namespace

void SpatialMesh::LoadMesh(const std::string &spec)
{
    source_ = spec;
    if (spec.empty()) throw std::invalid_argument("Mesh path or builtin name cannot be empty.");
    if (fs::exists(spec))
    {
        mesh_ = std::make_unique<mfem::Mesh>(spec.c_str(), 1, 1);
        return;
    }
    using mfem::Element;
    if (spec == "unit-square" || spec == "unit-square-tri")
        mesh_ = std::make_unique<mfem::Mesh>(mfem::Mesh::MakeCartesian2D(8, 8, Element::TRIANGLE, true, 1.0, 1.0));
    else if (spec == "unit-square-quad")
        mesh_ = std::make_unique<mfem::Mesh>(mfem::Mesh::MakeCartesian2D(8, 8, Element::QUADRILATERAL, true, 1.0, 1.0));
    else if (spec == "unit-cube" || spec == "unit-cube-tet")
        mesh_ = std::make_unique<mfem::Mesh>(mfem::Mesh::MakeCartesian3D(4, 4, 4, Element::TETRAHEDRON, 1.0, 1.0, 1.0));
    else if (spec == "unit-cube-hex")
        mesh_ = std::make_unique<mfem::Mesh>(mfem::Mesh::MakeCartesian3D(4, 4, 4, Element::HEXAHEDRON, 1.0, 1.0, 1.0));
    else throw std::invalid_argument("Unrecognized built-in mesh name: " + spec);
}

void SpatialMesh::UniformRefine(int levels)
{
    if (!mesh_) return;
    for (int i = 0; i < levels; ++i) mesh_->UniformRefinement();
}

void SpatialMesh::ScaleCoordinates(double factor)
{
    if (!mesh_ || factor == 1.0) return;
    if (mfem::GridFunction *nodes = mesh_->GetNodes())
    {
        (*nodes) *= factor;
        return;
    }
    const int sdim = mesh_->SpaceDimension();
    for (int i = 0; i < mesh_->GetNV(); ++i)
    {
        mfem::real_t *v = mesh_->GetVertex(i);
        for (int d = 0; d < sdim; ++d) v[d] *= factor;
    }
}

void SpatialMesh::LoadMeshFromConfig(const std::string &config_path)
{
    std::ifstream in(config_path);
    if (!in) throw std::runtime_error("Failed to open config file: " + config_path);
    std::vector<std::string> lines;
    for (std::string l; std::getline(in, l);)
    {
        const size_t hash = l.find('#');
        if (hash != std::string::npos) l.erase(hash);
        l = Strip(l);
        if (!l.empty()) lines.push_back(l);
    }
    std::string path;
    bool in_mesh = false;
    for (const std::string &l : lines)
    {
        if (StartsWith(l, "mesh:")) { in_mesh = true; continue; }
        if (in_mesh && StartsWith(l, "path:")) { path = After(l, "path:"); break; }
        if (!in_mesh && StartsWith(l, "mesh_path:")) { path = After(l, "mesh_path:"); break; }
    }
    if (path.empty()) throw std::runtime_error("No mesh path found in config: " + config_path);
    if (path.size() >= 2 && ((path.front() == '\'' && path.back() == '\'') || (path.front() == '"' && path.back() == '"')))
        path = path.substr(1, path.size() - 2);
    LoadMesh(path);

    // boundary_conditions: a list of "- attr: A" / "temperature: T" pairs
    isothermal_bc_.clear();
    bool in_bc = false;
    int attr = -1;
    auto take_temperature = [&](const std::string &v) {
        try { isothermal_bc_[attr] = std::stod(v); }
        catch (const std::exception &) { std::cerr << "Warning: failed to parse temperature value: " << v << std::endl; }
    };
    for (const std::string &l : lines)
    {
        if (StartsWith(l, "boundary_conditions:")) { in_bc = true; continue; }
        if (!in_bc) continue;
        std::string item = l;
        const bool dash = StartsWith(l, "-");
        if (dash)
        {
            attr = -1;
            item = Strip(l.substr(1));
        }
        if (StartsWith(item, "attr:")) attr = std::stoi(After(item, "attr:"));
        else if (StartsWith(item, "temperature:") && attr >= 0)
        {
            take_temperature(After(item, "temperature:"));
            if (!dash) attr = -1;
        }
    }
}

void SpatialMesh::BuildDGSpace(int order, mfem::Ordering::Type ordering, const std::string &log_path)
{
    if (!mesh_) throw std::runtime_error("Mesh must be loaded before building FE space.");
    if (order < 0) throw std::invalid_argument("Polynomial order must be non-negative.");
    order_ = order;
    fec_ = std::make_unique<mfem::L2_FECollection>(order, mesh_->Dimension());
    fes_ = std::make_unique<mfem::FiniteElementSpace>(mesh_.get(), fec_.get(), 1, ordering);

    std::string path = log_path;
    if (path.empty())
    {
        std::string base = fs::path(source_).has_stem() ? fs::path(source_).stem().string() : source_;
        if (base.empty()) base = "mesh";
        for (char &c : base)
            if (c == '/' || c == '\\' || c == ' ') c = '_';
        std::ostringstream os;
        os << "output/log/mesh_" << base << "_p" << order_ << "_dim" << mesh_->Dimension() << ".txt";
        path = os.str();
    }
    if (path == "-") return;  // extension: "-" suppresses the (O(ne) text) summary
    try
    {
        fs::create_directories(fs::path(path).parent_path());
        std::ofstream os(path);
        if (os)
        {
            os << Summary();
            std::cout << "Mesh summary written to: " << fs::path(path) << std::endl;
        }
        else std::cerr << "Failed to open log file for writing: " << fs::path(path) << std::endl;
    }
    catch (const std::exception &ex)
    {
        std::cerr << "Failed to write log: " << ex.what() << std::endl;
    }
}

std::string SpatialMesh::Summary() const
{
    const mfem::Mesh &m = *mesh_;
    std::ostringstream r;
    const int ne = m.GetNE(), sdim = m.SpaceDimension(), nf = m.GetNumFaces();
    r << "Mesh and DG space summary\n"
      << "  mesh source          : " << source_ << "\n"
      << "  dimension            : " << m.Dimension() << "\n"
      << "  element count        : " << ne << "\n"
      << "  boundary elem count  : " << m.GetNBE() << "\n"
      << "  vertex count         : " << m.GetNV() << "\n"
      << "  element geometry     : " << (ne > 0 ? mfem::Geometry::Name[m.GetElementBaseGeometry(0)] : "unknown") << "\n"
      << "  DG polynomial order  : " << order_ << "\n"
      << "  FE space ndofs       : " << fes_->GetNDofs() << "\n"
      << "  FE space vdim        : " << fes_->GetVDim() << "\n"
      << "  ordering             : " << (fes_->GetOrdering() == mfem::Ordering::byNODES ? "byNODES" : "byVDIM") << "\n";
    std::vector<int> attr(nf, 0);
    for (int be = 0; be < m.GetNBE(); ++be)
    {
        const int f = m.GetBdrElementFaceIndex(be);
        if (f >= 0 && f < nf) attr[f] = m.GetBdrAttribute(be);
    }
    std::vector<std::vector<int>> faces_of(ne);
    for (int f = 0; f < nf; ++f)
    {
        int e1 = -1, e2 = -1;
        m.GetFaceElements(f, &e1, &e2);
        if (e1 >= 0) faces_of[e1].push_back(f);
        if (e2 >= 0) faces_of[e2].push_back(f);
    }
    r << "Element details (vertices and faces):\n";
    mfem::Array<int> verts;
    for (int e = 0; e < ne; ++e)
    {
        r << "  elem " << e << "\n    vertices:";
        m.GetElementVertices(e, verts);
        for (int i = 0; i < verts.Size(); ++i)
        {
            const double *x = m.GetVertex(verts[i]);
            r << " v" << verts[i] << "(";
            for (int d = 0; d < sdim; ++d) r << x[d] << (d + 1 < sdim ? "," : "");
            r << ")";
        }
        r << "\n";
        for (int f : faces_of[e])
        {
            int e1 = -1, e2 = -1;
            m.GetFaceElements(f, &e1, &e2);
            const int nb = e1 == e ? e2 : e1;
            r << "    face " << f << " neigh=" << nb;
            if (nb < 0) r << " boundary attr=" << attr[f];
            else r << " interior";
            r << "\n";
        }
    }
    return r.str();
}
}  // namespace pbte
