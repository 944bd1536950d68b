// mfem_lite: command-line parser and the text stand-in for the ParaView writer.
#include "mfem.hpp"

#include <filesystem>
#include <fstream>
#include <iomanip>

namespace mfem
{
void ParaViewDataCollection::Save()
{
    namespace fs = std::filesystem;
    fs::path dir = fs::path(prefix_.empty() ? "." : prefix_) / name_;
    fs::create_directories(dir);
    std::ofstream ofs(dir / "vertex_values.txt");
    MFEM_VERIFY(static_cast<bool>(ofs), "cannot open " << (dir / "vertex_values.txt").string());
    ofs << std::setprecision(17);
    static const double ref_tri[3][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    static const double ref_sq[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
    static const double ref_tet[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static const double ref_cube[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    for (const auto &f : fields_)
    {
        const GridFunction &gf = *f.second;
        const int vdim = gf.FESpace()->GetVDim();
        ofs << "field " << f.first << " vdim " << vdim << " elements " << mesh_->GetNE() << "\n";
        for (int e = 0; e < mesh_->GetNE(); ++e)
        {
            const Geometry::Type g = mesh_->GetElementBaseGeometry(e);
            const double(*ref)[3] = g == Geometry::TRIANGLE ? ref_tri : g == Geometry::SQUARE ? ref_sq : g == Geometry::TETRAHEDRON ? ref_tet : ref_cube;
            for (int k = 0; k < Geometry::NumVerts[g]; ++k)
            {
                IntegrationPoint ip;
                ip.Set3(ref[k][0], ref[k][1], ref[k][2]);
                ofs << e << " " << k;
                for (int d = 1; d <= vdim; ++d) ofs << " " << gf.GetValue(e, ip, d);
                ofs << "\n";
            }
        }
    }
}

// ------------------------------------------------------------------ OptionsParser
void OptionsParser::AddOption(std::string *var, const char *s, const char *l, const char *descr, bool)
{
    opts_.push_back({0, var, s, l, "", "", descr});
}
void OptionsParser::AddOption(const char **var, const char *s, const char *l, const char *descr, bool)
{
    opts_.push_back({1, var, s, l, "", "", descr});
}
void OptionsParser::AddOption(int *var, const char *s, const char *l, const char *descr, bool)
{
    opts_.push_back({2, var, s, l, "", "", descr});
}
void OptionsParser::AddOption(double *var, const char *s, const char *l, const char *descr, bool)
{
    opts_.push_back({3, var, s, l, "", "", descr});
}
void OptionsParser::AddOption(bool *var, const char *en_s, const char *en_l, const char *dis_s, const char *dis_l,
                              const char *descr, bool)
{
    opts_.push_back({4, var, en_s, en_l, dis_s, dis_l, descr});
}

void OptionsParser::Parse()
{
    good_ = true;
    for (int i = 1; i < argc_; ++i)
    {
        const std::string a = argv_[i];
        bool matched = false;
        for (Opt &o : opts_)
        {
            if (o.kind == 4)
            {
                if (a == o.s || a == o.l) { *static_cast<bool *>(o.var) = true; matched = true; break; }
                if (a == o.ds || a == o.dl) { *static_cast<bool *>(o.var) = false; matched = true; break; }
                continue;
            }
            if (a != o.s && a != o.l) continue;
            if (i + 1 >= argc_) { good_ = false; return; }
            const char *val = argv_[++i];
            try
            {
                switch (o.kind)
                {
                    case 0: *static_cast<std::string *>(o.var) = val; break;
                    case 1: *static_cast<const char **>(o.var) = val; break;
                    case 2: *static_cast<int *>(o.var) = std::stoi(val); break;
                    case 3: *static_cast<double *>(o.var) = std::stod(val); break;
                    default: break;
                }
            }
            catch (const std::exception &) { good_ = false; return; }
            matched = true;
            break;
        }
        if (!matched) { good_ = false; return; }
    }
}

void OptionsParser::PrintUsage(std::ostream &os) const
{
    os << "Usage: " << (argc_ > 0 ? argv_[0] : "prog") << " [options] ...\nOptions:\n";
    for (const Opt &o : opts_)
    {
        os << "   " << o.s << ", " << o.l;
        if (o.kind == 4) os << ", " << o.ds << ", " << o.dl;
        os << "\n\t" << o.descr << "\n";
    }
}

void OptionsParser::PrintOptions(std::ostream &os) const
{
    os << "Options used:\n";
    for (const Opt &o : opts_)
    {
        os << "   ";
        switch (o.kind)
        {
            case 0: os << o.l << " " << *static_cast<std::string *>(o.var); break;
            case 1: os << o.l << " " << (*static_cast<const char **>(o.var) ? *static_cast<const char **>(o.var) : ""); break;
            case 2: os << o.l << " " << *static_cast<int *>(o.var); break;
            case 3: os << o.l << " " << *static_cast<double *>(o.var); break;
            case 4: os << (*static_cast<bool *>(o.var) ? o.l : o.dl); break;
            default: break;
        }
        os << "\n";
    }
}

}  // namespace mfem
