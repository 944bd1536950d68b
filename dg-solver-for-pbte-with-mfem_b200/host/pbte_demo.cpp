// pbte_demo: command-line driver with the reference's options (-m -c -o -r, src/PhononBTE.cpp:36-53) and
// call sequence (:84-411), running the sweep on the GPU through the kept class API.
#include "AngularQuadrature.hpp"
#include "AngularSweepOrder.hpp"
#include "ElementIntegrator.hpp"
#include "MacroscopicQuantities.hpp"
#include "PBTESolver.hpp"
#include "PhononProperties.hpp"
#include "SpatialMesh.hpp"
#include "Utils.hpp"

#include <yaml-cpp/yaml.h>

#include <chrono>
#include <iostream>
#include <sstream>

int main(int argc, char *argv[])
{
    std::string mesh_spec, config_path = "config/config.yaml", material_path = "config/si.yaml";
    int order = 1, refine = 0;
    bool write_fields = true;
    mfem::OptionsParser args(argc, argv);
    args.AddOption(&mesh_spec, "-m", "--mesh", "Mesh file or builtin name; default: taken from the config.");
    args.AddOption(&config_path, "-c", "--config", "Path to config YAML (defaults to config/config.yaml).");
    args.AddOption(&material_path, "-s", "--material", "Path to the material YAML (defaults to config/si.yaml).");
    args.AddOption(&order, "-o", "--order", "Polynomial order for DG/L2 finite element space.");
    args.AddOption(&refine, "-r", "--refine", "Uniform refinement levels (applied before building space).");
    args.AddOption(&write_fields, "-w", "--write-fields", "-nw", "--no-write-fields", "Write coefficient/field dumps after the solve.");
    args.Parse();
    if (!args.Good())
    {
        args.PrintUsage(std::cout);
        return 1;
    }
    args.PrintOptions(std::cout);

    pbte::SpatialMesh spatial;
    pbte::PhononMaterial material;
    try
    {
        if (!mesh_spec.empty()) spatial.LoadMesh(mesh_spec);
        else spatial.LoadMeshFromConfig(config_path);
        material = pbte::PhononProperties::LoadMaterial(material_path);
        spatial.ScaleCoordinates(material.ref_len);
        spatial.UniformRefine(refine);
        spatial.BuildDGSpace(order);
    }
    catch (const std::exception &ex)
    {
        std::cerr << "Error constructing mesh/FE space: " << ex.what() << "\n";
        return 1;
    }
    pbte::AngleDiscretizationOptions angle_opts;
    try
    {
        angle_opts = pbte::AngleQuadrature::LoadOptionsFromConfig(config_path);
    }
    catch (const std::exception &ex)
    {
        std::cerr << "Warning: failed to read angular options from config: " << ex.what() << ". Using built-in defaults.\n";
    }
    pbte::AngleQuadrature quad;
    try
    {
        quad = pbte::AngleQuadrature::Build(angle_opts);
    }
    catch (const std::exception &ex)
    {
        std::cerr << "Error constructing angular quadrature: " << ex.what() << "\n";
        return 1;
    }
    std::cout << "Angular quadrature built: dim=" << angle_opts.dimension << ", polar pts=" << quad.PolarAngles().size()
              << ", azimuth pts=" << quad.AzimuthAngles().size() << ", directions=" << quad.Directions().size()
              << ", total weight=" << quad.TotalWeight() << std::endl;
    std::ostringstream tag;
    tag << "dim" << angle_opts.dimension << "_np" << angle_opts.polar_points << "_" << pbte::AngleQuadrature::SchemeName(angle_opts.polar_scheme)
        << "_na" << angle_opts.azimuth_points << "_" << pbte::AngleQuadrature::SchemeName(angle_opts.azimuth_scheme) << ".txt";
    quad.WriteToFile("output/log/angles_" + tag.str());

    pbte::DGElementIntegrator integrator(spatial.FESpace());
    const auto element_data = integrator.AssembleAll();
    const pbte::PhononProperties props = pbte::PhononProperties::Build(material);
    props.WriteToFile("output/log/phonon_properties.txt");

    pbte::MacroscopicQuantities macro(spatial.FESpace(), props, quad);
    macro.Reset();
    double tol = 1e-7;
    int max_iter = 1000;
    try
    {
        const YAML::Node cfg = YAML::LoadFile(config_path);
        if (cfg["numerical"])
        {
            if (cfg["numerical"]["tolerance"]) tol = cfg["numerical"]["tolerance"].as<double>();
            if (cfg["numerical"]["max_iter"]) max_iter = cfg["numerical"]["max_iter"].as<int>();
        }
    }
    catch (const std::exception &ex)
    {
        std::cerr << "Warning: failed to read numerical config (using defaults): " << ex.what() << std::endl;
    }

    try
    {
        const auto t0 = std::chrono::steady_clock::now();
        pbte::AngularSweepOrder sweep = pbte::AngularSweepOrder::Build(spatial.Mesh(), quad);
        sweep.WriteToFile(quad, spatial.Mesh(), "output/log/sweep_" + tag.str());
        pbte::PBTESolver solver(spatial.Mesh(), spatial.FESpace(), quad, sweep, element_data, props, spatial.IsothermalBoundaryTemps(),
                                pbte::CachePolicy::FullLU, tol, max_iter);
        auto coeff = solver.CreateInitialCoefficients();
        const auto t1 = std::chrono::steady_clock::now();
        const double res = solver.Solve(coeff, macro);
        const auto t2 = std::chrono::steady_clock::now();
        const double sweep_s = solver.LastSolveSweepMs() * 1e-3;
        std::cout << "[GPU] final residual = " << res << " after " << solver.LastSolveIterations() << " iterations; setup "
                  << std::chrono::duration<double>(t1 - t0).count() << " s, solve " << std::chrono::duration<double>(t2 - t1).count()
                  << " s (device " << sweep_s << " s), " << solver.UpdatesPerSweep() * solver.LastSolveIterations() / std::max(sweep_s, 1e-12)
                  << " DOF*angle*band updates/s" << std::endl;
        if (write_fields)
        {
            // the dumps the reference's driver writes after the solve (src/PhononBTE.cpp:399-411)
            utils::DumpCoefficients(coeff, quad, true);
            utils::DumpTemperature(macro.Tc(), true);
            utils::DumpElementIntegrals(element_data, true);
            macro.WriteParaView("pbte_fields");
            if (spatial.Mesh().SpaceDimension() == 2) macro.Write2DSliceTemperature("output/2D/results/T_slice.txt", 100, 100);
        }
    }
    catch (const std::exception &ex)
    {
        std::cerr << "Error in solve: " << ex.what() << "\n";
        return 1;
    }
    return 0;
}
