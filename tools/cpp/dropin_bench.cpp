// Times the PLUGIN call: Opm::b200::EulerUpstream<GI, RP, BC>::transportSolve as a reference driver makes it
// (SimulatorTester.hpp:83-85) -- gathering pressure_sol.outflux(f) over all half-faces, a pageable
// std::vector<double>& saturation in and out, the PCIe copies, the device solve -- on the FlatGrid mock with the
// reference's own ReservoirPropertyCapillary<3> and BasicBoundaryConditions objects.  Not a parity test
// (tests/cpp/dropin_test.cpp is); built by oracle/Makefile where /root/reference exists, the binary travels to the GPU box.
//   dropin_bench <dir for rock files> <nx> <ny> <nz> <substeps> <calls> [devices "0,1"]
// Prints one JSON line: cell-substeps/s of the whole call, of the device part, and the host phases.
#include "../../tests/cpp/fixtures.hpp"

#include <chrono>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage: dropin_bench dir nx ny nz substeps calls [devices]\n"); return 2; }
    const std::string dir = argv[1];
    const int nx = std::atoi(argv[2]), ny = std::atoi(argv[3]), nz = std::atoi(argv[4]), substeps = std::atoi(argv[5]),
              calls = std::atoi(argv[6]);
    typedef Opm::ReservoirPropertyCapillary<3> RP;
    Rng rng(2024);
    GI grid;
    BCs bc;
    double t0 = now();
    buildGrid(grid, bc, nx, ny, nz, rng, false);
    const int N = grid.numberOfCells();
    RP rp;
    initProps(rp, N, rng, 1, dir, false);
    FlatFlux flux;
    const double v[3] = { 1e-6, 5e-7, 2.5e-7 };
    const flatgrid::Data& d = grid.data();
    flux.v.reserve(d.hf_area.size());
    for (size_t h = 0; h < d.hf_area.size(); ++h)
        flux.v.push_back(((v[0]*d.hf_normal[3*h] + v[1]*d.hf_normal[3*h + 1]) + v[2]*d.hf_normal[3*h + 2])*d.hf_area[h]);
    std::vector<double> sat(N);
    for (int c = 0; c < N; ++c) sat[c] = 0.3 + 0.2*(rng.next() - 0.5);
    Opm::SparseVector<double> inj(N);
    GI::Vector g(0.0);
    g[2] = -9.80665;
    const double t_build = now() - t0;

    Opm::parameter::ParameterGroup param;
    param.insertParameter("minimum_small_steps", substeps);
    param.insertParameter("maximum_small_steps", substeps);
    param.insertParameter("method_capillary", false);
    if (argc > 7) param.insertParameter("b200_devices", std::string(argv[7]));
    Opm::b200::EulerUpstream<GI, RP, BCs> dev;
    t0 = now();
    dev.init(param, grid, rp, bc);
    const double t_init = now() - t0;
    // a time step well inside the CFL limit of the viscous term: substeps * 0.25 * dt_cfl would need the CFL time; a
    // small fixed horizon keeps the saturations in range for this flux field
    const double time = 10.0*substeps;
    dev.transportSolve(sat, time, g, flux, inj);                     // warm-up (pins nothing: caller buffers stay pageable)
    t0 = now();
    double dev_ms = 0.0;
    for (int k = 0; k < calls; ++k) {
        dev.transportSolve(sat, time, g, flux, inj);
        dev_ms += dev.lastReport().device_ms;
    }
    const double wall = now() - t0;
    std::printf("{\"tool\": \"dropin_bench\", \"cells\": %d, \"substeps_per_call\": %d, \"calls\": %d, \"devices\": \"%s\", "
                "\"plugin_call_cell_substeps_per_s\": %.6e, \"device_substep_loop_cell_substeps_per_s\": %.6e, "
                "\"ms_per_call\": %.3f, \"device_ms_per_call\": %.3f, \"initObj_s\": %.2f, \"grid_build_s\": %.2f, "
                "\"pin_cache\": \"%s\", \"note\": \"transportSolve of the C++ drop-in: outflux gather + pageable saturation vector + PCIe copies + solve\"}\n",
                N, substeps, calls, argc > 7 ? argv[7] : "0", double(N)*substeps*calls/wall, double(N)*substeps*calls/(dev_ms*1e-3),
                1e3*wall/calls, dev_ms/calls, t_init, t_build, std::getenv("EU_PIN_CACHE") ? std::getenv("EU_PIN_CACHE") : "0");
    return 0;
}
