// TEST INFRASTRUCTURE.  Drop-in check of opm-porsol_b200/host/opm/porsol/euler/{EulerUpstream,EulerUpstreamResidual}.hpp
// and b200/Diagnostics.hpp:
// the reference's Opm::EulerUpstream (from /root/reference, unmodified) and Opm::b200::EulerUpstream are
// instantiated with the SAME GridInterface (FlatGrid mock), the reference's own
// ReservoirPropertyCapillary<3> / ReservoirPropertyCapillaryAnisotropicRelperm<3> and
// BasicBoundaryConditions<true,true> objects, driven with identical inputs, and compared.
// Built here by oracle/Makefile into oracle/_ref/dropin_test (needs /root/reference); run on the GPU box
// by tests/test_dropin_cpp.py.
#include "fixtures.hpp"

template <class RP, class Flux>
static int runCase(const char* name, int nx, int ny, int nz, int n_rocks, bool aniso, bool periodic_x, int mode,
                   double tol, const std::string& dir, bool expect_failure, const char* devices = 0)
{
    Rng rng(12345 + nx*7 + ny*13 + nz*31 + n_rocks);
    GI grid;
    BCs bc;
    buildGrid(grid, bc, nx, ny, nz, rng, periodic_x);
    const int N = grid.numberOfCells();
    RP rp;
    initProps(rp, N, rng, n_rocks, dir, aniso);
    Flux flux;
    const double v[3] = { 1e-6, -4e-7, 2.5e-7 };
    const flatgrid::Data& d = grid.data();
    for (size_t h = 0; h < d.hf_area.size(); ++h) {
        const double vn = (v[0]*d.hf_normal[3*h] + v[1]*d.hf_normal[3*h + 1]) + v[2]*d.hf_normal[3*h + 2];
        flux.v.push_back(vn*d.hf_area[h]);
    }
    std::vector<double> s0(N);
    for (int c = 0; c < N; ++c) s0[c] = 0.1 + 0.8*rng.next();
    Opm::SparseVector<double> inj(N);
    inj.addElement(expect_failure ? 5e-4 : 1.5e-7, 3);
    inj.addElement(-1.0e-7, N - 2);
    GI::Vector g(0.0);
    g[2] = -9.80665; g[0] = 0.2;

    Opm::parameter::ParameterGroup param;
    param.insertParameter("maximum_small_steps", expect_failure ? 2 : 40);
    param.insertParameter("b200_mode", mode);
    if (devices) param.insertParameter("b200_devices", std::string(devices));      // several devices of this process
    Opm::EulerUpstream<GI, RP, BCs> ref;
    ref.init(param, grid, rp, bc);
    Opm::b200::EulerUpstream<GI, RP, BCs> dev;
    dev.init(param, grid, rp, bc);

    const double time = expect_failure ? 3.0e5 : 2.0e4;
    std::vector<double> a(s0), b(s0);
    std::string ea, eb;
    for (int call = 0; call < 2; ++call) {        // two IMPES steps: state carried over
        try { ref.transportSolve(a, time, g, flux, inj); } catch (const std::exception& e) { ea = e.what(); }
        try { dev.transportSolve(b, time, g, flux, inj); } catch (const std::exception& e) { eb = e.what(); }
        if (!ea.empty() || !eb.empty()) break;
    }
    double maxdiff = 0.0;
    for (int c = 0; c < N; ++c) maxdiff = std::max(maxdiff, std::fabs(a[c] - b[c]));
    bool ok;
    if (expect_failure) {
        // "Saturation out of range in EulerUpstream: Cell <c>   sat <s>": same cell, same text up to the value
        const std::string pa = ea.substr(0, ea.find("sat ")), pb = eb.substr(0, eb.find("sat "));
        ok = !ea.empty() && pa == pb && dev.lastReport().attempts == 11;
    } else {
        ok = ea.empty() && eb.empty() && maxdiff <= tol;
    }
    std::printf("%-34s N=%5d mode=%d steps=%4d attempts=%2d maxdiff=%.3e %s%s\n", name, N, mode, dev.lastReport().nsteps,
                dev.lastReport().attempts, maxdiff, ok ? "OK" : "FAILED", (ea.empty() && eb.empty()) ? "" : (" [" + eb + "]").c_str());
    return ok ? 0 : 1;
}

// Opm::EulerUpstreamResidual (reference) against Opm::b200::EulerUpstreamResidual, called the way
// ImplicitCapillarity::transportSolve does (ImplicitCapillarity_impl.hpp:176-183: capillary = false, result negated
// into injection rates) and with all terms; then the post-transport loops of SimulatorUtilities.hpp against
// b200/Diagnostics.hpp on the same objects.
template <class RP, class Flux>
static int runResidualCase(const char* name, int nx, int ny, int nz, int n_rocks, bool aniso, bool periodic_x, int mode,
                           double rel_tol, const std::string& dir)
{
    Rng rng(777 + nx*7 + ny*13 + nz*31 + n_rocks);
    GI grid;
    BCs bc;
    buildGrid(grid, bc, nx, ny, nz, rng, periodic_x);
    const int N = grid.numberOfCells();
    RP rp;
    initProps(rp, N, rng, n_rocks, dir, aniso);
    Flux flux;
    const double v[3] = { 1e-6, -4e-7, 2.5e-7 };
    const flatgrid::Data& d = grid.data();
    for (size_t h = 0; h < d.hf_area.size(); ++h) {
        const double vn = (v[0]*d.hf_normal[3*h] + v[1]*d.hf_normal[3*h + 1]) + v[2]*d.hf_normal[3*h + 2];
        flux.v.push_back(vn*d.hf_area[h]*(0.9 + 0.2*rng.next()));
    }
    std::vector<double> s0(N);
    for (int c = 0; c < N; ++c) s0[c] = 0.1 + 0.8*rng.next();
    Opm::SparseVector<double> inj(N);
    inj.addElement(1.5e-7, 3);
    inj.addElement(-1.0e-7, N - 2);
    GI::Vector g(0.0);
    g[2] = -9.80665; g[0] = 0.2;

    Opm::EulerUpstreamResidual<GI, RP, BCs> ref(grid, rp, bc);
    Opm::b200::EulerUpstreamResidual<GI, RP, BCs> dev;
    dev.setDevice(0, mode);
    dev.initObj(grid, rp, bc);
    bool ok = &dev.grid() == &grid && &dev.reservoirProperties() == &rp && &dev.boundaryConditions() == &bc;
    double worst = 0.0;
    const bool flags[3][3] = { { true, true, false }, { true, true, true }, { false, true, true } };
    for (int k = 0; k < 3; ++k) {
        std::vector<double> a, b(7, 1.0);            // sat_delta is cleared and resized by the call
        if (flags[k][2]) ref.computeCapPressures(s0);
        ref.computeResidual(s0, g, flux, inj, flags[k][0], flags[k][1], flags[k][2], a);
        dev.computeResidual(s0, g, flux, inj, flags[k][0], flags[k][1], flags[k][2], b);
        ok = ok && int(b.size()) == N;
        double scale = 0.0, diff = 0.0;
        for (int c = 0; c < N && ok; ++c) { scale = std::max(scale, std::fabs(a[c])); diff = std::max(diff, std::fabs(a[c] - b[c])); }
        worst = std::max(worst, diff/scale);
        ok = ok && diff <= rel_tol*scale;
    }
    // diagnostics: bit-identical in every mode
    std::vector<GI::Vector> cv_ref, cv_dev, vw_ref, vo_ref, vw_dev, vo_dev;
    Opm::estimateCellVelocity(cv_ref, grid, flux);
    Opm::b200::estimateCellVelocity(cv_dev, dev);
    Opm::computePhaseVelocities(vw_ref, vo_ref, rp, s0, cv_ref);
    Opm::b200::computePhaseVelocities(vw_dev, vo_dev, dev, s0, cv_dev);
    std::vector<double> pc_ref, pc_dev, ff_dev;
    Opm::computeCapPressure(pc_ref, rp, s0);
    Opm::b200::computeCapPressure(pc_dev, dev, s0);
    Opm::b200::computeFractionalFlow(ff_dev, dev, s0);
    dev.computeCapPressures(s0);
    bool diag_ok = int(cv_dev.size()) == N;
    for (int c = 0; c < N && diag_ok; ++c) {
        for (int k = 0; k < 3; ++k) {
            diag_ok = diag_ok && cv_ref[c][k] == cv_dev[c][k] && vw_ref[c][k] == vw_dev[c][k] && vo_ref[c][k] == vo_dev[c][k];
        }
        diag_ok = diag_ok && pc_ref[c] == pc_dev[c] && dev.capPressures()[c] == pc_ref[c] && ff_dev[c] == rp.fractionalFlow(c, s0[c]);
    }
    std::printf("%-34s N=%5d mode=%d residual rel.diff=%.3e diagnostics %s %s\n", name, N, mode, worst,
                diag_ok ? "bit-identical" : "DIFFER", (ok && diag_ok) ? "OK" : "FAILED");
    return (ok && diag_ok) ? 0 : 1;
}

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    typedef Opm::ReservoirPropertyCapillary<3> RPS;
    typedef Opm::ReservoirPropertyCapillaryAnisotropicRelperm<3> RPA;
    int bad = 0;
    bad += runCase<RPS, IterFlux>("scalar 2 rocks periodic strict", 7, 5, 4, 2, false, true, 1, 0.0, dir, false);
    bad += runCase<RPS, IterFlux>("scalar 2 rocks periodic fast", 7, 5, 4, 2, false, true, 2, 1e-9, dir, false);
    bad += runCase<RPS, FlatFlux>("scalar 3 rocks dirichlet auto", 6, 6, 5, 3, false, false, 0, 1e-9, dir, false);
    bad += runCase<RPS, FlatFlux>("scalar no rocks strict", 5, 4, 6, 0, false, true, 1, 0.0, dir, false);
    bad += runCase<RPA, IterFlux>("tensor 2 rocks periodic strict", 5, 4, 3, 2, true, true, 1, 0.0, dir, false);
    bad += runCase<RPA, FlatFlux>("tensor 2 rocks periodic auto=fast", 5, 4, 3, 2, true, true, 0, 1e-9, dir, false);
    bad += runCase<RPS, IterFlux>("failure after 10 retries strict", 5, 4, 3, 1, false, false, 1, 0.0, dir, true);
    bad += runCase<RPS, IterFlux>("failure after 10 retries fast", 5, 4, 3, 1, false, false, 2, 0.0, dir, true);
    if (eu_device_count() >= 2) {
        // the grid split into two slabs over two devices of this process: STRICT stays bit-identical to the reference
        bad += runCase<RPS, IterFlux>("2 devices: scalar 2 rocks strict", 7, 5, 8, 2, false, true, 1, 0.0, dir, false, "0,1");
        bad += runCase<RPS, FlatFlux>("2 devices: scalar 3 rocks fast", 16, 4, 12, 3, false, false, 2, 1e-9, dir, false, "0,1");
        bad += runCase<RPS, IterFlux>("2 devices: failure strict", 5, 4, 6, 1, false, false, 1, 0.0, dir, true, "0,1");
    } else {
        std::printf("(one device: the 2-device cases are skipped)\n");
    }
    bad += runResidualCase<RPS, IterFlux>("residual scalar 2 rocks strict", 6, 5, 4, 2, false, true, 1, 0.0, dir);
    bad += runResidualCase<RPS, FlatFlux>("residual scalar 3 rocks fast", 6, 5, 4, 3, false, false, 2, 1e-12, dir);
    bad += runResidualCase<RPS, IterFlux>("residual scalar no rocks auto", 5, 4, 4, 0, false, true, 0, 1e-12, dir);
    bad += runResidualCase<RPA, IterFlux>("residual tensor 2 rocks strict", 5, 4, 3, 2, true, true, 1, 0.0, dir);
    bad += runResidualCase<RPA, IterFlux>("residual tensor 2 rocks fast", 5, 4, 3, 2, true, true, 2, 1e-12, dir);
    std::printf("%s\n", bad ? "DROPIN TEST FAILED" : "DROPIN TEST PASSED");
    return bad;
}
