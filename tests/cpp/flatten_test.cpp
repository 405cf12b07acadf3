// TEST INFRASTRUCTURE.  Host-side logic of the drop-in headers (opm-porsol_b200/host) checked WITHOUT a GPU: the headers
// are linked against the recording stub of tests/cpp/abi_stub.cpp instead of libeuler_b200.so, the reference's own
// ReservoirProperty / BoundaryConditions classes sit on the FlatGrid mock, and what arrives at the C ABI is compared
// with what the reference's objects say:
//   - the flattening walk (cell / face iteration order, neighbours, geometry, boundary tags, periodic partners),
//   - the fluid description (viscosities, densities, CFL factors, rock tables and rock ids -- checked through the
//     reference's own phaseMobility / capillaryPressure at sample saturations),
//   - parameter keys and defaults (EulerUpstream_impl.hpp:95-108), the two extra keys,
//   - flux gathering in half-face order for both PressureSolution flavours, sources from the SparseVector,
//   - the mapping of ABI status codes to the reference's exceptions, re-flattening on every initObj,
//   - EulerUpstreamResidual's argument passing and the Diagnostics wrappers' shapes.
// Built by oracle/Makefile into oracle/_ref/flatten_test (needs /root/reference); run by tests/test_dropin_cpp.py.
#include "fixtures.hpp"
#include "abi_stub.hpp"

static int g_bad = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("  CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_bad; } } while (0)

static double interp(const std::vector<double>& x, const std::vector<double>& y, int b, int e, double v)
{
    int i = b;
    while (i + 2 < e && v >= x[i + 1]) ++i;
    return (y[i + 1] - y[i])/(x[i + 1] - x[i])*(v - x[i]) + y[i];
}

template <class RP> struct Kind { enum { tensor = 0 }; };
template <> struct Kind<Opm::ReservoirPropertyCapillaryAnisotropicRelperm<3> > { enum { tensor = 1 }; };

static double mobEntry(const Opm::ScalarMobility& m, int) { return m.mob; }
static double mobEntry(const Opm::TensorMobility<3>& m, int axis) { return m.mob(axis, axis); }

template <class RP, class Flux>
static void flattenCase(const char* name, int nx, int ny, int nz, int n_rocks, bool aniso, bool periodic_x, const std::string& dir)
{
    const int bad0 = g_bad;
    Rng rng(4242 + nx*7 + ny*13 + nz*31 + n_rocks);
    GI grid;
    BCs bc;
    buildGrid(grid, bc, nx, ny, nz, rng, periodic_x);
    const int N = grid.numberOfCells();
    RP rp;
    initProps(rp, N, rng, n_rocks, dir, aniso);
    const flatgrid::Data& d = grid.data();
    const long long H = d.hf_offset[N];

    Opm::parameter::ParameterGroup param;
    param.insertParameter("courant_number", 0.3);
    param.insertParameter("method_capillary", false);
    param.insertParameter("use_cfl_gravity", false);
    param.insertParameter("maximum_small_steps", 55);
    param.insertParameter("check_sat", false);
    param.insertParameter("clamp_sat", true);
    param.insertParameter("b200_mode", 2);
    Opm::b200::EulerUpstream<GI, RP, BCs> dev;
    dev.init(param, grid, rp, bc);
    StubRecording& r = stub_recording();

    // parameters: given keys taken, the others at the reference's defaults (EulerUpstream_impl.hpp:59-73)
    CHECK(r.params.courant_number == 0.3 && r.params.method_viscous == 1 && r.params.method_gravity == 1 && r.params.method_capillary == 0);
    CHECK(r.params.use_cfl_viscous == 1 && r.params.use_cfl_gravity == 0 && r.params.use_cfl_capillary == 1);
    CHECK(r.params.minimum_small_steps == 1 && r.params.maximum_small_steps == 55 && r.params.check_sat == 0 && r.params.clamp_sat == 1);
    CHECK(r.cfg.mode == 2 && r.cfg.device == 0 && r.cfg.world_size == 1 && r.cfg.own_begin == 0 && r.cfg.own_end == N);
    // grid: the reference's iteration order is the half-face index
    CHECK(r.grid_ended && r.n_global == N && r.n_local == N && r.n_hf == H && int(r.hf_count.size()) == N);
    CHECK((long long)r.hf_neighbour.size() == H && (long long)r.hf_area.size() == H && (long long)r.hf_normal.size() == 3*H);
    size_t nb = 0;
    std::vector<int> hf_of_bid(bc.size(), -1);
    for (long long h = 0; h < H; ++h) if (d.hf_bid[h] > 0) hf_of_bid[d.hf_bid[h]] = int(h);
    for (int c = 0; c < N && g_bad == bad0; ++c) {
        CHECK(r.hf_count[c] == d.hf_offset[c + 1] - d.hf_offset[c]);
        CHECK(r.cell_volume[c] == d.cell_volume[c] && r.porosity[c] == rp.porosity(c));
        typename RP::PermTensor K = rp.permeability(c);
        for (int i = 0; i < 3; ++i) {
            CHECK(r.cell_centroid[3*c + i] == d.cell_centroid[3*c + i]);
            for (int j = 0; j < 3; ++j) CHECK(r.permeability[9*c + 3*i + j] == K(i, j));
        }
        for (int h = d.hf_offset[c]; h < d.hf_offset[c + 1]; ++h) {
            CHECK(r.hf_area[h] == d.hf_area[h]);
            for (int i = 0; i < 3; ++i) CHECK(r.hf_normal[3*h + i] == d.hf_normal[3*h + i] && r.hf_centroid[3*h + i] == d.hf_centroid[3*h + i]);
            const int bid = d.hf_bid[h];
            if (bid == 0) {
                CHECK(r.hf_neighbour[h] == d.hf_neighbour[h] && r.hf_neighbour[h] >= 0);
                continue;
            }
            CHECK(r.hf_neighbour[h] == -1);
            CHECK(nb < r.bnd_hf.size() && r.bnd_hf[nb] == h);
            if (nb >= r.bnd_hf.size()) break;
            if (bc.satCond(bid).isPeriodic()) {
                const int ph = hf_of_bid[bc.getPeriodicPartner(bid)];
                int pc = 0;
                while (d.hf_offset[pc + 1] <= ph) ++pc;
                CHECK(r.bnd_kind[nb] == EU_HF_PERIODIC && r.bnd_partner_cell[nb] == pc && r.bnd_partner_face[nb] == ph - d.hf_offset[pc]);
            } else {
                CHECK(r.bnd_kind[nb] == EU_HF_DIRICHLET && r.bnd_sat[nb] == bc.satCond(bid).saturation());
            }
            ++nb;
        }
    }
    CHECK(nb == r.bnd_hf.size());
    // fluid: scalars straight from the property object; tables and rock ids through the reference's own curves
    CHECK(r.fluid.mobility_kind == (aniso ? EU_MOB_DIAGONAL : EU_MOB_SCALAR) && r.fluid.n_rocks == n_rocks);
    CHECK(r.fluid.cfl_factor[0] == rp.cflFactor() && r.fluid.cfl_factor[1] == rp.cflFactorGravity() && r.fluid.cfl_factor[2] == rp.cflFactorCapillary());
    CHECK(r.fluid.density[0] - r.fluid.density[1] == rp.densityDifference());
    CHECK((n_rocks > 0) == !r.rock_id.empty());
    const double samples[3] = { 0.17, 0.52, 0.88 };
    for (int c = 0; c < N && g_bad == bad0; c += 3) {
        for (int q = 0; q < 3; ++q) {
            const double s = samples[q];
            for (int phase = 0; phase < 2; ++phase) {
                for (int axis = 0; axis < (aniso ? 3 : 1); ++axis) {
                    typename RP::Mobility m;
                    rp.phaseMobility(phase, c, s, m.mob);
                    const double want = mobEntry(m, axis);
                    double kr;
                    if (n_rocks == 0) {
                        kr = phase == 0 ? s*s : (1 - s)*(1 - s);
                    } else {
                        const int rk = r.rock_id[c], b = r.tab_offset[rk], e = r.tab_offset[rk + 1];
                        const int col = aniso ? 1 + 3*phase + axis : phase;
                        kr = interp(r.tab_s, r.tab_cols[col], b, e, s);
                    }
                    CHECK(std::fabs(kr/r.fluid.viscosity[phase] - want) <= 1e-14*std::fabs(want));
                }
            }
            double pc;
            if (n_rocks == 0) {
                pc = 1e5*(1 - s);
            } else {
                const int rk = r.rock_id[c], b = r.tab_offset[rk], e = r.tab_offset[rk + 1];
                pc = interp(r.tab_s, r.tab_cols[aniso ? 0 : 2], b, e, s);
                if (!aniso && r.fluid.use_jfunction_scaling) {
                    typename RP::PermTensor K = rp.permeability(c);
                    pc = pc*r.fluid.sigma_cos_theta/std::sqrt((K(0, 0) + K(1, 1) + K(2, 2))/(3*rp.porosity(c)));
                }
            }
            const double want = rp.capillaryPressure(c, s);
            CHECK(std::fabs(pc - want) <= 1e-13*std::fabs(want));
        }
    }

    // transportSolve: inputs in half-face order, sources as (cell, rate) pairs, result written back
    Flux flux;
    for (long long h = 0; h < H; ++h) flux.v.push_back(1e-6*(rng.next() - 0.5));
    std::vector<double> sat(N);
    for (int c = 0; c < N; ++c) sat[c] = rng.next();
    const std::vector<double> sat0(sat);
    Opm::SparseVector<double> inj(N);
    inj.addElement(1.5e-7, 3);
    inj.addElement(-1.0e-7, N - 2);
    GI::Vector g(0.0);
    g[0] = 0.2; g[1] = -0.1; g[2] = -9.80665;
    const int allocs_before = r.n_host_alloc;
    dev.transportSolve(sat, 1234.5, g, flux, inj);
    CHECK(r.time == 1234.5 && r.gravity[0] == 0.2 && r.gravity[1] == -0.1 && r.gravity[2] == -9.80665);
    CHECK(r.flux == flux.v && r.sat_in == sat0);
    CHECK(r.src_cell.size() == 2 && r.src_cell[0] == 3 && r.src_cell[1] == N - 2 && r.src_rate[0] == 1.5e-7 && r.src_rate[1] == -1.0e-7);
    for (int c = 0; c < N; ++c) CHECK(sat[c] == sat0[c] + 1.0);
    CHECK(dev.lastReport().nsteps == 7);
    CHECK(r.n_host_alloc == allocs_before);           // the flux buffer was allocated once, at initObj

    // status codes -> the reference's exceptions (EulerUpstream_impl.hpp:344-346, CflCalculator.hpp:75-77)
    std::string msg;
    r.next_status = EU_ERR_SAT_RANGE; r.next_attempts = 11; r.next_bad_cell = 5; r.next_bad_value = 1.25;
    try { dev.transportSolve(sat, 1.0, g, flux, inj); } catch (const std::runtime_error& e) { msg = e.what(); }
    CHECK(msg.find("Saturation out of range in EulerUpstream: Cell 5   sat 1.25") != std::string::npos);
    msg.clear();
    r.next_status = EU_ERR_CFL_ZERO;
    try { dev.transportSolve(sat, 1.0, g, flux, inj); } catch (const std::runtime_error& e) { msg = e.what(); }
    CHECK(msg.find("Cfl computation gave dt = 0.0") != std::string::npos);
    msg.clear();
    r.next_status = EU_ERR_CUDA;
    try { dev.transportSolve(sat, 1.0, g, flux, inj); } catch (const std::runtime_error& e) { msg = e.what(); }
    CHECK(msg.find("EulerUpstream (B200): stub failure") != std::string::npos);
    r.next_status = EU_OK; r.next_attempts = 1;
    std::vector<double> wrong(N + 1, 0.0);
    msg.clear();
    try { dev.transportSolve(wrong, 1.0, g, flux, inj); } catch (const std::runtime_error& e) { msg = e.what(); }
    CHECK(!msg.empty());

    // setCourantNumber and a second initObj: parameters forwarded, everything flattened again
    dev.setCourantNumber(0.45);
    CHECK(stub_recording().params.courant_number == 0.45);
    const int destroyed = stub_recording().n_destroy, chunks = stub_recording().n_chunks;
    dev.initObj(grid, rp, bc);
    CHECK(stub_recording().n_destroy == destroyed + 1 && stub_recording().n_chunks >= 1 && (long long)stub_recording().hf_area.size() == H);
    (void)chunks;

    // EulerUpstreamResidual mirror and the Diagnostics wrappers
    Opm::b200::EulerUpstreamResidual<GI, RP, BCs> res(grid, rp, bc);
    CHECK(&res.grid() == &grid && &res.reservoirProperties() == &rp && &res.boundaryConditions() == &bc);
    std::vector<double> delta(3, 9.0);
    res.computeResidual(sat0, g, flux, inj, true, false, true, delta);
    StubRecording& r2 = stub_recording();
    CHECK(int(delta.size()) == N && delta[N - 1] == double(N - 1));
    CHECK(r2.methods[0] == 1 && r2.methods[1] == 0 && r2.methods[2] == 1 && r2.flux == flux.v && r2.sat_in == sat0 && r2.src_cell.size() == 2);
    res.computeCapPressures(sat0);
    CHECK(int(res.capPressures().size()) == N && res.capPressures()[1] == 2.0*sat0[1]);
    std::vector<GI::Vector> cv, vw, vo;
    Opm::b200::estimateCellVelocity(cv, res);
    CHECK(int(cv.size()) == N && cv[1][2] == 5.0);
    Opm::b200::computePhaseVelocities(vw, vo, res, sat0, cv);
    CHECK(int(vw.size()) == N && vw[1][2] == 1.25 && vo[1][2] == 3.75);
    std::vector<double> pc, ff;
    Opm::b200::computeCapPressure(pc, res, sat0);
    Opm::b200::computeFractionalFlow(ff, res, sat0);
    CHECK(int(pc.size()) == N && pc[2] == 2.0*sat0[2] && int(ff.size()) == N && ff[2] == 0.5*sat0[2]);

    std::printf("%-40s N=%5d half-faces=%6lld boundary=%5zu  %s\n", name, N, H, nb, g_bad == bad0 ? "OK" : "FAILED");
}

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    typedef Opm::ReservoirPropertyCapillary<3> RPS;
    typedef Opm::ReservoirPropertyCapillaryAnisotropicRelperm<3> RPA;
    flattenCase<RPS, IterFlux>("scalar 2 rocks, periodic x, iterator flux", 6, 5, 4, 2, false, true, dir);
    flattenCase<RPS, FlatFlux>("scalar 3 rocks, dirichlet, flat flux", 5, 4, 3, 3, false, false, dir);
    flattenCase<RPS, FlatFlux>("scalar no rocks, periodic x", 4, 4, 4, 0, false, true, dir);
    flattenCase<RPA, IterFlux>("tensor 2 rocks, periodic x", 5, 4, 3, 2, true, true, dir);
    std::printf("%s\n", g_bad ? "FLATTEN TEST FAILED" : "FLATTEN TEST PASSED");
    return g_bad ? 1 : 0;
}
