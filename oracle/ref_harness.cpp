// TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load what this file builds.
//
// C-callable harness around the UNMODIFIED reference headers
//   /root/reference/opm/porsol/euler/EulerUpstream.hpp (+_impl, EulerUpstreamResidual*,
//   CflCalculator.hpp) and opm/porsol/common/{ReservoirPropertyCapillary*,RockJfunc,
//   BoundaryConditions,Matrix,MatrixInverse}.hpp,
// compiled where they lie, against oracle/shims (third-party stand-ins) and the FlatGrid
// mock GridInterface.  Built by oracle/Makefile into oracle/_ref/libeuler_ref.so.
// Nothing of the reference is copied: this file only instantiates and calls it.
#include <opm/porsol/euler/EulerUpstream.hpp>
#include <opm/porsol/common/ReservoirPropertyCapillary.hpp>
#ifdef REF_WITH_ANISO
#include <opm/porsol/common/ReservoirPropertyCapillaryAnisotropicRelperm.hpp>
#endif
#include <opm/porsol/common/BoundaryConditions.hpp>
#include <opm/porsol/common/SimulatorUtilities.hpp>
#include <opm/porsol/common/BoundaryPeriodicity.hpp>

#include "FlatGrid.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <memory>
#include <chrono>

namespace {

    typedef flatgrid::Grid GI;
    typedef Opm::BasicBoundaryConditions<true, true> BCs;

    struct FlatFlux {
        const double* hf_flux;
        double outflux(const GI::CellIterator::FaceIterator& f) const { return hf_flux[f->halfFaceIndex()]; }
    };

    template <class RP>
    struct Exposed : public Opm::EulerUpstream<GI, RP, BCs> {
        typedef Opm::EulerUpstream<GI, RP, BCs> Base;
        using Base::smallTimeStep;
        using Base::computeCflTime;
        using Base::residual_computer_;
        using Base::residual_;
        using Base::method_viscous_;
        using Base::method_gravity_;
        using Base::method_capillary_;
    };

    struct HarnessBase {
        virtual ~HarnessBase() {}
        GI grid;
        BCs bc;
        std::string last_error;
        int last_nsteps;
        int last_attempts;
        double last_seconds;
        virtual void setParams(const Opm::parameter::ParameterGroup& p) = 0;
        virtual void initObj() = 0;
        virtual int transportSolve(std::vector<double>& sat, double time, const GI::Vector& g,
                                   const FlatFlux& flux, const Opm::SparseVector<double>& inj) = 0;
        virtual int smallStep(std::vector<double>& sat, double dt, const GI::Vector& g,
                              const FlatFlux& flux, const Opm::SparseVector<double>& inj,
                              std::vector<double>& residual) = 0;
        virtual void cflTimes(const GI::Vector& g, const FlatFlux& flux, double* out3, double* total) = 0;
        virtual void cflFactors(double* out3) const = 0;
        virtual double capPressure(int cell, double s) const = 0;
        virtual void mobility(int phase, int cell, double s, double* out9) const = 0;
        virtual double fracFlow(int cell, double s) const = 0;
        virtual double porosity(int cell) const = 0;
        // EulerUpstreamResidual::computeResidual with explicit method flags (Residual_impl.hpp:472-505)
        virtual void computeResidual(const std::vector<double>& sat, const GI::Vector& g, const FlatFlux& flux,
                                     const Opm::SparseVector<double>& inj, bool mv, bool mg, bool mc,
                                     std::vector<double>& out) = 0;
        // SimulatorUtilities.hpp:153-170, :219-230
        virtual void phaseVelocities(const std::vector<double>& sat, const std::vector<GI::Vector>& cv,
                                     std::vector<GI::Vector>& vw, std::vector<GI::Vector>& vo) const = 0;
        virtual void capPressures(const std::vector<double>& sat, std::vector<double>& out) const = 0;
    };

    template <class RP> struct MobOut;
    template <> struct MobOut<Opm::ReservoirPropertyCapillary<3> > {
        static void get(const Opm::ReservoirPropertyCapillary<3>& rp, int phase, int cell, double s, double* out9)
        {
            double m; rp.phaseMobility(phase, cell, s, m);
            for (int i = 0; i < 9; ++i) out9[i] = 0.0;
            out9[0] = out9[4] = out9[8] = m;
        }
    };
#ifdef REF_WITH_ANISO
    template <> struct MobOut<Opm::ReservoirPropertyCapillaryAnisotropicRelperm<3> > {
        static void get(const Opm::ReservoirPropertyCapillaryAnisotropicRelperm<3>& rp, int phase, int cell, double s, double* out9)
        {
            Opm::TensorMobility<3> m; rp.phaseMobility(phase, cell, s, m.mob);
            for (int i = 0; i < 9; ++i) out9[i] = m.mob.data()[i];
        }
    };
#endif

    template <class RP>
    struct Harness : public HarnessBase {
        RP rp;
        Exposed<RP> solver;

        void setParams(const Opm::parameter::ParameterGroup& p) { solver.init(p); }
        void initObj() { solver.initObj(grid, rp, bc); }

        int transportSolve(std::vector<double>& sat, double time, const GI::Vector& g,
                           const FlatFlux& flux, const Opm::SparseVector<double>& inj)
        {
            // The reference keeps the step count and retry count in locals
            // (EulerUpstream_impl.hpp:163,183); its VERBOSE build prints one
            // "Doing <n> steps ..." line per attempt (:189-193).  Capture and parse.
            std::ostringstream captured;
            std::streambuf* old = std::cout.rdbuf(captured.rdbuf());
            int status = 0;
            last_error.clear();
            std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
            try {
                solver.transportSolve(sat, time, g, flux, inj);
            } catch (const std::exception& e) {
                status = 1;
                last_error = e.what();
            } catch (...) {
                status = 1;
                last_error = "unknown exception";
            }
            last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::cout.rdbuf(old);
            last_nsteps = -1;
            last_attempts = 0;
            std::istringstream lines(captured.str());
            std::string line;
            while (std::getline(lines, line)) {
                if (line.compare(0, 6, "Doing ") == 0) {
                    std::istringstream ls(line.substr(6));
                    ls >> last_nsteps;
                    ++last_attempts;
                }
                const std::string key = "Seconds taken by transport solver: ";
                if (line.compare(0, key.size(), key) == 0) {
                    std::istringstream ls(line.substr(key.size()));
                    ls >> last_seconds;   // the reference's own StopWatch around the substep loop
                }
            }
            return status;
        }

        int smallStep(std::vector<double>& sat, double dt, const GI::Vector& g,
                      const FlatFlux& flux, const Opm::SparseVector<double>& inj,
                      std::vector<double>& residual)
        {
            int status = 0;
            last_error.clear();
            try {
                solver.smallTimeStep(sat, dt, g, flux, inj);
            } catch (const std::exception& e) {
                status = 1;
                last_error = e.what();
            }
            residual = solver.residual_;
            return status;
        }

        void cflTimes(const GI::Vector& g, const FlatFlux& flux, double* out3, double* total)
        {
            out3[0] = out3[1] = out3[2] = 1e99;
            last_error.clear();
            try {
                out3[0] = Opm::cfl_calculator::findCFLtimeVelocity(grid, rp, flux);
                out3[1] = Opm::cfl_calculator::findCFLtimeGravity(grid, rp, g);
                out3[2] = Opm::cfl_calculator::findCFLtimeCapillary(grid, rp);
                std::ostringstream captured;
                std::streambuf* old = std::cout.rdbuf(captured.rdbuf());
                std::vector<double> dummy;
                *total = solver.computeCflTime(dummy, 0.0, g, flux);
                std::cout.rdbuf(old);
            } catch (const std::exception& e) {
                last_error = e.what();
                *total = -1.0;
            }
        }
        void cflFactors(double* out3) const
        {
            out3[0] = rp.cflFactor(); out3[1] = rp.cflFactorGravity(); out3[2] = rp.cflFactorCapillary();
        }
        double capPressure(int cell, double s) const { return rp.capillaryPressure(cell, s); }
        void mobility(int phase, int cell, double s, double* out9) const { MobOut<RP>::get(rp, phase, cell, s, out9); }
        double fracFlow(int cell, double s) const { return rp.fractionalFlow(cell, s); }
        double porosity(int cell) const { return rp.porosity(cell); }
        void computeResidual(const std::vector<double>& sat, const GI::Vector& g, const FlatFlux& flux,
                             const Opm::SparseVector<double>& inj, bool mv, bool mg, bool mc, std::vector<double>& out)
        {
            // like smallTimeStep (EulerUpstream_impl.hpp:362-369): the cached capillary pressures first
            if (mc) solver.residual_computer_.computeCapPressures(sat);
            solver.residual_computer_.computeResidual(sat, g, flux, inj, mv, mg, mc, out);
        }
        void phaseVelocities(const std::vector<double>& sat, const std::vector<GI::Vector>& cv,
                             std::vector<GI::Vector>& vw, std::vector<GI::Vector>& vo) const
        {
            Opm::computePhaseVelocities(vw, vo, rp, sat, cv);
        }
        void capPressures(const std::vector<double>& sat, std::vector<double>& out) const
        {
            Opm::computeCapPressure(out, rp, sat);
        }
    };

    void writeTable(const std::string& fname, int npts, int ncol, const double* const* cols)
    {
        std::ofstream os(fname.c_str());
        os << "# generated by oracle/ref_harness.cpp\n";
        os << std::setprecision(17);
        for (int i = 0; i < npts; ++i) {
            for (int c = 0; c < ncol; ++c) { os << (c ? " " : "") << cols[c][i]; }
            os << "\n";
        }
    }

    // Minimal Dune GridView stand-in for findPeriodicPartners (BoundaryPeriodicity.hpp:86-177): one element that owns
    // all boundary intersections, each with a unique boundary id, a centroid and an area.
    struct BFaceView {
        const double* centroid;
        const double* area;
        int n;
        struct Geometry {
            Dune::FieldVector<double, 3> c;
            double a;
            Dune::FieldVector<double, 3> center() const { return c; }
            double volume() const { return a; }
        };
        struct Intersection {
            const BFaceView* v;
            int i;
            int boundaryId() const { return i + 1; }
            Geometry geometry() const
            {
                Geometry g;
                for (int d = 0; d < 3; ++d) g.c[d] = v->centroid[3*i + d];
                g.a = v->area[i];
                return g;
            }
        };
        struct IntersectionIterator {
            Intersection x;
            const Intersection* operator->() const { return &x; }
            IntersectionIterator& operator++() { ++x.i; return *this; }
            bool operator!=(const IntersectionIterator& o) const { return x.i != o.x.i; }
        };
        struct Element {};
        struct ElementIterator {
            int pos;
            Element e;
            const Element& operator*() const { return e; }
            ElementIterator& operator++() { ++pos; return *this; }
            bool operator!=(const ElementIterator& o) const { return pos != o.pos; }
        };
        enum { dimension = 3 };
        template <int codim> struct Codim { typedef ElementIterator Iterator; };
        template <int codim> ElementIterator begin() const { ElementIterator it; it.pos = 0; return it; }
        template <int codim> ElementIterator end() const { ElementIterator it; it.pos = 1; return it; }
        IntersectionIterator ibegin(const Element&) const { IntersectionIterator it; it.x.v = this; it.x.i = 0; return it; }
        IntersectionIterator iend(const Element&) const { IntersectionIterator it; it.x.v = this; it.x.i = n; return it; }
    };

    GI::Vector vec3(const double* p) { GI::Vector v; v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; return v; }

    Opm::SparseVector<double> makeInj(int N, int n_src, const int* src_cell, const double* src_rate)
    {
        Opm::SparseVector<double> inj(N);
        for (int i = 0; i < n_src; ++i) { inj.addElement(src_rate[i], src_cell[i]); }
        return inj;
    }
} // anon

extern "C" {

// mobility_kind: 0 = ReservoirPropertyCapillary<3> (scalar mobility, RockJfunc tables S,krw,kro,J)
//                1 = ReservoirPropertyCapillaryAnisotropicRelperm<3> (tables per phase: pcow,S,krxx,kryy,krzz)
// For kind 1 the table columns are passed as tab_a..: phase-0 file {pc,S,kx,ky,kz} and phase-1 file
// share S and pc; tab_krw/kro/J are reused as described in tests (see oracle/README.md).
void* ref_create(int N, const int* hf_offset, const int* hf_nbr, const int* hf_bid,
                 const double* hf_area, const double* hf_normal, const double* hf_centroid,
                 const double* cell_volume, const double* cell_centroid,
                 const double* poro, const double* perm9,
                 const int* rock_id, int n_rocks, const int* tab_offset,
                 const double* tab_s, const double* const* tab_cols, int n_tab_cols,
                 int use_j, double sigma, double theta,
                 const double* visc2, const double* dens2,
                 int n_bid, const int* bid_kind, const double* bid_sat, const int* bid_partner,
                 int mobility_kind, const char* tmpdir)
{
    HarnessBase* hb = 0;
    try {
        const int H = hf_offset[N];
        std::shared_ptr<Opm::Deck> deck(new Opm::Deck);
        deck->dims_[0] = N; deck->dims_[1] = 1; deck->dims_[2] = 1;
        std::vector<int> global_cell(N);
        for (int c = 0; c < N; ++c) global_cell[c] = c;
        deck->setDouble("PORO", std::vector<double>(poro, poro + N));
        const char* names[9] = { "PERMX", "PERMXY", "PERMXZ", 0, "PERMY", "PERMYZ", 0, 0, "PERMZ" };
        for (int k = 0; k < 9; ++k) {
            if (!names[k]) continue;
            std::vector<double> v(N);
            bool nonzero = false;
            for (int c = 0; c < N; ++c) { v[c] = perm9[9*c + k]; nonzero = nonzero || v[c] != 0.0; }
            const bool diag = (k == 0 || k == 4 || k == 8);
            if (diag || nonzero) deck->setDouble(names[k], v);
        }
        if (rock_id) {
            std::vector<int> satnum(N);
            for (int c = 0; c < N; ++c) satnum[c] = rock_id[c] + 1;
            deck->setInt("SATNUM", satnum);
        }
        std::string rocklist;
        std::string dir = std::string(tmpdir) + "/";
        if (n_rocks > 0) {
            rocklist = dir + "rocklist.txt";
            std::ofstream rl(rocklist.c_str());
            rl << n_rocks << "\n";
            for (int r = 0; r < n_rocks; ++r) {
                const int b = tab_offset[r], n = tab_offset[r+1] - tab_offset[r];
                std::ostringstream fn; fn << "rock" << r;
                if (mobility_kind == 0) {
                    // Statoil format: S krw kro J (RockJfunc.hpp:162-218)
                    const double* cols[4] = { tab_s + b, tab_cols[0] + b, tab_cols[1] + b, tab_cols[2] + b };
                    writeTable(dir + fn.str() + ".txt", n, 4, cols);
                    rl << fn.str() << ".txt\n";
                } else {
                    // Aniso format per phase: pcow S krxx kryy krzz (RockAnisotropicRelperm.hpp:109-152)
                    // tab_cols = { pc, kx0, ky0, kz0, kx1, ky1, kz1 }
                    const double* c0[5] = { tab_cols[0] + b, tab_s + b, tab_cols[1] + b, tab_cols[2] + b, tab_cols[3] + b };
                    const double* c1[5] = { tab_cols[0] + b, tab_s + b, tab_cols[4] + b, tab_cols[5] + b, tab_cols[6] + b };
                    writeTable(dir + fn.str() + "_w.txt", n, 5, c0);
                    writeTable(dir + fn.str() + "_o.txt", n, 5, c1);
                    rl << fn.str() << "_w.txt " << fn.str() << "_o.txt\n";
                }
            }
        }
        (void)n_tab_cols;

        if (mobility_kind == 0) {
            Harness<Opm::ReservoirPropertyCapillary<3> >* h = new Harness<Opm::ReservoirPropertyCapillary<3> >;
            hb = h;
            h->rp.setViscosities(visc2[0], visc2[1]);
            h->rp.setDensities(dens2[0], dens2[1]);
            std::ostringstream captured; std::streambuf* old = std::cout.rdbuf(captured.rdbuf());
            h->rp.init(deck, global_cell, 0.0, n_rocks > 0 ? &rocklist : 0, use_j != 0, sigma, theta);
            std::cout.rdbuf(old);
        }
#ifdef REF_WITH_ANISO
        else if (mobility_kind == 1) {
            typedef Opm::ReservoirPropertyCapillaryAnisotropicRelperm<3> RPA;
            Harness<RPA>* h = new Harness<RPA>;
            hb = h;
            h->rp.setViscosities(visc2[0], visc2[1]);
            h->rp.setDensities(dens2[0], dens2[1]);
            std::ostringstream captured; std::streambuf* old = std::cout.rdbuf(captured.rdbuf());
            h->rp.init(deck, global_cell, 0.0, n_rocks > 0 ? &rocklist : 0, false, sigma, theta);
            std::cout.rdbuf(old);
        }
#endif
        else {
            return 0;
        }

        flatgrid::Data& d = hb->grid.data();
        d.num_cells = N;
        d.hf_offset.assign(hf_offset, hf_offset + N + 1);
        d.hf_neighbour.assign(hf_nbr, hf_nbr + H);
        d.hf_bid.assign(hf_bid, hf_bid + H);
        d.hf_area.assign(hf_area, hf_area + H);
        d.hf_normal.assign(hf_normal, hf_normal + 3*H);
        d.hf_centroid.assign(hf_centroid, hf_centroid + 3*H);
        d.cell_volume.assign(cell_volume, cell_volume + N);
        d.cell_centroid.assign(cell_centroid, cell_centroid + 3*N);

        hb->bc.resize(n_bid);
        for (int b = 0; b < n_bid; ++b) {
            if (bid_kind[b] == 1) {
                hb->bc.satCond(b) = Opm::SatBC(Opm::SatBC::Periodic, 0.0);
                hb->bc.flowCond(b) = Opm::FlowBC(Opm::FlowBC::Periodic, 0.0);
                if (bid_partner[b] > b) hb->bc.setPeriodicPartners(b, bid_partner[b]);
            } else {
                hb->bc.satCond(b) = Opm::SatBC(Opm::SatBC::Dirichlet, bid_sat[b]);
            }
        }
        hb->initObj();
        return hb;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_create failed: %s\n", e.what());
        delete hb;
        return 0;
    }
}

void ref_destroy(void* h) { delete static_cast<HarnessBase*>(h); }

void ref_set_params(void* hv, double courant, int mv, int mg, int mc, int cv, int cg, int cc,
                    int min_steps, int max_steps, int check_sat, int clamp_sat)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    Opm::parameter::ParameterGroup p;
    p.insertParameter("courant_number", courant);
    p.insertParameter("method_viscous", mv != 0);
    p.insertParameter("method_gravity", mg != 0);
    p.insertParameter("method_capillary", mc != 0);
    p.insertParameter("use_cfl_viscous", cv != 0);
    p.insertParameter("use_cfl_gravity", cg != 0);
    p.insertParameter("use_cfl_capillary", cc != 0);
    p.insertParameter("minimum_small_steps", min_steps);
    p.insertParameter("maximum_small_steps", max_steps);
    p.insertParameter("check_sat", check_sat != 0);
    p.insertParameter("clamp_sat", clamp_sat != 0);
    h->setParams(p);
}

int ref_transport_solve(void* hv, double* sat, double time, const double* gravity, const double* hf_flux,
                        int n_src, const int* src_cell, const double* src_rate,
                        int* nsteps, int* attempts, double* loop_seconds)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    const int N = h->grid.numberOfCells();
    std::vector<double> s(sat, sat + N);
    FlatFlux flux = { hf_flux };
    Opm::SparseVector<double> inj = makeInj(N, n_src, src_cell, src_rate);
    int status = h->transportSolve(s, time, vec3(gravity), flux, inj);
    std::copy(s.begin(), s.end(), sat);
    if (nsteps) *nsteps = h->last_nsteps;
    if (attempts) *attempts = h->last_attempts;
    if (loop_seconds) *loop_seconds = h->last_seconds;
    return status;
}

int ref_small_step(void* hv, double* sat, double dt, const double* gravity, const double* hf_flux,
                   int n_src, const int* src_cell, const double* src_rate, double* residual_out)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    const int N = h->grid.numberOfCells();
    std::vector<double> s(sat, sat + N), res;
    FlatFlux flux = { hf_flux };
    Opm::SparseVector<double> inj = makeInj(N, n_src, src_cell, src_rate);
    int status = h->smallStep(s, dt, vec3(gravity), flux, inj, res);
    std::copy(s.begin(), s.end(), sat);
    if (residual_out && int(res.size()) == N) std::copy(res.begin(), res.end(), residual_out);
    return status;
}

void ref_cfl_times(void* hv, const double* gravity, const double* hf_flux, double* out3, double* total)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    FlatFlux flux = { hf_flux };
    h->cflTimes(vec3(gravity), flux, out3, total);
}

void ref_compute_residual(void* hv, const double* sat, const double* gravity, const double* hf_flux,
                          int n_src, const int* src_cell, const double* src_rate, int mv, int mg, int mc, double* out)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    const int N = h->grid.numberOfCells();
    std::vector<double> s(sat, sat + N), res;
    FlatFlux flux = { hf_flux };
    Opm::SparseVector<double> inj = makeInj(N, n_src, src_cell, src_rate);
    h->computeResidual(s, vec3(gravity), flux, inj, mv != 0, mg != 0, mc != 0, res);
    std::copy(res.begin(), res.end(), out);
}

// estimateCellVelocity (SimulatorUtilities.hpp:59-86); out: 3 doubles per cell
void ref_cell_velocity(void* hv, const double* hf_flux, double* out)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    FlatFlux flux = { hf_flux };
    std::vector<GI::Vector> cv;
    Opm::estimateCellVelocity(cv, h->grid, flux);
    for (size_t c = 0; c < cv.size(); ++c) for (int d = 0; d < 3; ++d) out[3*c + d] = cv[c][d];
}

// computePhaseVelocities (SimulatorUtilities.hpp:153-170)
void ref_phase_velocities(void* hv, const double* sat, const double* cell_velocity, double* vw_out, double* vo_out)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    const int N = h->grid.numberOfCells();
    std::vector<double> s(sat, sat + N);
    std::vector<GI::Vector> cv(N), vw, vo;
    for (int c = 0; c < N; ++c) for (int d = 0; d < 3; ++d) cv[c][d] = cell_velocity[3*c + d];
    h->phaseVelocities(s, cv, vw, vo);
    for (int c = 0; c < N; ++c) for (int d = 0; d < 3; ++d) { vw_out[3*c + d] = vw[c][d]; vo_out[3*c + d] = vo[c][d]; }
}

// computeCapPressure (SimulatorUtilities.hpp:219-230)
void ref_cap_pressures(void* hv, const double* sat, double* out)
{
    HarnessBase* h = static_cast<HarnessBase*>(hv);
    const int N = h->grid.numberOfCells();
    std::vector<double> s(sat, sat + N), pc;
    h->capPressures(s, pc);
    std::copy(pc.begin(), pc.end(), out);
}

// findPeriodicPartners + match (BoundaryPeriodicity.hpp:86-177, .cpp:25-49) on a flat list of boundary faces.
// Returns 0, or 1 when the reference throws.
int ref_find_periodic_partners(int n, const double* centroid, const double* area, const int* is_periodic6, double tol,
                               int* canon_pos, int* partner, double* side_areas6)
{
    BFaceView view = { centroid, area, n };
    std::vector<Opm::BoundaryFaceInfo> info;
    std::array<double, 6> sa;
    std::array<bool, 6> per;
    for (int k = 0; k < 6; ++k) per[k] = is_periodic6[k] != 0;
    std::ostringstream captured;
    std::streambuf* old = std::cerr.rdbuf(captured.rdbuf());
    int status = 0;
    try {
        Opm::findPeriodicPartners(info, sa, view, per, tol);
    } catch (const std::exception&) {
        status = 1;
    }
    std::cerr.rdbuf(old);
    if (status) return status;
    for (int k = 0; k < 6; ++k) side_areas6[k] = sa[k];
    for (size_t q = 0; q < info.size(); ++q) {
        canon_pos[info[q].face_index] = info[q].canon_pos;
        partner[info[q].face_index] = info[q].partner_face_index;
    }
    return 0;
}

void ref_cfl_factors(void* hv, double* out3) { static_cast<HarnessBase*>(hv)->cflFactors(out3); }
double ref_cap_pressure(void* hv, int cell, double s) { return static_cast<HarnessBase*>(hv)->capPressure(cell, s); }
void ref_mobility(void* hv, int phase, int cell, double s, double* out9) { static_cast<HarnessBase*>(hv)->mobility(phase, cell, s, out9); }
double ref_frac_flow(void* hv, int cell, double s) { return static_cast<HarnessBase*>(hv)->fracFlow(cell, s); }
const char* ref_last_error(void* hv) { return static_cast<HarnessBase*>(hv)->last_error.c_str(); }

// writeField (SimulatorUtilities.hpp:288-298): returns 1 if the reference threw
int ref_write_field(const double* field, int n, const char* filename)
{
    try {
        Opm::writeField(std::vector<double>(field, field + n), filename);
    } catch (...) {
        return 1;
    }
    return 0;
}

} // extern "C"
