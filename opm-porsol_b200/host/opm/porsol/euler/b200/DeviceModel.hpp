// Shared by the drop-in headers of this directory (EulerUpstream.hpp, EulerUpstreamResidual.hpp): owns the
// eu_handle of the C ABI (include/euler_b200.h) and does the one walk over the caller's GridInterface /
// ReservoirProperties / BoundaryConditions objects that flattens them, in the reference's cell and face
// iteration order (EulerUpstream_impl.hpp:124, EulerUpstreamResidual_impl.hpp:407-421), into the chunks
// eu_grid_append takes.  Also gathers pressure_sol.outflux(f) for every half-face
// (IncompFlowSolverHybrid.hpp:426-433) and turns an Opm::SparseVector into (cell, rate) pairs.
#ifndef OPM_B200_DEVICEMODEL_HEADER
#define OPM_B200_DEVICEMODEL_HEADER

#include <opm/common/ErrorMacros.hpp>
#include <opm/core/utility/SparseVector.hpp>

#include <opm/porsol/euler/b200/FluidExtractor.hpp>

#include <euler_b200.h>

#include <algorithm>
#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace Opm {
namespace b200 {

    namespace detail {
        // pressure_sol.outflux(int half_face) is the flat accessor of IncompFlowSolverHybrid's FlowSolution
        // (IncompFlowSolverHybrid.hpp:430-433); use it when present, else outflux(face iterator) (:426-429).
        template <class PS>
        auto flatOutflux(const PS& ps, int hf, int) -> decltype(double(ps.outflux(hf))) { return ps.outflux(hf); }
        template <class PS>
        struct HasFlatOutflux {
            template <class T> static auto test(int) -> decltype(std::declval<const T&>().outflux(int(0)), std::true_type());
            template <class T> static std::false_type test(...);
            static const bool value = decltype(test<PS>(0))::value;
        };
    }

    /// Flat array of half-face fluxes in page-locked host memory (eu_host_alloc), so that the copy to the device inside
    /// eu_transport_solve runs at the PCIe rate; falls back to pageable memory if the allocation is refused.
    class FluxBuffer
    {
    public:
        FluxBuffer() : p_(0), n_(0), pinned_(false) {}
        ~FluxBuffer() { release(); }
        FluxBuffer(const FluxBuffer&) = delete;
        FluxBuffer& operator=(const FluxBuffer&) = delete;
        void assign(std::size_t n)
        {
            release();
            p_ = static_cast<double*>(eu_host_alloc(static_cast<unsigned long long>(n)*sizeof(double)));
            pinned_ = p_ != 0;
            if (!p_) p_ = new double[n ? n : 1];
            n_ = n;
            std::fill(p_, p_ + n_, 0.0);
        }
        double* data() { return p_; }
        const double* data() const { return p_; }
        std::size_t size() const { return n_; }
        double& operator[](std::size_t i) { return p_[i]; }
    private:
        void release()
        {
            if (p_) { if (pinned_) eu_host_free(p_); else delete[] p_; }
            p_ = 0; n_ = 0; pinned_ = false;
        }
        double* p_;
        std::size_t n_;
        bool pinned_;
    };

    template <class GridInterface, class ReservoirProperties, class BoundaryConditions>
    class DeviceModel
    {
    public:
        typedef typename GridInterface::CellIterator CIt;
        typedef typename CIt::FaceIterator FIt;
        typedef typename FIt::Vector Vector;

        DeviceModel() : pgrid_(0), prp_(0), pbc_(0), num_cells_(0) {}
        ~DeviceModel() { release(); }
        DeviceModel(const DeviceModel&) = delete;
        DeviceModel& operator=(const DeviceModel&) = delete;

        void release()
        {
            for (size_t r = 0; r < ranks_.size(); ++r) if (ranks_[r].h) eu_destroy(ranks_[r].h);
            ranks_.clear();
        }
        bool ready() const { return !ranks_.empty(); }
        /// number of devices the grid is spread over
        int numDevices() const { return int(ranks_.size()); }
        /// the device solver (C ABI handle); with several devices the one of the first slab -- the operator and
        /// diagnostics entry points of euler_b200.h are single-device and refuse a decomposed solver themselves
        eu_handle handle() const { return ranks_.empty() ? 0 : ranks_[0].h; }
        int numCells() const { return num_cells_; }
        const GridInterface& grid() const { return *pgrid_; }
        const ReservoirProperties& reservoirProperties() const { return *prp_; }
        const BoundaryConditions& boundaryConditions() const { return *pbc_; }
        /// pressure_sol.outflux of every half-face, in upload order (filled by gatherFluxes)
        const FluxBuffer& fluxes() const { return hf_flux_; }

        void check(int rc, const char* who) const
        {
            if (rc != EU_OK) OPM_THROW(std::runtime_error, who << " (B200): " << eu_last_error(handle()));
        }

        /// (Re)creates the device solver for these objects.  Like the reference (EulerUpstreamResidual.hpp:106-108)
        /// pointers to the three objects are kept; call again after changing any of them.
        void create(int device, int mode, const eu_params& par, const GridInterface& grid,
                    const ReservoirProperties& resprop, const BoundaryConditions& boundary, const char* who)
        {
            create(std::vector<int>(1, device), mode, par, grid, resprop, boundary, who);
        }

        /// Several devices of this process (parameter b200_devices=0,1,..): the cells are split into contiguous index
        /// ranges (z-slabs for the natural ordering of Cartesian and corner-point grids), one solver per device, each
        /// holding the cells of its neighbours that its faces reach as ghosts; the solvers exchange ghost saturations
        /// device to device (peer-to-peer stores) and transportSolve drives them from one host thread each.  The result
        /// does not depend on the number of devices (faces on a slab boundary are evaluated on both sides from identical
        /// operands): bit-identical in STRICT mode.  The caller stays single-process, as in the reference
        /// (SimulatorBase.hpp:204-214).
        void create(const std::vector<int>& devices, int mode, const eu_params& par, const GridInterface& grid,
                    const ReservoirProperties& resprop, const BoundaryConditions& boundary, const char* who)
        {
            pgrid_ = &grid;
            prp_ = &resprop;
            pbc_ = &boundary;
            who_ = who;
            release();
            if (devices.empty()) OPM_THROW(std::runtime_error, who << " (B200): empty device list");
            mode_ = mode;
            par_ = par;
            devices_ = devices;
            flattenAndUpload();
        }

        void setParams(const eu_params& par)
        {
            par_ = par;
            for (size_t r = 0; r < ranks_.size(); ++r) eu_set_params(ranks_[r].h, &par);
        }

        /// eu_transport_solve on one device, or on every slab from a host thread each
        int transportSolve(std::vector<double>& saturation, double time, const double gravity[3], const std::vector<int>& src_cell,
                           const std::vector<double>& src_rate, eu_report* report)
        {
            if (ranks_.size() == 1) {
                return eu_transport_solve(ranks_[0].h, saturation.data(), time, gravity, hf_flux_.data(), int(src_cell.size()),
                                          src_cell.data(), src_rate.data(), report);
            }
            const size_t W = ranks_.size();
            std::vector<int> rcs(W, EU_OK);
            std::vector<eu_report> reps(W);
            std::vector<std::thread> pool;
            for (size_t r = 0; r < W; ++r) {
                pool.emplace_back([&, r]() {
                    Rank& k = ranks_[r];
                    // private copy of the slab's cells: ghost entries are inputs, own entries come back
                    k.sat.assign(saturation.begin() + k.lo, saturation.begin() + k.hi);
                    rcs[r] = eu_transport_solve(k.h, k.sat.data(), time, gravity, hf_flux_.data() + k.hf_lo, int(src_cell.size()),
                                                src_cell.data(), src_rate.data(), &reps[r]);
                });
            }
            for (std::thread& th : pool) th.join();
            int rc = EU_OK;
            for (size_t r = 0; r < W; ++r) {
                const Rank& k = ranks_[r];
                std::copy(k.sat.begin() + (k.own_lo - k.lo), k.sat.begin() + (k.own_hi - k.lo), saturation.begin() + k.own_lo);
                if (rcs[r] != EU_OK && rc == EU_OK) { rc = rcs[r]; *report = reps[r]; failed_rank_ = int(r); }
            }
            if (rc == EU_OK) {
                *report = reps[0];
                for (size_t r = 1; r < W; ++r) {
                    report->device_ms = std::max(report->device_ms, reps[r].device_ms);
                    report->kernel_launches += reps[r].kernel_launches;
                }
            } else {
                // the reference stops at the lowest failing cell of the first failing substep: every rank reports its own
                // lowest one for that substep (or none)
                int best = -1;
                for (size_t r = 0; r < W; ++r)
                    if (rcs[r] == EU_ERR_SAT_RANGE && reps[r].bad_cell >= 0 && (best < 0 || reps[r].bad_cell < reps[size_t(best)].bad_cell)) best = int(r);
                if (best >= 0) { *report = reps[size_t(best)]; failed_rank_ = best; }
            }
            return rc;
        }
        /// eu_last_error of the rank whose transportSolve failed last
        const char* lastError() const { return eu_last_error(ranks_.empty() ? 0 : ranks_[size_t(failed_rank_)].h); }

        template <class PressureSolution>
        void gatherFluxes(const PressureSolution& ps)
        {
            gatherFluxes(ps, std::integral_constant<bool, detail::HasFlatOutflux<PressureSolution>::value>());
        }

        static void sources(const Opm::SparseVector<double>& injection_rates, std::vector<int>& cell, std::vector<double>& rate)
        {
            cell.clear();
            rate.clear();
            for (int i = 0; i < injection_rates.nonzeroSize(); ++i) {
                cell.push_back(injection_rates.nonzeroIndex(i));
                rate.push_back(injection_rates.nonzeroElement(i));
            }
        }

    private:
        void check(int rc) const { check(rc, who_.c_str()); }

        // one slab: a solver on one device holding the cells [lo, hi) (own [own_lo, own_hi) plus ghosts)
        struct Rank {
            eu_handle h;
            int device, lo, hi, own_lo, own_hi;
            long long hf_lo, hf_hi;
            std::vector<double> sat;
            // chunk under construction
            std::vector<int> hf_count, hf_nbr, bnd_hf, bnd_kind, bnd_pcell, bnd_pface, rock;
            std::vector<double> area, normal, centroid, bnd_sat, vol, ccent, poro, perm;
            int chunk_first;
            Rank() : h(0), device(0), lo(0), hi(0), own_lo(0), own_hi(0), hf_lo(0), hf_hi(0), chunk_first(0) {}
            void clearChunk()
            {
                hf_count.clear(); hf_nbr.clear(); bnd_hf.clear(); bnd_kind.clear(); bnd_pcell.clear(); bnd_pface.clear();
                rock.clear(); area.clear(); normal.clear(); centroid.clear(); bnd_sat.clear(); vol.clear(); ccent.clear();
                poro.clear(); perm.clear();
            }
        };

        // min / max over the slabs of this process: every rank's thread calls in with the same sequence of reductions
        struct Reducer {
            std::mutex m;
            std::condition_variable cv;
            int world, arrived;
            unsigned long long gen;
            double acc[8], result[8];
            Reducer() : world(1), arrived(0), gen(0) {}
            static void call(void* user, double* values, int n, int op)
            {
                Reducer* R = static_cast<Reducer*>(user);
                std::unique_lock<std::mutex> lk(R->m);
                if (R->arrived == 0) for (int i = 0; i < n && i < 8; ++i) R->acc[i] = values[i];
                else for (int i = 0; i < n && i < 8; ++i) R->acc[i] = op == 0 ? std::min(R->acc[i], values[i]) : std::max(R->acc[i], values[i]);
                if (++R->arrived == R->world) {
                    for (int i = 0; i < n && i < 8; ++i) R->result[i] = R->acc[i];
                    R->arrived = 0;
                    ++R->gen;
                    R->cv.notify_all();
                } else {
                    const unsigned long long g = R->gen;
                    R->cv.wait(lk, [R, g]() { return R->gen != g; });
                }
                for (int i = 0; i < n && i < 8; ++i) values[i] = R->result[i];
            }
        };

        void flushChunk(Rank& k)
        {
            if (k.hf_count.empty()) return;
            eu_grid_chunk ch;
            ch.first_cell = k.chunk_first; ch.n_cells = int(k.hf_count.size());
            ch.hf_count = k.hf_count.data(); ch.hf_neighbour = k.hf_nbr.data();
            ch.hf_area = k.area.data(); ch.hf_normal = k.normal.data(); ch.hf_centroid = k.centroid.data();
            ch.n_bnd = int(k.bnd_hf.size());
            ch.bnd_hf = k.bnd_hf.data(); ch.bnd_kind = k.bnd_kind.data(); ch.bnd_sat = k.bnd_sat.data();
            ch.bnd_partner_cell = k.bnd_pcell.data(); ch.bnd_partner_face = k.bnd_pface.data();
            ch.cell_volume = k.vol.data(); ch.cell_centroid = k.ccent.data();
            ch.porosity = k.poro.data(); ch.permeability = k.perm.data();
            ch.rock_id = n_rocks_ > 0 ? k.rock.data() : 0;
            if (eu_grid_append(k.h, &ch) != EU_OK) OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(k.h));
            k.chunk_first += ch.n_cells;
            k.clearChunk();
        }

        // One walk in the reference's order (EulerUpstream_impl.hpp:124, EulerUpstreamResidual_impl.hpp:407-421),
        // uploaded in chunks so that no second copy of a large grid is ever held on the host.
        void flattenAndUpload()
        {
            const GridInterface& g = *pgrid_;
            const ReservoirProperties& rp = *prp_;
            const BoundaryConditions& bc = *pbc_;
            num_cells_ = g.numberOfCells();
            const int W = int(devices_.size());
            // pass 1: counts, cell numbering check, periodic boundary id -> (cell, local face); for a decomposition also the
            // half-face offset of every cell, how far a face reaches in cell index and the most common large offset (the
            // plane size of a natural ordering), which the slab boundaries are aligned to
            long long H = 0;
            int pos = 0, maxbid = 0, reach = 0;
            std::vector<long long> hf_off;
            std::map<int, long long> offsets;
            if (W > 1) hf_off.reserve(size_t(num_cells_) + 1);
            for (CIt c = g.cellbegin(); c != g.cellend(); ++c, ++pos) {
                if (c->index() != pos) {
                    OPM_THROW(std::runtime_error, who_ << " (B200): cell index must equal iteration order");
                }
                if (W > 1) hf_off.push_back(H);
                for (FIt f = c->facebegin(); f != c->faceend(); ++f) {
                    ++H;
                    if (f->boundary()) maxbid = std::max(maxbid, int(f->boundaryId()));
                    else if (W > 1) {
                        const int d = f->neighbourCellIndex() - pos;
                        reach = std::max(reach, d < 0 ? -d : d);
                        if (d > 0) ++offsets[d];
                    }
                }
            }
            if (W > 1) hf_off.push_back(H);
            std::vector<std::pair<int, int> > bid_to_face(maxbid + 1, std::make_pair(-1, -1));
            for (CIt c = g.cellbegin(); c != g.cellend(); ++c) {
                for (FIt f = c->facebegin(); f != c->faceend(); ++f) {
                    if (f->boundary() && bc.satCond(*f).isPeriodic()) {
                        bid_to_face[f->boundaryId()] = std::make_pair(int(c->index()), int(f->localIndex()));
                    }
                }
            }
            hf_flux_.assign(size_t(H));
            // slabs
            ranks_.assign(size_t(W), Rank());
            int plane = 0;
            for (std::map<int, long long>::const_iterator it = offsets.begin(); it != offsets.end(); ++it)
                if (it->second*4 >= num_cells_ && it->first > plane) plane = it->first;
            if (plane <= 0 || num_cells_ % plane != 0) plane = 1;
            if (W > 1) {
                // a periodic partner is a neighbour like any other: the slabs' ghost range must reach it
                for (size_t b = 0; b < bid_to_face.size(); ++b) {
                    if (bid_to_face[b].first >= 0) {
                        const std::pair<int, int>& p = bid_to_face[bc.getPeriodicPartner(int(b))];
                        if (p.first >= 0) reach = std::max(reach, std::abs(p.first - bid_to_face[b].first));
                    }
                }
            }
            const int ext = ((reach + plane - 1)/plane)*plane;
            for (int r = 0; r < W; ++r) {
                Rank& k = ranks_[size_t(r)];
                k.device = devices_[size_t(r)];
                const long long units = num_cells_/plane;
                k.own_lo = int(units*r/W)*plane;
                k.own_hi = int(units*(r + 1)/W)*plane;
                if (r == W - 1) k.own_hi = num_cells_;
                if (k.own_hi <= k.own_lo) OPM_THROW(std::runtime_error, who_ << " (B200): more devices than slabs of the grid");
                k.lo = W > 1 ? std::max(0, k.own_lo - ext) : 0;
                k.hi = W > 1 ? std::min(num_cells_, k.own_hi + ext) : num_cells_;
                k.hf_lo = W > 1 ? hf_off[size_t(k.lo)] : 0;
                k.hf_hi = W > 1 ? hf_off[size_t(k.hi)] : H;
                k.chunk_first = k.lo;
                eu_config cfg;
                cfg.abi_version = EU_ABI_VERSION;
                cfg.device = k.device;
                cfg.mode = mode_;
                cfg.rank = r; cfg.world_size = W; cfg.own_begin = k.own_lo; cfg.own_end = k.own_hi;
                if (eu_create(&cfg, &k.h) != EU_OK) {
                    OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(0));
                }
                if (eu_set_params(k.h, &par_) != EU_OK) OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(k.h));
                if (eu_grid_begin(k.h, num_cells_, k.hi - k.lo, k.hf_hi - k.hf_lo) != EU_OK)
                    OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(k.h));
            }
            // fluid: viscosities, densities, CFL factors, rock tables
            FluidDescription fd;
            FluidExtractor<ReservoirProperties>::extract(rp, num_cells_, fd);
            eu_fluid fl;
            fd.fill(fl);
            n_rocks_ = fd.n_rocks;
            for (int r = 0; r < W; ++r)
                if (eu_set_fluid(ranks_[size_t(r)].h, &fl) != EU_OK) OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(ranks_[size_t(r)].h));
            // pass 2: every cell goes to the slab(s) that hold it, in chunks
            const int chunk_cells = 1 << 18;
            int first_active = 0;              // slabs below this one are complete
            for (CIt c = g.cellbegin(); c != g.cellend(); ++c) {
                const int ci = c->index();
                while (first_active < W && ci >= ranks_[size_t(first_active)].hi) { flushChunk(ranks_[size_t(first_active)]); ++first_active; }
                for (int r = first_active; r < W && ranks_[size_t(r)].lo <= ci; ++r) {
                    Rank& k = ranks_[size_t(r)];
                    if (ci >= k.hi) continue;
                    int cnt = 0;
                    for (FIt f = c->facebegin(); f != c->faceend(); ++f, ++cnt) {
                        const Vector nrm = f->normal();
                        const Vector fc = f->centroid();
                        k.area.push_back(f->area());
                        for (int d = 0; d < 3; ++d) { k.normal.push_back(nrm[d]); k.centroid.push_back(fc[d]); }
                        if (f->boundary()) {
                            k.hf_nbr.push_back(-1);
                            k.bnd_hf.push_back(int(k.hf_nbr.size()) - 1);
                            if (bc.satCond(*f).isPeriodic()) {
                                const std::pair<int, int>& p = bid_to_face[bc.getPeriodicPartner(f->boundaryId())];
                                if (p.first < 0) OPM_THROW(std::runtime_error, "periodic face without a partner face");
                                k.bnd_kind.push_back(EU_HF_PERIODIC);
                                k.bnd_sat.push_back(0.0);
                                k.bnd_pcell.push_back(p.first);
                                k.bnd_pface.push_back(p.second);
                            } else {
                                k.bnd_kind.push_back(EU_HF_DIRICHLET);
                                k.bnd_sat.push_back(bc.satCond(*f).saturation());
                                k.bnd_pcell.push_back(-1);
                                k.bnd_pface.push_back(-1);
                            }
                        } else {
                            k.hf_nbr.push_back(f->neighbourCellIndex());
                        }
                    }
                    k.hf_count.push_back(cnt);
                    k.vol.push_back(c->volume());
                    const Vector cc = c->centroid();
                    for (int d = 0; d < 3; ++d) k.ccent.push_back(cc[d]);
                    k.poro.push_back(rp.porosity(ci));
                    typename ReservoirProperties::PermTensor K = rp.permeability(ci);
                    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) k.perm.push_back(K(i, j));
                    k.rock.push_back(fd.rockOfCell(ci));
                    if (int(k.hf_count.size()) >= chunk_cells) flushChunk(k);
                }
            }
            for (int r = 0; r < W; ++r) {
                flushChunk(ranks_[size_t(r)]);
                if (eu_grid_end(ranks_[size_t(r)].h) != EU_OK) OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(ranks_[size_t(r)].h));
            }
            if (W > 1) {
                // ghost exchange between the devices: one opaque blob per slab, handed to every slab; the reductions
                // (CFL times, range-check flag) go through this object
                reducer_.world = W;
                std::vector<std::vector<char> > blobs(static_cast<size_t>(W));
                std::vector<const void*> ptrs(static_cast<size_t>(W));
                std::vector<int> sizes(static_cast<size_t>(W));
                for (int r = 0; r < W; ++r) {
                    blobs[size_t(r)].resize(size_t(eu_comm_blob_size(ranks_[size_t(r)].h)));
                    if (eu_comm_export(ranks_[size_t(r)].h, blobs[size_t(r)].data()) != EU_OK)
                        OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(ranks_[size_t(r)].h));
                    ptrs[size_t(r)] = blobs[size_t(r)].data();
                    sizes[size_t(r)] = int(blobs[size_t(r)].size());
                }
                for (int r = 0; r < W; ++r) {
                    if (eu_comm_connect(ranks_[size_t(r)].h, W, ptrs.data(), sizes.data()) != EU_OK)
                        OPM_THROW(std::runtime_error, who_ << " (B200): " << eu_last_error(ranks_[size_t(r)].h));
                    eu_comm_set_allreduce(ranks_[size_t(r)].h, &Reducer::call, &reducer_);
                }
            }
        }

        // flat accessor: a const read per half-face, spread over the host cores for large grids
        template <class PressureSolution>
        void gatherFluxes(const PressureSolution& ps, std::true_type)
        {
            const int H = int(hf_flux_.size());
            double* out = hf_flux_.data();
            unsigned nt = H >= (1 << 20) ? std::min(std::thread::hardware_concurrency(), 32u) : 1u;
            if (nt <= 1) {
                for (int hf = 0; hf < H; ++hf) out[hf] = detail::flatOutflux(ps, hf, 0);
                return;
            }
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; ++t) {
                const int lo = int((long long)H*t/nt), hi = int((long long)H*(t + 1)/nt);
                pool.emplace_back([&ps, out, lo, hi]() { for (int hf = lo; hf < hi; ++hf) out[hf] = detail::flatOutflux(ps, hf, 0); });
            }
            for (std::thread& th : pool) th.join();
        }
        template <class PressureSolution>
        void gatherFluxes(const PressureSolution& ps, std::false_type)
        {
            size_t hf = 0;
            for (CIt c = pgrid_->cellbegin(); c != pgrid_->cellend(); ++c) {
                for (FIt f = c->facebegin(); f != c->faceend(); ++f) hf_flux_[hf++] = ps.outflux(f);
            }
        }

        const GridInterface* pgrid_;
        const ReservoirProperties* prp_;
        const BoundaryConditions* pbc_;
        std::vector<Rank> ranks_;
        std::vector<int> devices_;
        Reducer reducer_;
        eu_params par_;
        int mode_ = 0;
        int n_rocks_ = 0;
        int failed_rank_ = 0;
        int num_cells_;
        std::string who_;
        FluxBuffer hf_flux_;
    };

} // namespace b200
} // namespace Opm

#endif // OPM_B200_DEVICEMODEL_HEADER
