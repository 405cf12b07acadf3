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
#include <cstddef>
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

        DeviceModel() : pgrid_(0), prp_(0), pbc_(0), handle_(0), num_cells_(0) {}
        ~DeviceModel() { release(); }
        DeviceModel(const DeviceModel&) = delete;
        DeviceModel& operator=(const DeviceModel&) = delete;

        void release() { if (handle_) { eu_destroy(handle_); handle_ = 0; } }
        bool ready() const { return handle_ != 0; }
        eu_handle handle() const { return handle_; }
        int numCells() const { return num_cells_; }
        const GridInterface& grid() const { return *pgrid_; }
        const ReservoirProperties& reservoirProperties() const { return *prp_; }
        const BoundaryConditions& boundaryConditions() const { return *pbc_; }
        /// pressure_sol.outflux of every half-face, in upload order (filled by gatherFluxes)
        const FluxBuffer& fluxes() const { return hf_flux_; }

        void check(int rc, const char* who) const
        {
            if (rc != EU_OK) OPM_THROW(std::runtime_error, who << " (B200): " << eu_last_error(handle_));
        }

        /// (Re)creates the device solver for these objects.  Like the reference (EulerUpstreamResidual.hpp:106-108)
        /// pointers to the three objects are kept; call again after changing any of them.
        void create(int device, int mode, const eu_params& par, const GridInterface& grid,
                    const ReservoirProperties& resprop, const BoundaryConditions& boundary, const char* who)
        {
            pgrid_ = &grid;
            prp_ = &resprop;
            pbc_ = &boundary;
            who_ = who;
            release();
            eu_config cfg;
            cfg.abi_version = EU_ABI_VERSION;
            cfg.device = device;
            cfg.mode = mode;
            cfg.rank = 0; cfg.world_size = 1; cfg.own_begin = 0; cfg.own_end = grid.numberOfCells();
            if (eu_create(&cfg, &handle_) != EU_OK) {
                OPM_THROW(std::runtime_error, who << " (B200): " << eu_last_error(0));
            }
            check(eu_set_params(handle_, &par));
            flattenAndUpload();
        }

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

        // One walk in the reference's order (EulerUpstream_impl.hpp:124, EulerUpstreamResidual_impl.hpp:407-421),
        // uploaded in chunks so that no second copy of a large grid is ever held on the host.
        void flattenAndUpload()
        {
            const GridInterface& g = *pgrid_;
            const ReservoirProperties& rp = *prp_;
            const BoundaryConditions& bc = *pbc_;
            num_cells_ = g.numberOfCells();
            // pass 1: counts, cell numbering check, periodic boundary id -> (cell, local face)
            long long H = 0;
            int pos = 0, maxbid = 0;
            for (CIt c = g.cellbegin(); c != g.cellend(); ++c, ++pos) {
                if (c->index() != pos) {
                    OPM_THROW(std::runtime_error, who_ << " (B200): cell index must equal iteration order");
                }
                for (FIt f = c->facebegin(); f != c->faceend(); ++f) {
                    ++H;
                    if (f->boundary()) maxbid = std::max(maxbid, int(f->boundaryId()));
                }
            }
            std::vector<std::pair<int, int> > bid_to_face(maxbid + 1, std::make_pair(-1, -1));
            for (CIt c = g.cellbegin(); c != g.cellend(); ++c) {
                for (FIt f = c->facebegin(); f != c->faceend(); ++f) {
                    if (f->boundary() && bc.satCond(*f).isPeriodic()) {
                        bid_to_face[f->boundaryId()] = std::make_pair(int(c->index()), int(f->localIndex()));
                    }
                }
            }
            hf_flux_.assign(size_t(H));
            check(eu_grid_begin(handle_, num_cells_, num_cells_, H));
            // fluid: viscosities, densities, CFL factors, rock tables
            FluidDescription fd;
            FluidExtractor<ReservoirProperties>::extract(rp, num_cells_, fd);
            eu_fluid fl;
            fd.fill(fl);
            check(eu_set_fluid(handle_, &fl));
            // pass 2: chunks
            const int chunk_cells = 1 << 18;
            std::vector<int> hf_count, hf_nbr, bnd_hf, bnd_kind, bnd_pcell, bnd_pface, rock;
            std::vector<double> area, normal, centroid, bnd_sat, vol, ccent, poro, perm;
            CIt c = g.cellbegin();
            int first = 0;
            while (c != g.cellend()) {
                hf_count.clear(); hf_nbr.clear(); bnd_hf.clear(); bnd_kind.clear(); bnd_pcell.clear(); bnd_pface.clear();
                rock.clear(); area.clear(); normal.clear(); centroid.clear(); bnd_sat.clear(); vol.clear(); ccent.clear();
                poro.clear(); perm.clear();
                int n = 0;
                for (; c != g.cellend() && n < chunk_cells; ++c, ++n) {
                    const int ci = c->index();
                    int cnt = 0;
                    for (FIt f = c->facebegin(); f != c->faceend(); ++f, ++cnt) {
                        const Vector nrm = f->normal();
                        const Vector fc = f->centroid();
                        area.push_back(f->area());
                        for (int d = 0; d < 3; ++d) { normal.push_back(nrm[d]); centroid.push_back(fc[d]); }
                        if (f->boundary()) {
                            hf_nbr.push_back(-1);
                            bnd_hf.push_back(int(hf_nbr.size()) - 1);
                            if (bc.satCond(*f).isPeriodic()) {
                                const std::pair<int, int>& p = bid_to_face[bc.getPeriodicPartner(f->boundaryId())];
                                if (p.first < 0) OPM_THROW(std::runtime_error, "periodic face without a partner face");
                                bnd_kind.push_back(EU_HF_PERIODIC);
                                bnd_sat.push_back(0.0);
                                bnd_pcell.push_back(p.first);
                                bnd_pface.push_back(p.second);
                            } else {
                                bnd_kind.push_back(EU_HF_DIRICHLET);
                                bnd_sat.push_back(bc.satCond(*f).saturation());
                                bnd_pcell.push_back(-1);
                                bnd_pface.push_back(-1);
                            }
                        } else {
                            hf_nbr.push_back(f->neighbourCellIndex());
                        }
                    }
                    hf_count.push_back(cnt);
                    vol.push_back(c->volume());
                    const Vector cc = c->centroid();
                    for (int d = 0; d < 3; ++d) ccent.push_back(cc[d]);
                    poro.push_back(rp.porosity(ci));
                    typename ReservoirProperties::PermTensor K = rp.permeability(ci);
                    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) perm.push_back(K(i, j));
                    rock.push_back(fd.rockOfCell(ci));
                }
                eu_grid_chunk ch;
                ch.first_cell = first; ch.n_cells = n;
                ch.hf_count = hf_count.data(); ch.hf_neighbour = hf_nbr.data();
                ch.hf_area = area.data(); ch.hf_normal = normal.data(); ch.hf_centroid = centroid.data();
                ch.n_bnd = int(bnd_hf.size());
                ch.bnd_hf = bnd_hf.data(); ch.bnd_kind = bnd_kind.data(); ch.bnd_sat = bnd_sat.data();
                ch.bnd_partner_cell = bnd_pcell.data(); ch.bnd_partner_face = bnd_pface.data();
                ch.cell_volume = vol.data(); ch.cell_centroid = ccent.data();
                ch.porosity = poro.data(); ch.permeability = perm.data();
                ch.rock_id = fd.n_rocks > 0 ? rock.data() : 0;
                check(eu_grid_append(handle_, &ch));
                first += n;
            }
            check(eu_grid_end(handle_));
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
        eu_handle handle_;
        int num_cells_;
        std::string who_;
        FluxBuffer hf_flux_;
    };

} // namespace b200
} // namespace Opm

#endif // OPM_B200_DEVICEMODEL_HEADER
