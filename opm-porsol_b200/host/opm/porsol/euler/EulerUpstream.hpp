// Drop-in replacement for opm-porsol's <opm/porsol/euler/EulerUpstream.hpp>.
//
// Same class template, same members and argument meaning as the reference
// (opm/porsol/euler/EulerUpstream.hpp:51-146, EulerUpstream_impl.hpp:59-218), but the work runs on a
// B200 through the C ABI of include/euler_b200.h:
//
//   init(param)                     -> eu_set_params            (same parameter keys and defaults)
//   initObj(grid, resprop, bc)      -> walks the GridInterface once, in the reference's cell / face
//                                      iteration order, flattens grid + properties + boundary conditions
//                                      and uploads them (eu_grid_begin/append/end, eu_set_fluid)
//   transportSolve(sat, time, g,    -> gathers pressure_sol.outflux(f) for every half-face, calls
//                  pressure_sol,       eu_transport_solve, throws the reference's exceptions
//                  injection_rates)
//
// Put opm-porsol_b200/host in front of the opm-porsol include path and SimulatorTraits.hpp:89-101,
// SimulatorBase.hpp:213 and examples/SimulatorTester.hpp:83-85 pick this class up unchanged
// (INTEGRATION.md).  Define EULER_B200_KEEP_REFERENCE to keep the name Opm::EulerUpstream free (the class
// is then only reachable as Opm::b200::EulerUpstream), e.g. to run both side by side in a test.
//
// Like the reference (EulerUpstreamResidual.hpp:106-108) the object keeps pointers to grid, properties
// and boundary conditions; call initObj again after changing any of them.
#ifndef OPM_B200_EULERUPSTREAM_HEADER
#define OPM_B200_EULERUPSTREAM_HEADER

#include <opm/common/ErrorMacros.hpp>
#include <opm/core/utility/parameters/ParameterGroup.hpp>
#include <opm/core/utility/SparseVector.hpp>

#include <opm/porsol/euler/b200/FluidExtractor.hpp>

#include <euler_b200.h>

#include <iostream>
#include <stdexcept>
#include <string>
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

    template <class GridInterface, class ReservoirProperties, class BoundaryConditions>
    class EulerUpstream
    {
    public:
        EulerUpstream() : pgrid_(0), prp_(0), pbc_(0), handle_(0), device_(0), mode_(EU_MODE_AUTO) { eu_default_params(&par_); }
        EulerUpstream(const GridInterface& grid, const ReservoirProperties& resprop, const BoundaryConditions& boundary)
            : pgrid_(0), prp_(0), pbc_(0), handle_(0), device_(0), mode_(EU_MODE_AUTO)
        {
            eu_default_params(&par_);
            initObj(grid, resprop, boundary);
        }
        ~EulerUpstream() { if (handle_) eu_destroy(handle_); }
        EulerUpstream(const EulerUpstream&) = delete;
        EulerUpstream& operator=(const EulerUpstream&) = delete;

        /// Same keys and defaults as the reference (EulerUpstream_impl.hpp:95-108), plus two optional ones:
        /// b200_device (CUDA ordinal, default 0) and b200_mode (0 auto, 1 strict = bit-identical, 2 fast).
        void init(const Opm::parameter::ParameterGroup& param)
        {
            par_.courant_number = param.getDefault("courant_number", par_.courant_number);
            par_.method_viscous = param.getDefault("method_viscous", par_.method_viscous != 0);
            par_.method_gravity = param.getDefault("method_gravity", par_.method_gravity != 0);
            par_.method_capillary = param.getDefault("method_capillary", par_.method_capillary != 0);
            par_.use_cfl_viscous = param.getDefault("use_cfl_viscous", par_.use_cfl_viscous != 0);
            par_.use_cfl_gravity = param.getDefault("use_cfl_gravity", par_.use_cfl_gravity != 0);
            par_.use_cfl_capillary = param.getDefault("use_cfl_capillary", par_.use_cfl_capillary != 0);
            par_.minimum_small_steps = param.getDefault("minimum_small_steps", par_.minimum_small_steps);
            par_.maximum_small_steps = param.getDefault("maximum_small_steps", par_.maximum_small_steps);
            par_.check_sat = param.getDefault("check_sat", par_.check_sat != 0);
            par_.clamp_sat = param.getDefault("clamp_sat", par_.clamp_sat != 0);
            device_ = param.getDefault("b200_device", device_);
            mode_ = param.getDefault("b200_mode", mode_);
            if (handle_) eu_set_params(handle_, &par_);
        }

        void init(const Opm::parameter::ParameterGroup& param, const GridInterface& grid,
                  const ReservoirProperties& resprop, const BoundaryConditions& boundary)
        {
            init(param);
            initObj(grid, resprop, boundary);
        }

        void initObj(const GridInterface& grid, const ReservoirProperties& resprop, const BoundaryConditions& boundary)
        {
            pgrid_ = &grid;
            prp_ = &resprop;
            pbc_ = &boundary;
            if (handle_) { eu_destroy(handle_); handle_ = 0; }
            eu_config cfg;
            cfg.abi_version = EU_ABI_VERSION;
            cfg.device = device_;
            cfg.mode = mode_;
            cfg.rank = 0; cfg.world_size = 1; cfg.own_begin = 0; cfg.own_end = grid.numberOfCells();
            if (eu_create(&cfg, &handle_) != EU_OK) {
                OPM_THROW(std::runtime_error, "EulerUpstream (B200): " << eu_last_error(0));
            }
            check(eu_set_params(handle_, &par_));
            flattenAndUpload();
        }

        void display()
        {
            using namespace std;
            cout << endl;
            cout << "Displaying some members of EulerUpstream" << endl;
            cout << endl;
            cout << "courant_number = " << par_.courant_number << endl;
        }

        void setCourantNumber(double cn)
        {
            par_.courant_number = cn;
            if (handle_) eu_set_params(handle_, &par_);
        }

        /// Report of the last transportSolve (step count, retries, CFL times, device time).
        const eu_report& lastReport() const { return report_; }

        template <class PressureSolution>
        void transportSolve(std::vector<double>& saturation, const double time,
                            const typename GridInterface::Vector& gravity, const PressureSolution& pressure_sol,
                            const Opm::SparseVector<double>& injection_rates) const
        {
            if (!handle_) OPM_THROW(std::runtime_error, "EulerUpstream (B200): initObj() has not been called");
            if (int(saturation.size()) != num_cells_) OPM_THROW(std::runtime_error, "saturation has the wrong size");
            gatherFluxes(pressure_sol, std::integral_constant<bool, detail::HasFlatOutflux<PressureSolution>::value>());
            std::vector<int> src_cell;
            std::vector<double> src_rate;
            for (int i = 0; i < injection_rates.nonzeroSize(); ++i) {
                src_cell.push_back(injection_rates.nonzeroIndex(i));
                src_rate.push_back(injection_rates.nonzeroElement(i));
            }
            const double g[3] = { gravity[0], gravity[1], gravity[2] };
            const int rc = eu_transport_solve(handle_, saturation.data(), time, g, hf_flux_.data(), int(src_cell.size()),
                                              src_cell.data(), src_rate.data(), &report_);
            for (int r = 1; r < report_.attempts; ++r) {
                OPM_MESSAGE("Warning: Transport failed, retrying with more steps.");
            }
            if (rc == EU_ERR_SAT_RANGE) {
                OPM_THROW(std::runtime_error, "Saturation out of range in EulerUpstream: Cell " << report_.bad_cell
                          << "   sat " << report_.bad_value);
            }
            if (rc == EU_ERR_CFL_ZERO) {
                OPM_THROW(std::runtime_error, "Cfl computation gave dt = 0.0");
            }
            check(rc);
        }

    protected:
        typedef typename GridInterface::CellIterator CIt;
        typedef typename CIt::FaceIterator FIt;
        typedef typename FIt::Vector Vector;

        void check(int rc) const
        {
            if (rc != EU_OK) OPM_THROW(std::runtime_error, "EulerUpstream (B200): " << eu_last_error(handle_));
        }

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
                    OPM_THROW(std::runtime_error, "EulerUpstream (B200): cell index must equal iteration order");
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
            hf_flux_.assign(size_t(H), 0.0);
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

        template <class PressureSolution>
        void gatherFluxes(const PressureSolution& ps, std::true_type) const
        {
            const int H = int(hf_flux_.size());
            for (int hf = 0; hf < H; ++hf) hf_flux_[hf] = detail::flatOutflux(ps, hf, 0);
        }
        template <class PressureSolution>
        void gatherFluxes(const PressureSolution& ps, std::false_type) const
        {
            size_t hf = 0;
            for (CIt c = pgrid_->cellbegin(); c != pgrid_->cellend(); ++c) {
                for (FIt f = c->facebegin(); f != c->faceend(); ++f) hf_flux_[hf++] = ps.outflux(f);
            }
        }

        const GridInterface* pgrid_;
        const ReservoirProperties* prp_;
        const BoundaryConditions* pbc_;
        mutable eu_handle handle_;
        eu_params par_;
        int device_;
        int mode_;
        int num_cells_;
        mutable std::vector<double> hf_flux_;
        mutable eu_report report_;
    };

} // namespace b200

#ifndef EULER_B200_KEEP_REFERENCE
    using b200::EulerUpstream;
#endif

} // namespace Opm

#endif // OPM_B200_EULERUPSTREAM_HEADER
