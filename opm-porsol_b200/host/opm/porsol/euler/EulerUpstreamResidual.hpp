// Drop-in replacement for opm-porsol's <opm/porsol/euler/EulerUpstreamResidual.hpp>.
//
// Same class template and public members as the reference (opm/porsol/euler/EulerUpstreamResidual.hpp:57-125,
// _impl.hpp:375-505); the residual is computed on a B200 through eu_compute_residual of include/euler_b200.h.
// The reference uses the class in two places: as a member of EulerUpstream (EulerUpstream.hpp:126-128) and in
// ImplicitCapillarity::transportSolve, which calls computeResidual with method_capillary = false and negates
// the result into injection rates for its capillary pressure solve (ImplicitCapillarity_impl.hpp:176-183).
//
//   initObj(grid, resprop, bc)       -> b200/DeviceModel.hpp: one walk in the reference's cell / face order
//   computeResidual(S, g, flow_sol,  -> gathers flow_sol.outflux(f), eu_compute_residual; sat_delta is cleared and
//       inj, mv, mg, mc, sat_delta)     resized to S.size() like the reference (:483-484)
//   computeCapPressures(S)           -> eu_compute_cap_pressures; the values are kept (capPressures()).  The
//                                       reference's callers must call it before computeResidual(..., mc = true)
//                                       (EulerUpstream_impl.hpp:362-369); here that order is allowed but not needed,
//                                       the device call evaluates pc(S) itself.
//   grid(), reservoirProperties(), boundaryConditions()
//
// Two additions: setDevice(ordinal, mode) before initObj (defaults: device 0, EU_MODE_AUTO), and
// capPressures().  Define EULER_B200_KEEP_REFERENCE to keep the name Opm::EulerUpstreamResidual free.
#ifndef OPM_B200_EULERUPSTREAMRESIDUAL_HEADER
#define OPM_B200_EULERUPSTREAMRESIDUAL_HEADER

#include <opm/common/ErrorMacros.hpp>
#include <opm/core/utility/SparseVector.hpp>

#include <opm/porsol/euler/b200/DeviceModel.hpp>

#include <euler_b200.h>

#include <stdexcept>
#include <vector>

namespace Opm {
namespace b200 {

    template <class GridInterface, class ReservoirProperties, class BoundaryConditions>
    class EulerUpstreamResidual
    {
    public:
        typedef typename GridInterface::CellIterator CIt;
        typedef typename CIt::FaceIterator FIt;
        typedef typename FIt::Vector Vector;
        typedef ReservoirProperties RP;

        EulerUpstreamResidual() : device_(0), mode_(EU_MODE_AUTO) {}
        EulerUpstreamResidual(const GridInterface& grid, const ReservoirProperties& resprop, const BoundaryConditions& boundary)
            : device_(0), mode_(EU_MODE_AUTO)
        {
            initObj(grid, resprop, boundary);
        }
        EulerUpstreamResidual(const EulerUpstreamResidual&) = delete;
        EulerUpstreamResidual& operator=(const EulerUpstreamResidual&) = delete;

        /// CUDA ordinal and arithmetic mode (EU_MODE_*) used by the next initObj.
        void setDevice(int device, int mode) { device_ = device; mode_ = mode; }

        void initObj(const GridInterface& grid, const ReservoirProperties& resprop, const BoundaryConditions& boundary)
        {
            eu_params par;
            eu_default_params(&par);           // the method flags are arguments of computeResidual
            model_.create(device_, mode_, par, grid, resprop, boundary, "EulerUpstreamResidual");
        }

        template <class FlowSolution>
        void computeResidual(const std::vector<double>& saturation, const typename GridInterface::Vector& gravity,
                             const FlowSolution& flow_sol, const Opm::SparseVector<double>& injection_rates,
                             const bool method_viscous, const bool method_gravity, const bool method_capillary,
                             std::vector<double>& sat_delta) const
        {
            if (!model_.ready()) OPM_THROW(std::runtime_error, "EulerUpstreamResidual (B200): initObj() has not been called");
            if (int(saturation.size()) != model_.numCells()) OPM_THROW(std::runtime_error, "saturation has the wrong size");
            sat_delta.clear();
            sat_delta.resize(saturation.size(), 0.0);
            model_.gatherFluxes(flow_sol);
            std::vector<int> src_cell;
            std::vector<double> src_rate;
            Model::sources(injection_rates, src_cell, src_rate);
            const double g[3] = { gravity[0], gravity[1], gravity[2] };
            model_.check(eu_compute_residual(model_.handle(), saturation.data(), g, model_.fluxes().data(), int(src_cell.size()),
                                             src_cell.data(), src_rate.data(), method_viscous, method_gravity, method_capillary,
                                             sat_delta.data()), "EulerUpstreamResidual");
        }

        void computeCapPressures(const std::vector<double>& saturation) const
        {
            if (!model_.ready()) OPM_THROW(std::runtime_error, "EulerUpstreamResidual (B200): initObj() has not been called");
            if (int(saturation.size()) != model_.numCells()) OPM_THROW(std::runtime_error, "saturation has the wrong size");
            cap_pressures_.resize(saturation.size());
            model_.check(eu_compute_cap_pressures(model_.handle(), saturation.data(), cap_pressures_.data()), "EulerUpstreamResidual");
        }
        /// What the last computeCapPressures stored (private in the reference, :118).
        const std::vector<double>& capPressures() const { return cap_pressures_; }

        const GridInterface& grid() const { return model_.grid(); }
        const ReservoirProperties& reservoirProperties() const { return model_.reservoirProperties(); }
        const BoundaryConditions& boundaryConditions() const { return model_.boundaryConditions(); }

        eu_handle deviceHandle() const { return model_.handle(); }

    private:
        typedef DeviceModel<GridInterface, ReservoirProperties, BoundaryConditions> Model;
        mutable Model model_;
        int device_;
        int mode_;
        mutable std::vector<double> cap_pressures_;
    };

} // namespace b200

#ifndef EULER_B200_KEEP_REFERENCE
    using b200::EulerUpstreamResidual;
#endif

} // namespace Opm

#endif // OPM_B200_EULERUPSTREAMRESIDUAL_HEADER
