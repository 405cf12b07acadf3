// Drop-in replacement for opm-porsol's <opm/porsol/euler/EulerUpstream.hpp>.
//
// Same class template, same members and argument meaning as the reference
// (opm/porsol/euler/EulerUpstream.hpp:51-146, EulerUpstream_impl.hpp:59-218), but the work runs on a
// B200 through the C ABI of include/euler_b200.h:
//
//   init(param)                     -> eu_set_params            (same parameter keys and defaults)
//   initObj(grid, resprop, bc)      -> walks the GridInterface once, in the reference's cell / face
//                                      iteration order, flattens grid + properties + boundary conditions
//                                      and uploads them (b200/DeviceModel.hpp: eu_grid_begin/append/end,
//                                      eu_set_fluid)
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

#include <opm/porsol/euler/b200/DeviceModel.hpp>

#include <euler_b200.h>

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
#include <vector>

namespace Opm {
namespace b200 {

    template <class GridInterface, class ReservoirProperties, class BoundaryConditions>
    class EulerUpstream
    {
    public:
        EulerUpstream() : device_(0), mode_(EU_MODE_AUTO) { eu_default_params(&par_); }
        EulerUpstream(const GridInterface& grid, const ReservoirProperties& resprop, const BoundaryConditions& boundary)
            : device_(0), mode_(EU_MODE_AUTO)
        {
            eu_default_params(&par_);
            initObj(grid, resprop, boundary);
        }
        EulerUpstream(const EulerUpstream&) = delete;
        EulerUpstream& operator=(const EulerUpstream&) = delete;

        /// Same keys and defaults as the reference (EulerUpstream_impl.hpp:95-108), plus three optional ones:
        /// b200_device (CUDA ordinal, default 0), b200_devices (comma-separated ordinals: the grid is split into slabs over
        /// these devices of the process) and b200_mode (0 auto, 1 strict = bit-identical, 2 fast).
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
            // b200_devices=0,1,..: several devices of this process, the grid split into slabs of contiguous cell indices
            const std::string list = param.getDefault("b200_devices", std::string());
            devices_.clear();
            for (std::string::size_type p = 0; p < list.size();) {
                const std::string::size_type q = list.find(',', p);
                const std::string tok = list.substr(p, q == std::string::npos ? std::string::npos : q - p);
                if (!tok.empty()) devices_.push_back(std::atoi(tok.c_str()));
                if (q == std::string::npos) break;
                p = q + 1;
            }
            if (model_.ready()) model_.setParams(par_);
        }

        void init(const Opm::parameter::ParameterGroup& param, const GridInterface& grid,
                  const ReservoirProperties& resprop, const BoundaryConditions& boundary)
        {
            init(param);
            initObj(grid, resprop, boundary);
        }

        void initObj(const GridInterface& grid, const ReservoirProperties& resprop, const BoundaryConditions& boundary)
        {
            if (devices_.size() > 1) model_.create(devices_, mode_, par_, grid, resprop, boundary, "EulerUpstream");
            else model_.create(devices_.empty() ? device_ : devices_[0], mode_, par_, grid, resprop, boundary, "EulerUpstream");
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
            if (model_.ready()) model_.setParams(par_);
        }

        /// Report of the last transportSolve (step count, retries, CFL times, device time).
        const eu_report& lastReport() const { return report_; }

        template <class PressureSolution>
        void transportSolve(std::vector<double>& saturation, const double time,
                            const typename GridInterface::Vector& gravity, const PressureSolution& pressure_sol,
                            const Opm::SparseVector<double>& injection_rates) const
        {
            if (!model_.ready()) OPM_THROW(std::runtime_error, "EulerUpstream (B200): initObj() has not been called");
            if (int(saturation.size()) != model_.numCells()) OPM_THROW(std::runtime_error, "saturation has the wrong size");
            model_.gatherFluxes(pressure_sol);
            std::vector<int> src_cell;
            std::vector<double> src_rate;
            Model::sources(injection_rates, src_cell, src_rate);
            const double g[3] = { gravity[0], gravity[1], gravity[2] };
            const int rc = model_.transportSolve(saturation, time, g, src_cell, src_rate, &report_);
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
            if (rc != EU_OK) OPM_THROW(std::runtime_error, "EulerUpstream (B200): " << model_.lastError());
        }

        /// The device solver behind this object (C ABI handle), e.g. for the diagnostics of euler_b200.h.
        eu_handle deviceHandle() const { return model_.handle(); }

    protected:
        typedef DeviceModel<GridInterface, ReservoirProperties, BoundaryConditions> Model;

        mutable Model model_;
        eu_params par_;
        int device_;
        std::vector<int> devices_;
        int mode_;
        mutable eu_report report_;
    };

} // namespace b200

#ifndef EULER_B200_KEEP_REFERENCE
    using b200::EulerUpstream;
#endif

} // namespace Opm

#endif // OPM_B200_EULERUPSTREAM_HEADER
