// Device versions of the per-cell loops opm-porsol's drivers run right after the transport step
// (opm/porsol/common/SimulatorUtilities.hpp), with the same names, argument order and results (bit-identical:
// the kernels of csrc/eu_diag.cu follow the reference's operation order).  The reference's functions take the
// grid / property objects; these take the B200 solver that already holds them flattened on the device
// (Opm::b200::EulerUpstream or Opm::b200::EulerUpstreamResidual, through deviceHandle()), and use the half-face
// fluxes of its last transportSolve / computeResidual.
//
//   estimateCellVelocity(cell_velocity, solver)                       SimulatorUtilities.hpp:59-86
//   computePhaseVelocities(v_water, v_oil, solver, saturation, v)     :153-170
//   computeCapPressure(cap_pressure, solver, saturation)              :219-230
//   computeFractionalFlow(frac_flow, solver, saturation)              the loop of writeVtkOutput, :273-279
#ifndef OPM_B200_DIAGNOSTICS_HEADER
#define OPM_B200_DIAGNOSTICS_HEADER

#include <opm/common/ErrorMacros.hpp>

#include <euler_b200.h>

#include <stdexcept>
#include <vector>

namespace Opm {
namespace b200 {

    namespace detail {
        inline void diagCheck(eu_handle h, int rc)
        {
            if (rc != EU_OK) OPM_THROW(std::runtime_error, "B200 diagnostics: " << eu_last_error(h));
        }
    }

    template <class Vec, class Solver>
    void estimateCellVelocity(std::vector<Vec>& cell_velocity, const Solver& solver)
    {
        eu_handle h = solver.deviceHandle();
        const int n = eu_local_cells(h);
        std::vector<double> flat(3*size_t(n));
        detail::diagCheck(h, eu_cell_velocity(h, flat.data()));
        cell_velocity.clear();
        cell_velocity.resize(n);
        for (int c = 0; c < n; ++c) for (int d = 0; d < 3; ++d) cell_velocity[c][d] = flat[3*size_t(c) + d];
    }

    template <class Vec, class Solver>
    void computePhaseVelocities(std::vector<Vec>& phase_velocity_water, std::vector<Vec>& phase_velocity_oil,
                                const Solver& solver, const std::vector<double>& saturation,
                                const std::vector<Vec>& cell_velocity)
    {
        eu_handle h = solver.deviceHandle();
        const size_t n = saturation.size();
        if (cell_velocity.size() != n || int(n) != eu_local_cells(h)) OPM_THROW(std::runtime_error, "size mismatch");
        std::vector<double> v(3*n), vw(3*n), vo(3*n);
        for (size_t c = 0; c < n; ++c) for (int d = 0; d < 3; ++d) v[3*c + d] = cell_velocity[c][d];
        detail::diagCheck(h, eu_phase_velocities(h, saturation.data(), v.data(), vw.data(), vo.data()));
        phase_velocity_water = cell_velocity;
        phase_velocity_oil = cell_velocity;
        for (size_t c = 0; c < n; ++c) {
            for (int d = 0; d < 3; ++d) { phase_velocity_water[c][d] = vw[3*c + d]; phase_velocity_oil[c][d] = vo[3*c + d]; }
        }
    }

    template <class Solver>
    void computeCapPressure(std::vector<double>& cap_pressure, const Solver& solver, const std::vector<double>& sat)
    {
        eu_handle h = solver.deviceHandle();
        cap_pressure.resize(sat.size());
        detail::diagCheck(h, eu_compute_cap_pressures(h, sat.data(), cap_pressure.data()));
    }

    template <class Solver>
    void computeFractionalFlow(std::vector<double>& frac_flow, const Solver& solver, const std::vector<double>& sat)
    {
        eu_handle h = solver.deviceHandle();
        frac_flow.resize(sat.size());
        detail::diagCheck(h, eu_fractional_flow(h, sat.data(), frac_flow.data()));
    }

} // namespace b200
} // namespace Opm

#endif // OPM_B200_DIAGNOSTICS_HEADER
