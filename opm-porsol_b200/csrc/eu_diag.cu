// Diagnostics on device-resident data: the per-cell loops opm-porsol's drivers run right after the
// transport step (common/SimulatorUtilities.hpp), so that the saturation and the fluxes do not have to
// travel back to the host only to be reduced there.  Compiled with -fmad=false: every kernel follows the
// reference's operation order and is bit-identical to it.
//
//   k_cell_velocity     estimateCellVelocity      SimulatorUtilities.hpp:59-86
//   k_phase_velocities  computePhaseVelocities    :153-170
//   k_fractional_flow   rp.fractionalFlow loop    :273-279
//                       (ReservoirPropertyCapillary_impl.hpp:83-88,
//                        ReservoirPropertyCapillaryAnisotropicRelperm_impl.hpp:57-72)
//   (computeCapPressure :219-230 is k_strict_pc of eu_setup.cu)
#include "eu_internal.h"
#include "eu_strict_math.cuh"

namespace {

constexpr int kThreads = 256;
inline int div_up(long long a, int b) { return int((a + b - 1)/b); }

// One thread per own cell; the half-faces of a cell are contiguous (CSR), 56 bytes each (flux + centroid):
// bound by HBM, one pass over hf_flux and hf_centroid.
__global__ void __launch_bounds__(kThreads) k_cell_velocity(EuGridDev g, const double* __restrict__ hf_flux, double* __restrict__ out)
{
    const int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.own_hi) return;
    const double cc[3] = { g.cell_centroid[3LL*c], g.cell_centroid[3LL*c + 1], g.cell_centroid[3LL*c + 2] };
    const double vol = g.cell_volume[c];
    double cv[3] = { 0.0, 0.0, 0.0 };
    for (int h = g.hf_offset[c]; h < g.hf_offset[c + 1]; ++h) {
        const double s = hf_flux[h]/vol;                 // v *= flux/c->volume()
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double v = g.hf_centroid[3LL*h + i];
            v -= cc[i];
            v *= s;
            cv[i] += v;
        }
    }
    const long long o = 3LL*(c - g.own_lo);
    out[o] = cv[0]; out[o + 1] = cv[1]; out[o + 2] = cv[2];
}

template <int KIND>
__device__ __forceinline__ double frac_flow(const EuGridDev& g, const EuTablesDev& t, int c, double s)
{
    const int rock = g.rock ? g.rock[c] : 0;
    double m1[9], m2[9];
    sm_mobility<KIND>(t, 0, rock, s, m1);
    sm_mobility<KIND>(t, 1, rock, s, m2);
    if (KIND == 0) return m1[0]/(m1[0] + m2[0]);
    double ff = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double l1 = m1[4*d], l2 = m2[4*d];
        ff += l1/(l1 + l2);
    }
    ff /= 3.0;
    return ff;
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) k_fractional_flow(EuGridDev g, EuTablesDev t, const double* __restrict__ S,
                                                               double* __restrict__ out)
{
    const int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.own_hi) return;
    out[c - g.own_lo] = frac_flow<KIND>(g, t, c, S[c]);
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) k_phase_velocities(EuGridDev g, EuTablesDev t, const double* __restrict__ S,
                                                                const double* __restrict__ cell_v, double* __restrict__ vw,
                                                                double* __restrict__ vo)
{
    const int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.own_hi) return;
    const double f = frac_flow<KIND>(g, t, c, S[c]);
    const double omf = 1.0 - f;
    const long long o = 3LL*(c - g.own_lo);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double v = cell_v[o + i];
        vw[o + i] = v*f;
        vo[o + i] = v*omf;
    }
}

} // namespace

void eu_launch_cell_velocity(const EuGridDev& g, const double* hf_flux, double* out, cudaStream_t st)
{
    const int n = g.own_hi - g.own_lo;
    if (n > 0) k_cell_velocity<<<div_up(n, kThreads), kThreads, 0, st>>>(g, hf_flux, out);
}
void eu_launch_fractional_flow(const EuGridDev& g, const EuTablesDev& t, const double* S, double* out, cudaStream_t st)
{
    const int n = g.own_hi - g.own_lo;
    if (n <= 0) return;
    if (t.kind == EU_MOB_SCALAR) k_fractional_flow<0><<<div_up(n, kThreads), kThreads, 0, st>>>(g, t, S, out);
    else                         k_fractional_flow<1><<<div_up(n, kThreads), kThreads, 0, st>>>(g, t, S, out);
}
void eu_launch_phase_velocities(const EuGridDev& g, const EuTablesDev& t, const double* S, const double* cell_v,
                                double* vw, double* vo, cudaStream_t st)
{
    const int n = g.own_hi - g.own_lo;
    if (n <= 0) return;
    if (t.kind == EU_MOB_SCALAR) k_phase_velocities<0><<<div_up(n, kThreads), kThreads, 0, st>>>(g, t, S, cell_v, vw, vo);
    else                         k_phase_velocities<1><<<div_up(n, kThreads), kThreads, 0, st>>>(g, t, S, cell_v, vw, vo);
}
