// Host-only helpers of the C ABI: no device needed.
//
// eu_compute_cfl_factors restates ReservoirPropertyCapillary<3>::computeCflFactors
// (common/ReservoirPropertyCapillary_impl.hpp:190-281, RockJfunc.hpp:114-127) for callers that do
// not hold the reference's property class (bench.py, the Python binding).  The drop-in C++ header
// does not use it: it asks the caller's ReservoirProperties object for cflFactor*() directly.
#include "../../include/euler_b200.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

struct Curve {
    const double* x;
    const double* y;
    int n;
    int index(double v) const                 // opm-core tableIndex(): binary search
    {
        int hi = n - 1;
        if (hi < 2) return 0;
        int lo = 0;
        const bool ascend = x[hi] > x[0];
        while (hi - lo > 1) {
            const int mid = (hi + lo)/2;
            if ((v >= x[mid]) == ascend) lo = mid; else hi = mid;
        }
        return lo;
    }
    double slope(double v) const { const int i = index(v); return (y[i + 1] - y[i])/(x[i + 1] - x[i]); }
    double operator()(double v) const { const int i = index(v); return (y[i + 1] - y[i])/(x[i + 1] - x[i])*(v - x[i]) + y[i]; }
};

struct RockCfl { double inv_dff, inv_dfg, cap; };

RockCfl sample_rock(const eu_fluid& f, int rock, double min_perm, double max_poro)
{
    const int samples = 257;
    const double delta = 1.0/double(samples - 1);
    auto frac_flows = [&](double s, double& ff, double& fg) {
        double l1, l2;
        if (rock < 0) {
            l1 = (s*s)/f.viscosity[0];
            l2 = ((1 - s)*(1 - s))/f.viscosity[1];
        } else {
            const int b = f.table_offset[rock], n = f.table_offset[rock + 1] - b;
            const Curve krw = { f.table_s + b, f.table_cols[0] + b, n };
            const Curve kro = { f.table_s + b, f.table_cols[1] + b, n };
            l1 = krw(s)/f.viscosity[0];
            l2 = kro(s)/f.viscosity[1];
        }
        ff = l1/(l1 + l2);
        fg = l1*l2/(l1 + l2);
    };
    auto dpc = [&](double s) -> double {
        if (rock < 0) return 0.0;
        const int b = f.table_offset[rock], n = f.table_offset[rock + 1] - b;
        const Curve J = { f.table_s + b, f.table_cols[2] + b, n };
        double d = J.slope(s);
        if (f.use_jfunction_scaling) {
            const double k = 1.0*min_perm;
            double tr = 0; tr += k; tr += k; tr += k;
            d = d*f.sigma_cos_theta/std::sqrt(tr/(3*max_poro));
        }
        return std::fabs(d);
    };
    double last_ff, last_fg;
    frac_flows(0.0, last_ff, last_fg);
    double max_dff = -1e100, max_dfg = -1e100, max_fg = last_fg, max_dpc = dpc(0.0);
    for (int i = 1; i < samples; ++i) {
        const double s = double(i)*delta;
        double ff, fg;
        frac_flows(s, ff, fg);
        max_dff = std::max(max_dff, std::fabs(ff - last_ff)/delta);
        max_dfg = std::max(max_dfg, std::fabs(fg - last_fg)/delta);
        max_fg = std::max(max_fg, fg);
        max_dpc = rock < 0 ? 0.0 : std::max(max_dpc, dpc(s));
        last_ff = ff;
        last_fg = fg;
    }
    RockCfl r = { 1.0/max_dff, 1.0/max_dfg, max_fg*max_dpc };
    return r;
}

} // namespace

extern "C" int eu_compute_cfl_factors(const eu_fluid* fluid, int n_cells, const double* porosity,
                                      const double* permeability, const int* rock_id, double out[3])
{
    if (!fluid || !out) return EU_ERR_ARG;
    if (fluid->mobility_kind != EU_MOB_SCALAR) return EU_ERR_UNSUPPORTED;
    if (fluid->n_rocks == 0) {
        const RockCfl r = sample_rock(*fluid, -1, 0.0, 0.0);
        out[0] = r.inv_dff; out[1] = r.inv_dfg; out[2] = r.cap;
        return EU_OK;
    }
    if (!porosity || !permeability) return EU_ERR_ARG;
    std::vector<double> min_perm(fluid->n_rocks, 1e100), max_poro(fluid->n_rocks, 0.0);
    for (int c = 0; c < n_cells; ++c) {
        const int r = rock_id ? rock_id[c] : 0;
        const double* K = permeability + 9*size_t(c);
        double tr = 0; tr += K[0]; tr += K[4]; tr += K[8];
        min_perm[r] = std::min(min_perm[r], tr/3.0);
        max_poro[r] = std::max(max_poro[r], porosity[c]);
    }
    out[0] = 1e100; out[1] = 1e100; out[2] = 0.0;
    for (int r = 0; r < fluid->n_rocks; ++r) {
        const RockCfl f = sample_rock(*fluid, r, min_perm[r], max_poro[r]);
        out[0] = std::min(out[0], f.inv_dff);
        out[1] = std::min(out[1], f.inv_dfg);
        out[2] = std::max(out[2], f.cap);
    }
    return EU_OK;
}
