/* TEST INFRASTRUCTURE -- not product code.  See euler_oracle.h for the rules.
 *
 * Plain-C restatement of the reference's explicit transport over flat arrays.
 * Compile with -ffp-contract=off.  Citations are relative to /root/reference.
 */
#include "euler_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---------------------------------------------------------------------------------
 * Tabulated functions.  opm-core 2016.04 NonuniformTableLinear::operator() ->
 * linearInterpolation(): tableIndex() binary search, slope form, extrapolates.
 * (third-party; restated identically in oracle/shims/opm/core/utility/NonuniformTableLinear.hpp;
 * call sites RockJfunc.hpp:70-88,107,110)
 * --------------------------------------------------------------------------------- */
static int eo_table_index(int size, const double* table, double x)
{
    int n = size - 1;
    if (n < 2) {
        return 0;
    }
    int jl = 0;
    int ju = n;
    int ascend = (table[n] > table[0]);
    while (ju - jl > 1) {
        int jm = (ju + jl)/2;
        if ((x >= table[jm]) == ascend) {
            jl = jm;
        } else {
            ju = jm;
        }
    }
    return jl;
}

double eo_table_eval(int n, const double* x, const double* y, double xv)
{
    int i1 = eo_table_index(n, x, xv);
    int i2 = i1 + 1;
    return (y[i2] - y[i1])/(x[i2] - x[i1])*(xv - x[i1]) + y[i1];
}

static double eo_table_deriv(int n, const double* x, const double* y, double xv)
{
    int i1 = eo_table_index(n, x, xv);
    int i2 = i1 + 1;
    return (y[i2] - y[i1])/(x[i2] - x[i1]);
}

static double rock_col(const eo_case* c, int rock, int col, double s)
{
    int b = c->tab_offset[rock];
    int n = c->tab_offset[rock + 1] - b;
    return eo_table_eval(n, c->tab_s + b, c->tab_cols[col] + b, s);
}

/* ---------------------------------------------------------------------------------
 * Mobility.  kind 0: ReservoirPropertyCapillary_impl.hpp:44-70,118-131,154-168 (kr/viscosity,
 * quadratic fallback without rocks).  kind 1: ReservoirPropertyCapillaryAnisotropicRelperm_impl.hpp:
 * 45-101 + RockAnisotropicRelperm.hpp:70-77 (zero-filled 3x3, diagonal from tables, every entry
 * divided by the viscosity).  Output: row-major 3x3; for kind 0 only mob9[0] is meaningful.
 * --------------------------------------------------------------------------------- */
void eo_mobility(const eo_case* c, int phase, int cell, double s, double* mob9)
{
    const double visc = c->visc[phase];
    if (c->mobility_kind == 0) {
        double kr;
        if (c->n_rocks > 0) {
            int r = c->rock_id ? c->rock_id[cell] : 0;
            kr = rock_col(c, r, phase, s);
        } else {
            kr = phase == 0 ? s*s : (1 - s)*(1 - s);
        }
        mob9[0] = kr/visc;
        return;
    }
    for (int i = 0; i < 9; ++i) mob9[i] = 0.0;
    if (c->n_rocks > 0) {
        int r = c->rock_id ? c->rock_id[cell] : 0;
        mob9[0] = rock_col(c, r, 1 + 3*phase + 0, s);
        mob9[4] = rock_col(c, r, 1 + 3*phase + 1, s);
        mob9[8] = rock_col(c, r, 1 + 3*phase + 2, s);
        for (int i = 0; i < 9; ++i) mob9[i] = mob9[i]/visc;
    } else {
        double kr = phase == 0 ? s*s : (1.0 - s)*(1.0 - s);
        double mob = kr/visc;
        mob9[0] = mob9[4] = mob9[8] = mob;
    }
}

/* ReservoirPropertyCapillary_impl.hpp:83-88; anisotropic: ..AnisotropicRelperm_impl.hpp:56-72 */
double eo_fractional_flow(const eo_case* c, int cell, double s)
{
    double m1[9], m2[9];
    eo_mobility(c, 0, cell, s, m1);
    eo_mobility(c, 1, cell, s, m2);
    if (c->mobility_kind == 0) {
        return m1[0]/(m1[0] + m2[0]);
    }
    double ff = 0.0;
    for (int d = 0; d < 3; ++d) {
        double l1 = m1[4*d], l2 = m2[4*d];
        ff += l1/(l1 + l2);
    }
    ff /= 3.0;
    return ff;
}

/* ReservoirPropertyCommon_impl.hpp:465-476; RockJfunc.hpp:99-112; RockAnisotropicRelperm.hpp:79-83 */
double eo_cap_pressure(const eo_case* c, int cell, double s)
{
    if (c->n_rocks > 0) {
        int r = c->rock_id ? c->rock_id[cell] : 0;
        if (c->mobility_kind == 1) {
            return rock_col(c, r, 0, s);
        }
        double J = rock_col(c, r, 2, s);
        if (c->use_j) {
            const double* K = c->perm + 9*cell;
            double tr = 0;                         /* Matrix.hpp:637-647 */
            tr += K[0]; tr += K[4]; tr += K[8];
            double sqrt_k_phi = sqrt(tr/(3*c->poro[cell]));
            return J*c->sigma_cos_theta/sqrt_k_phi;
        }
        return J;
    }
    return 1e5*(1 - s);
}

/* ---------------------------------------------------------------------------------
 * small dense helpers with the reference's operation order
 * --------------------------------------------------------------------------------- */
/* Matrix.hpp:667-682: res = 0; for col: for row: res[row] += A(row,col)*x[col] */
static void prod3(const double* A, const double* x, double* res)
{
    res[0] = res[1] = res[2] = 0.0;
    for (int col = 0; col < 3; ++col) {
        for (int row = 0; row < 3; ++row) {
            res[row] += A[3*row + col]*x[col];
        }
    }
}
/* dune-cornerpoint Volumes.hpp inner(): std::inner_product from 0 */
static double inner3(const double* a, const double* b)
{
    double r = 0.0;
    r = r + a[0]*b[0];
    r = r + a[1]*b[1];
    r = r + a[2]*b[2];
    return r;
}
static double two_norm3(const double* a)
{
    double r = 0.0;
    r += a[0]*a[0];
    r += a[1]*a[1];
    r += a[2]*a[2];
    return sqrt(r);
}
/* opm-core Average.hpp: r = a; r += b; r *= 0.5 */
static void aver9(const double* a, const double* b, double* r)
{
    for (int i = 0; i < 9; ++i) { double t = a[i]; t += b[i]; t *= 0.5; r[i] = t; }
}

/* LAPACK dgetrf+dgetri for a 3x3 (Matrix.hpp:780-802), same unblocked algorithm as
 * oracle/ref_lapack3.cpp.  Row-major storage is passed as is (inv(A)^T == inv(A^T)). */
static int invert3(double* A)
{
    const int n = 3, ld = 3;
    int ipiv[3];
    int info = 0;
    for (int j = 0; j < n; ++j) {
        int p = j;
        double best = fabs(A[j + j*ld]);
        for (int i = j + 1; i < n; ++i) {
            if (fabs(A[i + j*ld]) > best) { best = fabs(A[i + j*ld]); p = i; }
        }
        ipiv[j] = p;
        if (A[p + j*ld] != 0.0) {
            if (p != j) {
                for (int k = 0; k < n; ++k) { double t = A[j + k*ld]; A[j + k*ld] = A[p + k*ld]; A[p + k*ld] = t; }
            }
            if (j < n - 1) {
                const double r = 1.0/A[j + j*ld];
                for (int i = j + 1; i < n; ++i) A[i + j*ld] *= r;
            }
        } else if (info == 0) {
            info = j + 1;
        }
        if (j < n - 1) {
            for (int k = j + 1; k < n; ++k) {
                const double akj = A[j + k*ld];
                for (int i = j + 1; i < n; ++i) A[i + k*ld] -= A[i + j*ld]*akj;
            }
        }
    }
    if (info != 0) return info;
    for (int j = 0; j < n; ++j) {
        A[j + j*ld] = 1.0/A[j + j*ld];
        const double ajj = -A[j + j*ld];
        for (int k = 0; k < j; ++k) {
            const double t = A[k + j*ld];
            if (t != 0.0) {
                for (int i = 0; i < k; ++i) A[i + j*ld] += t*A[i + k*ld];
                A[k + j*ld] = t*A[k + k*ld];
            }
        }
        for (int i = 0; i < j; ++i) A[i + j*ld] *= ajj;
    }
    double work[3];
    for (int j = n - 2; j >= 0; --j) {
        for (int i = j + 1; i < n; ++i) { work[i] = A[i + j*ld]; A[i + j*ld] = 0.0; }
        for (int k = j + 1; k < n; ++k) {
            const double t = -work[k];
            for (int i = 0; i < n; ++i) A[i + j*ld] += t*A[i + k*ld];
        }
    }
    for (int j = n - 2; j >= 0; --j) {
        const int jp = ipiv[j];
        if (jp != j) {
            for (int i = 0; i < n; ++i) { double t = A[i + j*ld]; A[i + j*ld] = A[i + jp*ld]; A[i + jp*ld] = t; }
        }
    }
    return 0;
}

/* Mobility wrappers: ScalarMobility (ReservoirPropertyCapillary.hpp:47-74) and
 * TensorMobility<3> (ReservoirPropertyCapillaryAnisotropicRelperm.hpp:49-100). */
static void mob_multiply(int kind, const double* m, const double* v, double* out)
{
    if (kind == 0) {
        out[0] = v[0]*m[0]; out[1] = v[1]*m[0]; out[2] = v[2]*m[0];
    } else {
        double t[3];
        prod3(m, v, t);
        out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
    }
}
static void mob_sum(int kind, const double* a, const double* b, double* r)
{
    int n = kind == 0 ? 1 : 9;
    for (int i = 0; i < n; ++i) r[i] = a[i] + b[i];
}
static void mob_average(int kind, const double* a, const double* b, double* r)
{
    int n = kind == 0 ? 1 : 9;
    for (int i = 0; i < n; ++i) r[i] = 0.5*(a[i] + b[i]);
}
static void mob_inverse(int kind, const double* a, double* r)
{
    if (kind == 0) {
        r[0] = 1.0/a[0];
    } else {
        for (int i = 0; i < 9; ++i) r[i] = a[i];
        invert3(r);
    }
}

/* EulerUpstreamResidual_impl.hpp:510-547 */
static void cap_gradient(const eo_case* c, int hf, int nbhf, int periodic_or_interior,
                         int cell, int nbcell, const double* cap_pressures, double* res)
{
    if (!periodic_or_interior) {
        res[0] = res[1] = res[2] = 0.0;
        return;
    }
    const double* cell_c = c->cell_centroid + 3*cell;
    const double* nb_c = c->cell_centroid + 3*nbcell;
    const double* f_c = c->hf_centroid + 3*hf;
    const double* nbf_c = c->hf_centroid + 3*nbhf;
    double a[3], b[3];
    for (int i = 0; i < 3; ++i) { a[i] = cell_c[i] - f_c[i]; b[i] = nb_c[i] - nbf_c[i]; }
    double d0 = two_norm3(a);
    double d1 = two_norm3(b);
    double cp0 = cap_pressures[cell];
    double cp1 = cap_pressures[nbcell];
    double val = (cp1 - cp0)/(d0 + d1);
    for (int i = 0; i < 3; ++i) {
        /* nb_c - nbf_c + f_c - cell_c, left to right */
        double t = nb_c[i] - nbf_c[i];
        t = t + f_c[i];
        t = t - cell_c[i];
        res[i] = t;
    }
    double nrm = two_norm3(res);
    for (int i = 0; i < 3; ++i) res[i] /= nrm;
    for (int i = 0; i < 3; ++i) res[i] *= val;
}

/* EulerUpstreamResidual_impl.hpp:459-467 (cap pressures), :472-505 (driver),
 * :100-300 (UpdateForCell), :403-421 (bid_to_face_) */
void eo_compute_residual(const eo_case* c, const double* sat, const double* gravity, const double* hf_flux,
                         int n_src, const int* src_cell, const double* src_rate,
                         double* cap_pressures, double* residual)
{
    const int N = c->N;
    const int kind = c->mobility_kind;
    int* bid_to_hf = (int*)malloc(sizeof(int)*(size_t)(c->n_bid > 0 ? c->n_bid : 1));
    for (int b = 0; b < c->n_bid; ++b) bid_to_hf[b] = -1;
    for (int cell = 0; cell < N; ++cell) {
        for (int hf = c->hf_offset[cell]; hf < c->hf_offset[cell + 1]; ++hf) {
            if (c->hf_nbr[hf] < 0 && c->bid_kind[c->hf_bid[hf]] == 1) bid_to_hf[c->hf_bid[hf]] = hf;
        }
    }
    /* half-face -> owning cell, needed for periodic partners */
    int* hf_cell = (int*)malloc(sizeof(int)*(size_t)(c->hf_offset[N] > 0 ? c->hf_offset[N] : 1));
    for (int cell = 0; cell < N; ++cell)
        for (int hf = c->hf_offset[cell]; hf < c->hf_offset[cell + 1]; ++hf) hf_cell[hf] = cell;

    if (c->method_capillary) {
        for (int cell = 0; cell < N; ++cell) cap_pressures[cell] = eo_cap_pressure(c, cell, sat[cell]);
    }
    for (int cell = 0; cell < N; ++cell) residual[cell] = 0.0;

    const double delta_rho = c->dens[0] - c->dens[1];
    int src_pos = 0;
    for (int c0 = 0; c0 < N; ++c0) {
        int cell[2];
        double cell_sat[2];
        cell[0] = c0;
        cell_sat[0] = sat[c0];
        for (int hf = c->hf_offset[c0]; hf < c->hf_offset[c0 + 1]; ++hf) {
            int nbhf = hf;
            int interior_like = 1;
            double dS = 0.0;
            if (c->hf_nbr[hf] < 0) {
                int bid = c->hf_bid[hf];
                if (c->bid_kind[bid] == 1) {
                    nbhf = bid_to_hf[c->bid_partner[bid]];
                    cell[1] = hf_cell[nbhf];
                    if (cell[0] > cell[1]) continue;
                    cell_sat[1] = sat[cell[1]];
                } else {
                    cell[1] = cell[0];
                    cell_sat[1] = c->bid_sat[bid];
                    interior_like = 0;
                }
            } else {
                cell[1] = c->hf_nbr[hf];
                if (cell[0] > cell[1]) continue;
                cell_sat[1] = sat[cell[1]];
            }
            const double loc_area = c->hf_area[hf];
            const double loc_flux = hf_flux[hf];
            const double* loc_normal = c->hf_normal + 3*hf;

            double aver_perm[9];
            aver9(c->perm + 9*cell[0], c->perm + 9*cell[1], aver_perm);
            double grav_influence[3];
            prod3(aver_perm, gravity, grav_influence);
            for (int i = 0; i < 3; ++i) grav_influence[i] *= delta_rho;
            const double G = c->method_gravity ? loc_area*inner3(loc_normal, grav_influence) : 0.0;
            const int triv_phase = G >= 0.0 ? 0 : 1;
            const int ups_cell = loc_flux >= 0.0 ? 0 : 1;
            double m_ups[2][9];
            eo_mobility(c, triv_phase, cell[ups_cell], cell_sat[ups_cell], m_ups[triv_phase]);
            const double sign_G[2] = { -1.0, 1.0 };
            double tmp[3], tmp2[3], tmp3[3];
            mob_multiply(kind, m_ups[triv_phase], grav_influence, tmp);
            double grav_flux_nontriv = sign_G[triv_phase]*loc_area*inner3(loc_normal, tmp);
            const int ups_cell_nontriv = (loc_flux + grav_flux_nontriv >= 0.0) ? 0 : 1;
            const int nontriv_phase = (triv_phase + 1) % 2;
            eo_mobility(c, nontriv_phase, cell[ups_cell_nontriv], cell_sat[ups_cell_nontriv], m_ups[nontriv_phase]);
            double m_tot[9], m_totinv[9];
            mob_sum(kind, m_ups[0], m_ups[1], m_tot);
            mob_inverse(kind, m_tot, m_totinv);

            double aver_sat = cell_sat[0];
            aver_sat += cell_sat[1];
            aver_sat *= 0.5;
            double m1c0[9], m1c1[9], m2c0[9], m2c1[9];
            eo_mobility(c, 0, cell[0], aver_sat, m1c0);
            eo_mobility(c, 0, cell[1], aver_sat, m1c1);
            eo_mobility(c, 1, cell[0], aver_sat, m2c0);
            eo_mobility(c, 1, cell[1], aver_sat, m2c1);
            double m_aver[2][9], m_aver_tot[9], m_aver_totinv[9];
            mob_average(kind, m1c0, m1c1, m_aver[0]);
            mob_average(kind, m2c0, m2c1, m_aver[1]);
            mob_sum(kind, m_aver[0], m_aver[1], m_aver_tot);
            mob_inverse(kind, m_aver_tot, m_aver_totinv);

            if (c->method_viscous) {
                double v[3] = { loc_normal[0], loc_normal[1], loc_normal[2] };
                for (int i = 0; i < 3; ++i) v[i] *= loc_flux;
                mob_multiply(kind, m_totinv, v, tmp);
                mob_multiply(kind, m_ups[0], tmp, tmp2);
                const double visc_change = inner3(loc_normal, tmp2);
                dS += visc_change;
            }
            if (c->method_gravity) {
                if (cell[0] != cell[1]) {
                    mob_multiply(kind, m_ups[1], grav_influence, tmp);
                    mob_multiply(kind, m_totinv, tmp, tmp2);
                    mob_multiply(kind, m_ups[0], tmp2, tmp3);
                    const double grav_change = loc_area*inner3(loc_normal, tmp3);
                    dS += grav_change;
                }
            }
            if (c->method_capillary) {
                double grad[3], cap_influence[3];
                cap_gradient(c, hf, nbhf, interior_like, cell[0], cell[1], cap_pressures, grad);
                prod3(aver_perm, grad, cap_influence);
                mob_multiply(kind, m_aver[1], cap_influence, tmp);
                mob_multiply(kind, m_aver_totinv, tmp, tmp2);
                mob_multiply(kind, m_aver[0], tmp2, tmp3);
                const double cap_change = loc_area*inner3(loc_normal, tmp3);
                dS += cap_change;
            }
            if (cell[0] != cell[1]) {
                residual[cell[0]] -= dS;
                residual[cell[1]] += dS;
            } else {
                residual[cell[0]] -= dS;
            }
        }
        /* source term: SparseVector::element (sorted cells) */
        double rate = 0.0;
        while (src_pos < n_src && src_cell[src_pos] < c0) ++src_pos;
        if (src_pos < n_src && src_cell[src_pos] == c0) rate = src_rate[src_pos];
        if (rate < 0.0) {
            rate *= eo_fractional_flow(c, c0, cell_sat[0]);
        }
        residual[c0] += rate;
    }
    free(bid_to_hf);
    free(hf_cell);
}

/* EulerUpstream_impl.hpp:355-385 (smallTimeStep), :336-349 (checkAndPossiblyClampSat),
 * porevol_ from :119-127 */
int eo_small_step(const eo_case* c, double* sat, double dt, const double* gravity, const double* hf_flux,
                  int n_src, const int* src_cell, const double* src_rate,
                  double* cap_pressures, double* residual, int* bad_cell, double* bad_value)
{
    eo_compute_residual(c, sat, gravity, hf_flux, n_src, src_cell, src_rate, cap_pressures, residual);
    for (int i = 0; i < c->N; ++i) {
        const double porevol = c->cell_volume[i]*c->poro[i];
        const double sat_change = dt*residual[i]/porevol;
        sat[i] += sat_change;
    }
    if (c->check_sat || c->clamp_sat) {
        for (int cell = 0; cell < c->N; ++cell) {
            if (sat[cell] > 1.0 || sat[cell] < 0.0) {
                if (c->clamp_sat) {
                    double v = sat[cell] < 1.0 ? sat[cell] : 1.0;   /* std::min(s, 1.0) */
                    sat[cell] = v > 0.0 ? v : 0.0;                    /* std::max(., 0.0) */
                } else if (sat[cell] > 1.001 || sat[cell] < -0.001) {
                    if (bad_cell) *bad_cell = cell;
                    if (bad_value) *bad_value = sat[cell];
                    return 1;
                }
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------------------------
 * Diagnostics the drivers run right after transport (common/SimulatorUtilities.hpp).
 * --------------------------------------------------------------------------------- */
/* estimateCellVelocity, SimulatorUtilities.hpp:59-86: per face v = centroid(f); v -= centroid(c);
 * v *= flux/volume; cell_v += v.  out: 3 doubles per cell. */
void eo_cell_velocity(const eo_case* c, const double* hf_flux, double* out)
{
    for (int cell = 0; cell < c->N; ++cell) {
        double cv[3] = { 0.0, 0.0, 0.0 };
        for (int hf = c->hf_offset[cell]; hf < c->hf_offset[cell + 1]; ++hf) {
            const double s = hf_flux[hf]/c->cell_volume[cell];
            for (int d = 0; d < 3; ++d) {
                double v = c->hf_centroid[3*(size_t)hf + d];
                v -= c->cell_centroid[3*(size_t)cell + d];
                v *= s;
                cv[d] += v;
            }
        }
        for (int d = 0; d < 3; ++d) out[3*(size_t)cell + d] = cv[d];
    }
}

/* computePhaseVelocities, SimulatorUtilities.hpp:153-170: v_w = v*f, v_o = v*(1.0 - f), f = rp.fractionalFlow */
void eo_phase_velocities(const eo_case* c, const double* sat, const double* cell_velocity, double* vw, double* vo)
{
    for (int cell = 0; cell < c->N; ++cell) {
        const double f = eo_fractional_flow(c, cell, sat[cell]);
        const double omf = 1.0 - f;
        for (int d = 0; d < 3; ++d) {
            const double v = cell_velocity[3*(size_t)cell + d];
            vw[3*(size_t)cell + d] = v*f;
            vo[3*(size_t)cell + d] = v*omf;
        }
    }
}

/* computeCapPressure, SimulatorUtilities.hpp:219-230 == EulerUpstreamResidual::computeCapPressures (:459-467) */
void eo_cap_pressures(const eo_case* c, const double* sat, double* out)
{
    for (int cell = 0; cell < c->N; ++cell) out[cell] = eo_cap_pressure(c, cell, sat[cell]);
}

/* CflCalculator.hpp:54-83 */
int eo_cfl_velocity(const eo_case* c, const double* hf_flux, double* dt_out)
{
    double dt = 1e100;
    for (int cell = 0; cell < c->N; ++cell) {
        double flux_p = 0.0, flux_n = 0.0;
        for (int hf = c->hf_offset[cell]; hf < c->hf_offset[cell + 1]; ++hf) {
            const double loc_flux = hf_flux[hf];
            if (loc_flux > 0) flux_p += loc_flux; else flux_n -= loc_flux;
        }
        double flux = flux_n > flux_p ? flux_n : flux_p;   /* std::max(flux_n, flux_p) */
        double loc_dt = (c->cfl_factor[0]*c->cell_volume[cell]*c->poro[cell])/flux;
        if (loc_dt == 0.0) return 2;
        if (loc_dt < dt) dt = loc_dt;
    }
    *dt_out = dt;
    return 0;
}

/* CflCalculator.hpp:90-134 */
double eo_cfl_gravity(const eo_case* c, const double* gravity)
{
    const double delta_rho = c->dens[0] - c->dens[1];
    double dt = 1e100;
    for (int cell = 0; cell < c->N; ++cell) {
        double flux = 0.0;
        for (int hf = c->hf_offset[cell]; hf < c->hf_offset[cell + 1]; ++hf) {
            double aver[9];
            const double* K;
            if (c->hf_nbr[hf] >= 0) {
                aver9(c->perm + 9*cell, c->perm + 9*c->hf_nbr[hf], aver);
                K = aver;
            } else {
                K = c->perm + 9*cell;
            }
            const double* n = c->hf_normal + 3*hf;
            double loc_gravity_flux = 0.0;
            for (int k = 0; k < 3; ++k) {
                for (int q = 0; q < 3; ++q) {
                    loc_gravity_flux += n[q]*(K[3*q + k]*gravity[k]*delta_rho);
                }
            }
            loc_gravity_flux *= c->hf_area[hf];
            if (loc_gravity_flux > 0) flux += loc_gravity_flux;
        }
        double loc_dt = (c->cfl_factor[1]*c->cell_volume[cell]*c->poro[cell])/flux;
        if (loc_dt < dt) dt = loc_dt;
    }
    return dt;
}

/* MatrixInverse.hpp:85-123 */
static void inverse3x3(const double* m, double* mi)
{
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double t1 = (e - f*h/i);
    double t2 = (c*h/i - b);
    double t3 = (f*g/i - d);
    double t4 = (a - c*g/i);
    double x = t4*t1 - t2*t3;
    mi[0] = t1/x;
    mi[1] = t2/x;
    mi[2] = -(c*t1 + f*t2)/(i*x);
    mi[3] = t3/x;
    mi[4] = t4/x;
    mi[5] = -(c*t3 + f*t4)/(i*x);
    mi[6] = -(g*t1 + h*t3)/(i*x);
    mi[7] = -(g*t2 + h*t4)/(i*x);
    mi[8] = 1/i + 1/(i*i*x)*(c*(g*t1 + h*t3) + f*(g*t2 + h*t4));
}

/* CflCalculator.hpp:142-176 */
double eo_cfl_capillary(const eo_case* c)
{
    double dt = 1e100;
    for (int cell = 0; cell < c->N; ++cell) {
        for (int hf = c->hf_offset[cell]; hf < c->hf_offset[cell + 1]; ++hf) {
            double aver[9], inv[9];
            const double* K;
            if (c->hf_nbr[hf] >= 0) {
                aver9(c->perm + 9*cell, c->perm + 9*c->hf_nbr[hf], aver);
                K = aver;
            } else {
                K = c->perm + 9*cell;
            }
            inverse3x3(K, inv);
            double d[3], t[3];
            for (int i = 0; i < 3; ++i) d[i] = c->hf_centroid[3*hf + i] - c->cell_centroid[3*cell + i];
            prod3(inv, d, t);
            double spatial = 0.0;                   /* FieldVector operator*: left fold from 0 */
            spatial += d[0]*t[0];
            spatial += d[1]*t[1];
            spatial += d[2]*t[2];
            double loc_dt = spatial/c->cfl_factor[2];
            dt = loc_dt < dt ? loc_dt : dt;          /* std::min(dt, loc_dt) */
        }
    }
    return dt;
}

/* EulerUpstream_impl.hpp:151-218 (transportSolve), :263-331 (computeCflTime) */
void eo_transport_solve(const eo_case* c, double* sat, double time_, const double* gravity, const double* hf_flux,
                        int n_src, const int* src_cell, const double* src_rate, eo_result* out)
{
    const int N = c->N;
    memset(out, 0, sizeof(*out));
    out->bad_cell = -1;
    double cfl_dt_v = 1e99, cfl_dt_g = 1e99, cfl_dt_c = 1e99;
    if (c->method_viscous && c->use_cfl_viscous) {
        if (eo_cfl_velocity(c, hf_flux, &cfl_dt_v)) { out->status = 2; return; }
    }
    if (c->method_gravity && c->use_cfl_gravity) cfl_dt_g = eo_cfl_gravity(c, gravity);
    if (c->method_capillary && c->use_cfl_capillary) cfl_dt_c = eo_cfl_capillary(c);
    out->cfl_dt[0] = cfl_dt_v; out->cfl_dt[1] = cfl_dt_g; out->cfl_dt[2] = cfl_dt_c;
    double m = cfl_dt_g < cfl_dt_v ? cfl_dt_g : cfl_dt_v;
    double cfl_dt = cfl_dt_c < m ? cfl_dt_c : m;
    cfl_dt *= c->courant;

    int nsteps;
    if (cfl_dt > time_) {
        nsteps = c->min_steps;
    } else {
        double a = ceil(time_/cfl_dt);
        double steps = ((double)INT_MAX < a) ? (double)INT_MAX : a;   /* std::min<double>(a, INT_MAX) */
        nsteps = (steps != steps) ? INT_MIN : (int)steps;
        nsteps = nsteps > c->min_steps ? nsteps : c->min_steps;
        nsteps = nsteps < c->max_steps ? nsteps : c->max_steps;
    }
    double dt = time_/nsteps;

    double* initial = (double*)malloc(sizeof(double)*(size_t)N);
    double* residual = (double*)malloc(sizeof(double)*(size_t)N);
    double* cap = (double*)malloc(sizeof(double)*(size_t)N);
    memcpy(initial, sat, sizeof(double)*(size_t)N);
    int finished = 0, repeats = 0;
    const int max_repeats = 10;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while (!finished) {
        int failed = 0;
        out->attempts++;
        for (int q = 0; q < nsteps; ++q) {
            out->substeps_executed++;
            if (eo_small_step(c, sat, dt, gravity, hf_flux, n_src, src_cell, src_rate, cap, residual,
                              &out->bad_cell, &out->bad_value)) {
                failed = 1;
                break;
            }
        }
        if (!failed) {
            finished = 1;
        } else {
            ++repeats;
            if (repeats > max_repeats) {
                out->status = 1;
                break;
            }
            nsteps *= 2;
            dt = time_/nsteps;
            memcpy(sat, initial, sizeof(double)*(size_t)N);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    out->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
    out->nsteps = nsteps;
    free(initial); free(residual); free(cap);
}

/* ReservoirPropertyCapillary_impl.hpp:190-281 (cflFracFlows, computeSingleRockCflFactors,
 * computeCflFactors); RockJfunc.hpp:114-127 (capPressDeriv).  Scalar-mobility class only. */
static void cfl_frac_flows(const eo_case* c, int rock, double s, double* ff_first, double* ff_gravity)
{
    double l1, l2;
    if (rock == -1) {
        l1 = (s*s)/c->visc[0];
        l2 = ((1 - s)*(1 - s))/c->visc[1];
    } else {
        l1 = rock_col(c, rock, 0, s)/c->visc[0];
        l2 = rock_col(c, rock, 1, s)/c->visc[1];
    }
    *ff_first = l1/(l1 + l2);
    *ff_gravity = l1*l2/(l1 + l2);
}

static double cap_press_deriv(const eo_case* c, int rock, double min_perm, double max_poro, double s)
{
    int b = c->tab_offset[rock];
    int n = c->tab_offset[rock + 1] - b;
    double dJ = eo_table_deriv(n, c->tab_s + b, c->tab_cols[2] + b, s);
    if (c->use_j) {
        /* eye(3)*min_perm: diagonal entries 1.0*min_perm */
        double k = 1.0*min_perm;
        double tr = 0; tr += k; tr += k; tr += k;
        double sqrt_k_phi = sqrt(tr/(3*max_poro));
        return dJ*c->sigma_cos_theta/sqrt_k_phi;
    }
    return dJ;
}

static void single_rock_cfl(const eo_case* c, int rock, double min_perm, double max_poro, double* out3)
{
    const int Ns = 257;
    double delta = 1.0/(double)(Ns - 1);
    double last_ff1, last_ffg;
    double max_der1 = -1e100, max_derg = -1e100;
    cfl_frac_flows(c, rock, 0.0, &last_ff1, &last_ffg);
    double max_ffg = last_ffg;
    double max_derpc = rock == -1 ? 0.0 : fabs(cap_press_deriv(c, rock, min_perm, max_poro, 0.0));
    for (int i = 1; i < Ns; ++i) {
        double s = (double)i*delta;
        double ff1, ffg;
        cfl_frac_flows(c, rock, s, &ff1, &ffg);
        double e1 = fabs(ff1 - last_ff1)/delta;
        double eg = fabs(ffg - last_ffg)/delta;
        max_der1 = max_der1 < e1 ? e1 : max_der1;      /* std::max(a,b) = a<b ? b : a */
        max_derg = max_derg < eg ? eg : max_derg;
        max_ffg = max_ffg < ffg ? ffg : max_ffg;
        if (rock != -1) {
            double d = fabs(cap_press_deriv(c, rock, min_perm, max_poro, s));
            max_derpc = max_derpc < d ? d : max_derpc;
        } else {
            max_derpc = 0.0;
        }
        last_ff1 = ff1;
        last_ffg = ffg;
    }
    out3[0] = 1.0/max_der1;
    out3[1] = 1.0/max_derg;
    out3[2] = max_ffg*max_derpc;
}

void eo_compute_cfl_factors(const eo_case* c, double* out3)
{
    if (c->n_rocks == 0) {
        single_rock_cfl(c, -1, 0.0, 0.0, out3);
        return;
    }
    double* min_perm = (double*)malloc(sizeof(double)*(size_t)c->n_rocks);
    double* max_poro = (double*)malloc(sizeof(double)*(size_t)c->n_rocks);
    for (int r = 0; r < c->n_rocks; ++r) { min_perm[r] = 1e100; max_poro[r] = 0.0; }
    for (int cell = 0; cell < c->N; ++cell) {
        int r = c->rock_id ? c->rock_id[cell] : 0;
        const double* K = c->perm + 9*cell;
        double tr = 0; tr += K[0]; tr += K[4]; tr += K[8];
        double v = tr/3.0;
        min_perm[r] = v < min_perm[r] ? v : min_perm[r];
        max_poro[r] = max_poro[r] < c->poro[cell] ? c->poro[cell] : max_poro[r];
    }
    out3[0] = 1e100; out3[1] = 1e100; out3[2] = 0.0;
    for (int r = 0; r < c->n_rocks; ++r) {
        double fac[3];
        single_rock_cfl(c, r, min_perm[r], max_poro[r], fac);
        out3[0] = fac[0] < out3[0] ? fac[0] : out3[0];
        out3[1] = fac[1] < out3[1] ? fac[1] : out3[1];
        out3[2] = out3[2] < fac[2] ? fac[2] : out3[2];
    }
    free(min_perm); free(max_poro);
}
