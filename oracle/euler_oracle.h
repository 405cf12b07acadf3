/* TEST INFRASTRUCTURE -- not product code.
 *
 * Plain-C CPU restatement of opm-porsol's explicit saturation transport
 * (opm/porsol/euler: EulerUpstream, EulerUpstreamResidual, CflCalculator) over the flat
 * half-face arrays of include/euler_b200.h.  Every function cites the reference
 * file:line it follows.  Serial, IEEE double, no FMA contraction (-ffp-contract=off):
 * operation order is the reference's, so results are bit-identical to the reference
 * headers compiled in oracle/_ref (tests/test_oracle_vs_reference.py pins that).
 *
 * Parity pinning: the reference ships no golden vectors for this path; this
 * restatement is pinned against outputs of the reference itself (oracle/_ref, built from
 * /root/reference in this container) and against tests/golden/ fixtures generated from it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or load this.  The product never does.
 */
#ifndef EULER_ORACLE_H
#define EULER_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eo_case {
    /* grid: CSR of half-faces in the reference's cell/face iteration order */
    int N;
    const int* hf_offset;      /* N+1 */
    const int* hf_nbr;         /* H; -1 on the boundary */
    const int* hf_bid;         /* H; boundary id, 0 interior */
    const double* hf_area;     /* H */
    const double* hf_normal;   /* 3H */
    const double* hf_centroid; /* 3H */
    const double* cell_volume; /* N */
    const double* cell_centroid; /* 3N */
    /* properties */
    const double* poro;        /* N */
    const double* perm;        /* 9N row-major */
    const int* rock_id;        /* N or NULL */
    int n_rocks;
    const int* tab_offset;     /* n_rocks+1 */
    const double* tab_s;
    /* mobility_kind 0: cols = {krw, kro, J}; kind 1: cols = {pc, kxw, kyw, kzw, kxo, kyo, kzo} */
    const double* tab_cols[7];
    int mobility_kind;
    int use_j;
    double sigma_cos_theta;
    double visc[2];
    double dens[2];
    double cfl_factor[3];      /* cflFactor, cflFactorGravity, cflFactorCapillary */
    /* boundary conditions by boundary id */
    int n_bid;
    const int* bid_kind;       /* 0 Dirichlet, 1 periodic */
    const double* bid_sat;
    const int* bid_partner;
    /* solver parameters (EulerUpstream_impl.hpp:59-73) */
    double courant;
    int method_viscous, method_gravity, method_capillary;
    int use_cfl_viscous, use_cfl_gravity, use_cfl_capillary;
    int min_steps, max_steps;
    int check_sat, clamp_sat;
} eo_case;

typedef struct eo_result {
    int status;          /* 0 ok, 1 saturation out of range after all retries, 2 cfl dt == 0 */
    int nsteps;          /* substeps of the last attempt */
    int attempts;        /* 1 + number of retries */
    long long substeps_executed; /* over all attempts, including the failed ones */
    int bad_cell;
    double bad_value;
    double cfl_dt[3];
    double seconds;      /* wall time of the substep loop */
} eo_result;

double eo_table_eval(int n, const double* x, const double* y, double xv);
void   eo_mobility(const eo_case* c, int phase, int cell, double s, double* mob9);
double eo_cap_pressure(const eo_case* c, int cell, double s);
double eo_fractional_flow(const eo_case* c, int cell, double s);

/* scratch: bid_to_hf (n_bid ints) is rebuilt inside; cap_pressures (N doubles) */
void eo_compute_residual(const eo_case* c, const double* sat, const double* gravity, const double* hf_flux,
                         int n_src, const int* src_cell, const double* src_rate,
                         double* cap_pressures, double* residual);
/* one substep; returns 0, or 1 with bad_cell/bad_value when the range check throws */
int eo_small_step(const eo_case* c, double* sat, double dt, const double* gravity, const double* hf_flux,
                  int n_src, const int* src_cell, const double* src_rate,
                  double* cap_pressures, double* residual, int* bad_cell, double* bad_value);
int eo_cfl_velocity(const eo_case* c, const double* hf_flux, double* dt);   /* returns 2 when a cell gives dt == 0 */
double eo_cfl_gravity(const eo_case* c, const double* gravity);
double eo_cfl_capillary(const eo_case* c);
void eo_transport_solve(const eo_case* c, double* sat, double time, const double* gravity, const double* hf_flux,
                        int n_src, const int* src_cell, const double* src_rate, eo_result* out);
/* diagnostics (common/SimulatorUtilities.hpp:59-86, :153-170, :219-230); vectors are 3 doubles per cell */
void eo_cell_velocity(const eo_case* c, const double* hf_flux, double* out);
void eo_phase_velocities(const eo_case* c, const double* sat, const double* cell_velocity, double* vw, double* vo);
void eo_cap_pressures(const eo_case* c, const double* sat, double* out);
/* ReservoirPropertyCapillary<3>::computeCflFactors restated (kind 0 only) */
void eo_compute_cfl_factors(const eo_case* c, double* out3);

#ifdef __cplusplus
}
#endif
#endif
