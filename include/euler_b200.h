/*
 * euler_b200.h -- C ABI of the B200-native explicit saturation transport.
 *
 * This is the boundary a host simulator binds to in place of opm-porsol's
 *     Opm::EulerUpstream<GridInterface, ReservoirProperties, BoundaryConditions>
 * (reference: opm/porsol/euler/EulerUpstream.hpp:51-146).  The reference has no FFI: its
 * "plugin API" is that class template, selected by SimulatorTraits.hpp:89-101 and called from
 * SimulatorBase.hpp:213 / examples/SimulatorTester.hpp:83-85.  The header-only C++ mirror of
 * that template (opm-porsol_b200/host/opm/porsol/euler/EulerUpstream.hpp) flattens whatever
 * grid / property / boundary-condition objects it is given and forwards to the entry points
 * below; INTEGRATION.md shows the two-line change a maintainer makes.
 *
 * Conventions
 *   - plain C types, caller-owned host buffers, no exceptions across the boundary;
 *   - every call returns EU_OK (0) or an error code; eu_last_error() gives the text;
 *   - all floating point data is IEEE double, all indices are 32-bit int;
 *   - cells are numbered in the reference's cell iteration order (== c->index(), as for
 *     CpGrid, GridInterfaceEuler.hpp:70-80); the half-faces of a cell are stored in the order
 *     `for (f = c->facebegin(); f != c->faceend(); ++f)`, i.e. "the reference's face order";
 *     a half-face index is the running count in that double loop and equals the index of
 *     FlowSolution::outflux(int hf) (IncompFlowSolverHybrid.hpp:430-433);
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device and the
 *     compiled sm_100a kernels, and fails with EU_ERR_CUDA otherwise.
 */
#ifndef EULER_B200_H
#define EULER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EU_ABI_VERSION 1

typedef struct eu_solver* eu_handle;

enum {
    EU_OK = 0,
    EU_ERR_ARG = 1,            /* bad argument / call order */
    EU_ERR_CUDA = 2,           /* CUDA runtime failure or no device */
    EU_ERR_SAT_RANGE = 3,      /* "Saturation out of range in EulerUpstream" after 10 retries
                                  (EulerUpstream_impl.hpp:203-212,344-346) */
    EU_ERR_CFL_ZERO = 4,       /* "Cfl computation gave dt = 0.0" (CflCalculator.hpp:75-77) */
    EU_ERR_UNSUPPORTED = 5,
    EU_ERR_COMM = 6
};

/* half-face kinds */
enum { EU_HF_INTERIOR = 0, EU_HF_DIRICHLET = 1, EU_HF_PERIODIC = 2 };

/* mobility kinds: ReservoirPropertyCapillary<3> (ScalarMobility) or
 * ReservoirPropertyCapillaryAnisotropicRelperm<3> (diagonal TensorMobility<3>) */
enum { EU_MOB_SCALAR = 0, EU_MOB_DIAGONAL = 1 };

/* arithmetic modes */
enum {
    EU_MODE_AUTO = 0,    /* fast wherever the FAST tables apply: scalar mobility, and the diagonal tensor mobility class on
                            any grid (axis-aligned normals: scalar formula per face axis; oblique normals: the
                            three-component kernel); strict only when the rock tables do not fit the FAST search
                            (nodes closer than 1/4096, more than 48 curve sets) */
    EU_MODE_STRICT = 1,  /* reference operation order, no FMA contraction: bit-identical to the reference */
    EU_MODE_FAST = 2     /* static per-face quantities pre-contracted to scalars, FMA allowed:
                            |dS| error ~1e-16 per substep (gate: 1e-12), identical step counts */
};

typedef struct eu_config {
    int abi_version;       /* EU_ABI_VERSION */
    int device;            /* CUDA device ordinal of this process */
    int mode;              /* EU_MODE_* */
    /* z-slab / index-range decomposition: this process owns global cells [own_begin, own_end).
       world_size == 1: own_begin = 0, own_end = number of cells. */
    int rank;
    int world_size;
    int own_begin;
    int own_end;
} eu_config;

/* The 11 scalars of EulerUpstream (EulerUpstream_impl.hpp:59-73; keys read at :95-108). */
typedef struct eu_params {
    double courant_number;       /* 0.5 */
    int method_viscous;          /* true */
    int method_gravity;          /* true */
    int method_capillary;        /* true */
    int use_cfl_viscous;         /* true */
    int use_cfl_gravity;         /* true */
    int use_cfl_capillary;       /* true */
    int minimum_small_steps;     /* 1 */
    int maximum_small_steps;     /* 10000 */
    int check_sat;               /* true */
    int clamp_sat;               /* false */
} eu_params;

/* One chunk of consecutive cells [first_cell, first_cell + n_cells) with their half-faces,
 * replacing the walk over GridInterfaceEuler (GridInterfaceEuler.hpp:126-216,335-363) and the
 * per-cell property accessors (ReservoirPropertyCommon.hpp:130,150; cell_to_rock_).
 * A rank of a decomposed run appends its own cells and every remote cell its faces touch
 * (ghost cells), in ascending global cell order. */
typedef struct eu_grid_chunk {
    int first_cell;
    int n_cells;
    const int* hf_count;          /* n_cells: number of half-faces of each cell */
    /* per half-face, n_hf = sum(hf_count) */
    const int* hf_neighbour;      /* neighbour cell (global), -1 on the boundary
                                     (Face::neighbourCellIndex(), GridInterfaceEuler.hpp:192-199) */
    const double* hf_area;        /* Face::area() */
    const double* hf_normal;      /* 3*n_hf, Face::normal() (unit outer normal) */
    const double* hf_centroid;    /* 3*n_hf, Face::centroid() */
    /* boundary half-faces of this chunk (satCond(face), BoundaryConditions.hpp:420-425) */
    int n_bnd;
    const int* bnd_hf;            /* chunk-relative half-face index, ascending */
    const int* bnd_kind;          /* EU_HF_DIRICHLET / EU_HF_PERIODIC */
    const double* bnd_sat;        /* SatBC::saturation() for Dirichlet faces */
    const int* bnd_partner_cell;  /* periodic: cell of bid_to_face_[getPeriodicPartner(bid)]
                                     (EulerUpstreamResidual_impl.hpp:118-120) */
    const int* bnd_partner_face;  /* periodic: localIndex() of that partner face */
    /* per cell */
    const double* cell_volume;    /* Cell::volume() */
    const double* cell_centroid;  /* 3*n_cells, Cell::centroid() */
    const double* porosity;       /* rp.porosity(c) */
    const double* permeability;   /* 9*n_cells row-major, rp.permeability(c) */
    const int* rock_id;           /* cell_to_rock_[c]; may be NULL when there are no rock tables */
} eu_grid_chunk;

/* Fluid and rock description (ReservoirPropertyCommon.hpp:232-248, RockJfunc.hpp:220-228,
 * RockAnisotropicRelperm.hpp:154-159). */
typedef struct eu_fluid {
    int mobility_kind;           /* EU_MOB_* */
    double viscosity[2];
    double density[2];
    double cfl_factor[3];        /* rp.cflFactor(), cflFactorGravity(), cflFactorCapillary() */
    int use_jfunction_scaling;   /* RockJfunc::use_jfunction_scaling_ (scalar mobility only) */
    double sigma_cos_theta;      /* RockJfunc::sigma_cos_theta_ */
    int n_rocks;                 /* 0: no rock tables -> quadratic relperm, pc = 1e5(1-S) */
    const int* table_offset;     /* n_rocks+1 offsets into the concatenated node arrays */
    const double* table_s;       /* saturation nodes */
    /* EU_MOB_SCALAR:   columns {krw, kro, J}
       EU_MOB_DIAGONAL: columns {pc, krxx_w, kryy_w, krzz_w, krxx_o, kryy_o, krzz_o} */
    const double* table_cols[7];
} eu_fluid;

/* Per-call report of eu_transport_solve. */
typedef struct eu_report {
    int status;                  /* EU_OK, EU_ERR_SAT_RANGE or EU_ERR_CFL_ZERO */
    int nsteps;                  /* nr_transport_steps of the last attempt */
    int attempts;                /* 1 + number of retries ("repeats") */
    long long substeps_executed; /* substeps launched over all attempts */
    int bad_cell;                /* first offending cell of the failing substep, or -1 */
    double bad_value;            /* its saturation */
    double cfl_dt[3];            /* viscous, gravity, capillary CFL times (1e99 when unused) */
    double dt;                   /* dt_transport of the last attempt */
    double device_ms;            /* CUDA-event time of the substep loop of the last attempt */
    int kernel_launches;         /* kernels launched by this call */
} eu_report;

/* ---- life cycle ------------------------------------------------------------------------ */
int eu_create(const eu_config* cfg, eu_handle* out);
/* number of CUDA devices this process sees (0 without a driver / device) */
int eu_device_count(void);
void eu_destroy(eu_handle h);
const char* eu_last_error(eu_handle h);     /* h may be NULL: error of the last failed eu_create */
int eu_abi_version(void);

/* EulerUpstream::init(param) (EulerUpstream_impl.hpp:95-108) */
int eu_set_params(eu_handle h, const eu_params* p);
void eu_default_params(eu_params* p);       /* EulerUpstream::EulerUpstream() (:59-73) */

/* ---- EulerUpstream::initObj(grid, resprop, boundary) (EulerUpstream_impl.hpp:119-127,
 *      EulerUpstreamResidual_impl.hpp:391-432) ------------------------------------------- */
/* n_local_cells / n_local_halffaces: exact totals of what eu_grid_append will deliver to this
 * rank (own + ghost); device storage is sized from them. */
int eu_grid_begin(eu_handle h, int n_cells_global, int n_local_cells, long long n_local_halffaces);
int eu_grid_append(eu_handle h, const eu_grid_chunk* chunk);
int eu_set_fluid(eu_handle h, const eu_fluid* fluid);
int eu_grid_end(eu_handle h);               /* builds the device structures */

/* number of local cells / half-faces held by this rank, in upload order (own + ghost) */
int eu_local_cells(eu_handle h);
long long eu_local_halffaces(eu_handle h);
/* the arithmetic mode in effect after eu_grid_end: EU_MODE_STRICT or EU_MODE_FAST (EU_MODE_AUTO resolves to FAST for
 * both mobility classes on any grid -- the tensor class on oblique normals runs the three-component FAST kernel, NOT
 * STRICT -- and to STRICT only when the rock tables do not fit the FAST interval search) */
int eu_resolved_mode(eu_handle h);
/* fraction of (slice, slot) pairs whose adjacency is described by an 8-byte descriptor instead of 32 records */
double eu_regular_fraction(eu_handle h);
/* the work plan of the FAST substep kernel, as built by the last substep / transportSolve (zeros before that and in
 * STRICT mode): out[0] = fraction of the own slices (32 cells) that belong to a slice class, out[1] = number of work
 * items, out[2] = longest march (slices per item), out[3] = mean march length over the class items.  When the box
 * kernel ran (local numbering c = x + nx*(y + ny*z): tiles swept along z with TMA-staged operands) out[0] = 1, out[1] =
 * work units (tile x z-chunk), out[2], out[3] = planes per unit; out[4] = 1 when the box kernel ran, else 0.  Tests use it
 * to assert which code path a parity case exercised; EU_BOX=0 in the environment keeps the slice-class kernel. */
int eu_work_plan(eu_handle h, double out[5]);

/* ---- EulerUpstream::transportSolve (EulerUpstream_impl.hpp:151-218) -------------------
 * saturation:  local cells (in/out; ghost entries are inputs only)
 * hf_flux:     pressure_sol.outflux(f) for every local half-face, upload order
 * sources:     injection_rates as (cell, rate) pairs with ascending global cell index
 * Host buffers; the copies to and from the device are part of the call. */
int eu_transport_solve(eu_handle h, double* saturation, double time, const double gravity[3],
                       const double* hf_flux, int n_src, const int* src_cell, const double* src_rate,
                       eu_report* report);

/* ---- flux hand-off from the host pressure solver (FlowSolution::outflux, IncompFlowSolverHybrid.hpp:426-438) -----
 * eu_transport_solve copies `saturation` and `hf_flux` over PCIe inside the call; from pageable memory that copy runs
 * at a fraction of the link rate.  Two remedies, both optional:
 *  - eu_host_alloc / eu_host_free: page-locked host memory for the caller's flat flux array (the drop-in C++ header
 *    gathers pressure_sol.outflux(f) into such a buffer);
 *  - opt-in, EU_PIN_CACHE=1 in the environment: buffers of >= 1 MiB passed to eu_transport_solve / eu_upload_state /
 *    eu_upload_saturation / eu_download_saturation / eu_compute_residual that are not page-locked yet are registered
 *    with the driver on first use and stay registered while the same (pointer, size) keeps coming (an IMPES loop
 *    passes the same vectors every step); at most 4 registrations per solver, released by eu_destroy or
 *    eu_host_unpin_all.  It is off by default because a registered buffer must not be freed or reallocated while the
 *    registration lives (the driver would keep transferring from the old pages): enable it only when the application
 *    keeps its vectors for the lifetime of the solver, or calls eu_host_unpin_all before releasing them. */
void* eu_host_alloc(unsigned long long bytes);
void eu_host_free(void* p);
void eu_host_unpin_all(eu_handle h);

/* ---- device-resident variant: state stays in HBM between calls ------------------------- */
int eu_upload_state(eu_handle h, const double* saturation, const double* hf_flux);
int eu_upload_saturation(eu_handle h, const double* saturation);
int eu_download_saturation(eu_handle h, double* saturation);
int eu_transport_solve_resident(eu_handle h, double time, const double gravity[3],
                                int n_src, const int* src_cell, const double* src_rate, eu_report* report);

/* ---- pieces exposed for per-substep parity tests --------------------------------------
 * eu_cfl_times: CflCalculator.hpp:54-176 on the resident state (always all three terms).
 * eu_small_step: EulerUpstream::smallTimeStep (:355-385) once, on the resident state;
 *                residual_out (local own cells) may be NULL. */
int eu_cfl_times(eu_handle h, const double gravity[3], double out[3]);
int eu_small_step(eu_handle h, double dt, const double gravity[3],
                  int n_src, const int* src_cell, const double* src_rate,
                  double* residual_out, int* bad_cell, double* bad_value);

/* ---- the residual as an operator: EulerUpstreamResidual::computeResidual(saturation, gravity, flow_sol,
 *      injection_rates, method_viscous, method_gravity, method_capillary, sat_delta)
 *      (EulerUpstreamResidual.hpp:83-91, _impl.hpp:472-505).  Second caller in the reference:
 *      ImplicitCapillarity::transportSolve (ImplicitCapillarity_impl.hpp:178-180, capillary = false).
 * saturation: all local cells (host); hf_flux: all local half-faces (host), or NULL to reuse the fluxes already
 * resident from eu_upload_state / the last eu_transport_solve; sat_delta: the own cells (host, out).
 * The method flags are arguments, as in the reference; the solver's eu_params are not consulted or changed,
 * nor is the resident saturation.  With method_capillary set the call includes computeCapPressures(saturation)
 * (:459-467), which the reference's callers issue separately beforehand (EulerUpstream_impl.hpp:362-369). */
int eu_compute_residual(eu_handle h, const double* saturation, const double gravity[3], const double* hf_flux,
                        int n_src, const int* src_cell, const double* src_rate,
                        int method_viscous, int method_gravity, int method_capillary, double* sat_delta);
/* EulerUpstreamResidual::computeCapPressures (:459-467): pc[c] = rp.capillaryPressure(c, saturation[c]) for all
 * local cells, reference operation order (bit-identical in every mode). */
int eu_compute_cap_pressures(eu_handle h, const double* saturation, double* cap_pressures);

/* ---- diagnostics on resident data, the per-cell / per-face loops the drivers run right after transport
 *      (SimulatorUtilities.hpp).  All use the resident half-face fluxes and operate on the own cells. ----
 * eu_cell_velocity: estimateCellVelocity (:59-86): v_c = (1/volume) sum_f flux_f (face centroid - cell centroid);
 *                   out 3 doubles per own cell.
 * eu_phase_velocities: computePhaseVelocities (:153-170): v_w = v f, v_o = v (1.0 - f) with f = rp.fractionalFlow(c, S[c]);
 *                   saturation NULL = the resident state (else all local cells, host), cell_velocity NULL = computed
 *                   from the resident fluxes (else 3 doubles per own cell, host); out 3 doubles per own cell and phase.
 * eu_fractional_flow: rp.fractionalFlow(c, S[c]) per own cell (writeVtkOutput loop, :273-279;
 *                   ReservoirPropertyCapillary_impl.hpp:83-88, ...AnisotropicRelperm_impl.hpp:57-72); saturation NULL =
 *                   the resident state.
 * All three follow the reference's operation order (bit-identical results in every arithmetic mode). */
int eu_cell_velocity(eu_handle h, double* cell_velocity);
int eu_phase_velocities(eu_handle h, const double* saturation, const double* cell_velocity,
                        double* water_velocity, double* oil_velocity);
int eu_fractional_flow(eu_handle h, const double* saturation, double* frac_flow);

/* ---- multi-GPU plumbing (one process per GPU; see DESIGN.md "Multi-GPU") ----------------
 * Each rank owns a contiguous range of global cells (z-slab) and holds the remote cells its faces touch
 * as ghost cells.  After every substep the new saturations (and, in FAST mode, capillary pressures) of
 * the cells a neighbour holds as ghosts are written straight into that neighbour's HBM with
 * peer-to-peer stores over NVLink (buffers exported with CUDA IPC), followed by an epoch flag; the
 * neighbour's next substep waits for the flag on its own stream.  Nothing goes through the host.
 * The host application only has to all-gather one opaque, variable-size blob per rank (eu_comm_export)
 * and hand all of them to eu_comm_connect, and to provide a tiny min/max reduction for the CFL times
 * and the range-check flag.  The library itself stays free of MPI/NCCL. */
int eu_comm_blob_size(eu_handle h);                          /* bytes of this rank's blob (after eu_grid_end) */
int eu_comm_export(eu_handle h, void* blob);                 /* fill this rank's blob */
int eu_comm_connect(eu_handle h, int n_blobs, const void* const* blobs, const int* blob_sizes);  /* rank order */
typedef void (*eu_allreduce_fn)(void* user, double* values, int n, int op /*0 = min, 1 = max*/);
int eu_comm_set_allreduce(eu_handle h, eu_allreduce_fn fn, void* user);
/* Pure host logic behind eu_comm_connect, exposed for tests: which of the cells in [own_begin, own_end)
 * does a peer hold as ghosts?  ghost_global ascending.  Writes, for every match, the global cell id and
 * the peer's local slot; returns the number of matches. */
int eu_comm_plan_sends(int own_begin, int own_end, int n_ghost, const int* ghost_global, const int* ghost_local,
                       int* send_global, int* send_peer_local);

/* ---- host-side helper, no device needed: ReservoirPropertyCapillary<3>::computeCflFactors
 *      (ReservoirPropertyCapillary_impl.hpp:190-281) for callers without the reference's
 *      property class (bench, Python).  perm/poro/rock_id cover n_cells cells. */
int eu_compute_cfl_factors(const eu_fluid* fluid, int n_cells, const double* porosity,
                           const double* permeability, const int* rock_id, double out[3]);

/* ---- host-side helper, no device needed: periodic partner matching of boundary faces, findPeriodicPartners
 *      (BoundaryPeriodicity.hpp:86-177) with match() (BoundaryPeriodicity.cpp:25-49) -- the step that produces the
 *      partner table behind eu_grid_chunk::bnd_partner_*.  Faces in the order the grid walk meets them.
 * in:  centroid[3n], area[n], is_periodic[6] (xmin, xmax, ymin, ...), spatial_tolerance (1e-6 in the reference)
 * out: canon_pos[n] (0 xmin, 1 xmax, 2 ymin, ...), partner[n] (index of the partner face, -1 if none), side_areas[6]
 * Returns EU_ERR_ARG when a centroid is not on the bounding box of all centroids (the reference throws).
 * O(n log n); identical to the reference wherever exactly one face qualifies as partner. */
int eu_match_periodic_faces(int n, const double* centroid, const double* area, const int is_periodic[6],
                            double spatial_tolerance, int* canon_pos, int* partner, double side_areas[6]);

/* ---- host-side helpers for the callers and data formats on either side of the path (no device needed) -----------
 * eu_build_face_indices: GridInterfaceEuler::buildFaceIndices (common/GridInterfaceEuler.hpp:498-611) over the flat CSR
 *      adjacency (hf_neighbour < 0 = boundary): unique face number per half-face -- interior faces in the order the
 *      cell walk discovers them, boundary faces after them; out: face_index[H], *num_faces, *max_faces_per_cell.
 * eu_post_process_fluxes: IncompFlowSolverHybrid::postProcessFluxes (mimetic/IncompFlowSolverHybrid.hpp:707-807): makes
 *      the out-fluxes of the two half-faces of a face exactly antisymmetric (periodic partners via partner_face[face],
 *      -1 or NULL = none), in place; *max_modification = the largest change.  Produces the hf_flux array that
 *      eu_transport_solve consumes.
 * eu_write_field: writeField (common/SimulatorUtilities.hpp:288-298), the per-step saturation file of the drivers. */
int eu_build_face_indices(int n_cells, const int* hf_offset, const int* hf_neighbour, int* face_index, int* num_faces,
                          int* max_faces_per_cell);
int eu_post_process_fluxes(int n_cells, const int* hf_offset, const int* hf_neighbour, const int* face_index, int num_faces,
                           const int* partner_face, double* hf_flux, double* max_modification);
int eu_write_field(const double* field, long long n, const char* filename);

/* Test hook (no device needed): the work units of the box kernel's sweep for a grid of nx x ny cells per plane, tiles of
 * tx x ty, own planes [z_lo, z_hi), bnd_lo / bnd_hi boundary planes whose cells a neighbour rank keeps as ghosts, and
 * grid_blocks resident blocks; spans != 0 asks for the equal-span partition (EU_BOX_UNITS=spans), lz > 0 fixes the chunk
 * length.  units4[4 i ..] = {x0 | y0 << 16, first plane, last plane + 1, flags (bit 0 / 1: pushes to the rank below /
 * above)} block after block, start[0 .. *n_blocks] the blocks' ranges.  Returns the number of units, -1 on bad arguments
 * (or a too small buffer).  No reference counterpart: the reference walks cells in a serial loop. */
int eu_debug_box_units(int nx, int ny, int tx, int ty, int z_lo, int z_hi, int bnd_lo, int bnd_hi, int grid_blocks,
                       int spans, int lz, int* units4, int max_units, int* start, int max_start, int* n_blocks);

#ifdef __cplusplus
}
#endif
#endif /* EULER_B200_H */
