// Internal declarations shared by the translation units of libeuler_b200.so.
//
//   eu_api.cu     host side of the C ABI (include/euler_b200.h): upload, step-count logic,
//                 retry loop, reports                      -- EulerUpstream_impl.hpp:95-218
//   eu_setup.cu   one-time structure building + STRICT arithmetic kernels, compiled with
//                 -fmad=false (bit-identical to the reference's operation order)
//   eu_fast.cu    the FAST substep kernel (pre-contracted face scalars, FMA allowed)
//
// Data layout in HBM (per rank; "local" = own + ghost cells in ascending global order):
//   "fat" static arrays, exactly what the host uploaded (kept for re-contraction when gravity
//   or the method flags change and for STRICT mode):
//       hf_offset[n+1] | hf_nbr[H] | hf_area[H] | hf_normal[3H] | hf_centroid[3H]
//       cell_volume[n] | cell_centroid[3n] | poro[n] | perm[9n] | rock[n]
//       bnd_kind/sat/partner_hf/partner_cell[nb]      (hf_nbr <= -2 encodes -2-bnd_index)
//   derived, STRICT:  owner_hf[H], strict list (int2 {owner hf, lo cell}) per half-face,
//                     porevol[n]
//   derived, FAST:    SELL-32 records int2 {nbr|code, face id} (slot-major inside a slice of
//                     32 cells: record of (slot j, lane l) at base[s] + 32*j + l),
//                     unique-face arrays q[F], G[F], T[F] in (slice, slot, lane) order,
//                     per cell porevol[n], pcscale[n], rock8[n]
//   state:            S[2][n] ping-pong, pc[2][n], hf_flux[H]
#ifndef EU_INTERNAL_H
#define EU_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/euler_b200.h"

#define EU_MAX_ROCKS 16          /* rock types the caller may pass */
#define EU_MAX_TABLES 48         /* curve sets in shared memory: one per rock, or three (x, y, z) for tensor mobility */
#define EU_SLICE 32
#define EU_REC_PAD (-1)

// ---- rock tables in device memory ----------------------------------------------------------
struct EuTablesDev {
    int kind;                 // EU_MOB_*
    int n_rocks;
    int n_nodes_total;
    int use_j;
    double sigma_cos_theta;
    double visc[2];
    double delta_rho;
    const int* offset;        // n_rocks+1
    const double* s;          // nodes
    const double* cols[7];    // raw columns (STRICT)
    // FAST (scalar mobility): per table interval, mobility = kr/visc and J in intercept/slope form
    // y(s) = a + b s (see eu_fast.cu), the interval bounds and a bucket index for the interval search.
    double inv_visc[2];
    int n_buckets;            // per rock; chosen so that a bucket holds at most one interior node
    const double* fcoef;      // 4 per node: a_w, b_w, a_o, b_o
    const double* fjcoef;     // 2 per node: a_J, b_J
    const double* fxb;        // node abscissae, last node of each rock = +inf
    const unsigned char* fbucket;   // n_rocks * n_buckets
};

// ---- grid / structure pointers handed to kernels -----------------------------------------
struct EuGridDev {
    int n_local;              // own + ghost
    int own_lo, own_hi;       // local index range of the cells this rank updates
    int cell_global0;         // global id = local id + cell_global0 only when contiguous (single range)
    long long H;
    const int* hf_offset;
    const int* hf_nbr;        // local neighbour id, or -2-bnd for boundary half-faces
    const double* hf_area;
    const double* hf_normal;
    const double* hf_centroid;
    const double* cell_volume;
    const double* cell_centroid;
    const double* poro;
    const double* perm;
    const int* rock;
    const int* bnd_kind;
    const double* bnd_sat;
    const int* bnd_partner_hf;     // local half-face index of the periodic partner
    const int* bnd_partner_cell;   // local cell of the periodic partner
    const int* local_to_global;    // n_local
};

struct EuStrictDev {
    const int2* list;         // per half-face slot of a cell: {owner hf, lo cell}, strict order
    const double* porevol;    // volume*poro
};

// Slice classes (FAST mode).  A slice of 32 own cells whose regular slots come in (-d, +d) pairs belongs to a
// class: the per-slot neighbour / face offsets are the same for every slice of the class, so the kernel keeps them
// in shared memory and never reads the slice's descriptors.  Slots are reordered so that slot 2p is the negative and
// slot 2p+1 the positive offset of pair p, and the pair with the largest offset that is a whole number of slices sits
// in slots 4/5: the kernel *marches* along it (slice s, s + D/32, ...), carrying the flux of face 5 of one cell over
// as face 4 of the next (the same face, evaluated once), together with the neighbour's saturation and mobilities.
// Slots that are not regular in this class (boundary faces, fault faces) are explicit: the regular code runs on a
// dummy face with zero flux for them (face id F, one past the real faces) and the real face is added from the slice's
// SELL records afterwards.  Slices wider than six slots (cells with split faces next to a fault plane) belong to a
// class as well: their slots 6.. are explicit.
#define EU_MAX_CLASSES 40
#define EU_ITEM_GENERIC 0xffff
struct EuSliceClass {
    int nb_off[6];            // regular slot: neighbour = cell + nb_off;  explicit: 0
    int fid_mul[6];           // regular slot: face = cell*1 + fid_off;    explicit: cell*0 + (the all-zero face F)
    int fid_off[6];
    int D;                    // march stride in cells (multiple of 32); 0 = items of this class have length 1
    int rec_mask;             // bit j: slot j of the SELL records is an explicit face of this class (j < 31)
    int pad[2];
};

struct EuFastDev {
    int n_slices;             // slices over all local cells
    int n_local;              // face id = plane*n_local + owner cell, plane = the owner's local face slot
    const int2* items;        // work items of the interior slice range: {first slice, length | class << 16}
    int n_items;
    const EuSliceClass* classes;   // n_classes <= EU_MAX_CLASSES
    int n_classes;
    const int* slice_base;    // n_slices+1, record offsets (multiples of 32)
    const int2* rec;          // explicit records, read only for the slots a descriptor marks irregular
    const int2* desc;         // per (slice, slot) at slice_base/32 + slot: {d, k}.  k >= 0: every lane's neighbour
                              // is cell + d and its face lives in plane k (regular slot, no record needed);
                              // k == -1: irregular slot, use rec; k == -2: no face in this slot
    const double2* qg;        // per unique face {q, G}: compacted flux of the current transportSolve, gravity scalar
    const double* T;
    const double* nn;         // n.n per face, or NULL when all normals are unit to 1e-13
    const double* inv_porevol;   // 1/(volume*poro)
    const double* pcscale;
    const unsigned char* rock8;
    const unsigned char* axis8;  // per unique face: axis of its (axis-aligned) normal -- FAST tensor mobility only, else NULL
    const double* fv;            // FAST tensor mobility on oblique normals: fv[k*fv_stride + face], k = 0..2 Gv, 3..5 n_k^2,
    long long fv_stride;         // 6..8 Tv (k_contract_t3); else NULL
    long long F;
    int prefetch;             // marches request a later cell's lines into L2: distance in march steps, 0 = off
    int l2_hint;              // streaming arrays are loaded with an evict-first L2 policy
};

struct EuStepArgs {
    double dt;
    int method_viscous, method_gravity, method_capillary;
    int check_sat, clamp_sat;
    int substep;              // index inside the attempt, for the failure key
    int n_src;
    const int* src_cell;      // local cell ids, ascending
    const double* src_rate;
    const double* S_in;
    double* S_out;
    const double* pc_in;      // FAST: pc(S_in) for all local cells; STRICT: scratch filled by k_strict_pc
    double* pc_out;
    const double2* lam_in;    // FAST built with -DEU_STORED_LAM: {lambda_w, lambda_o}(S_in) of the own cells; else NULL
    double2* lam_out;
    double* residual_out;     // optional
    unsigned long long* fail_key;   // min over failing (substep<<32 | local cell)
    double gravity[3];
};

// Halo exchange fused into the FAST substep kernel (one process per GPU, peers mapped with CUDA IPC).
// The own slices next to a slab boundary form up to two ranges at the two ends of the own slice range; they
// are processed FIRST, their new saturations are also stored into the neighbour's ghost slots (NVLink P2P),
// and the last warp to finish a neighbour's share publishes the epoch in that neighbour's flag word.
// Before touching ghosts the boundary warps wait for the neighbours' flags of the previous epoch.
struct EuHaloDev {
    int enabled;
    int a_hi;                  // boundary range A = [slice_lo, a_hi)
    int b_lo;                  // boundary range B = [b_lo, slice_hi)
    const int* dst[2];         // per cell of the range (from its first slice): ghost slot in the peer, or -1
    double* peer_S[2];         // the peer's S_out buffer
    double* peer_pc[2];        // the peer's pc_out buffer (FAST + capillary) or NULL
    unsigned* counter[2];      // finished-slice counter (shared when both ranges feed the same peer)
    unsigned total[2];         // slices that make the counter complete
    unsigned* peer_flag[2];    // flag word in the peer, written with `epoch`
    const unsigned* my_flags;  // [world]: flag words the peers write
    int n_wait;
    int wait_rank[2];
    unsigned epoch;
    long long timeout_cycles;
    int* err_flag;
};

// ---- launchers implemented in eu_setup.cu (all -fmad=false) --------------------------------
struct EuSetupOut;   // opaque to eu_fast.cu
void eu_launch_translate_nbr(int* hf_nbr, long long H, const int* range_first, const int* range_count,
                             const int* range_local, int n_ranges, cudaStream_t st);
void eu_launch_owner(const EuGridDev& g, int* owner_hf, int* err_flag, cudaStream_t st);
void eu_launch_strict_list(const EuGridDev& g, const int* owner_hf, int2* list, cudaStream_t st);
void eu_launch_porevol(const EuGridDev& g, double* porevol, cudaStream_t st);
void eu_launch_canonical_slots(const EuGridDev& g, unsigned char* slot_of_hf, int* slice_width, cudaStream_t st);
void eu_launch_assign_fid(const EuGridDev& g, const int* owner_hf, const unsigned char* slot_of_hf, const int ax[3], int box,
                          int* fid_of_hf, cudaStream_t st);
// votes for candidate positive neighbour offsets (n_cand <= 16): how many interior faces have neighbour = cell + cand[q]
void eu_launch_offset_votes(const EuGridDev& g, const int* cand, int n_cand, unsigned long long* votes, cudaStream_t st);
void eu_launch_cell_mask(const EuGridDev& g, const int* slice_base, const int2* rec, unsigned short* cmask, int* list, int* count,
                         cudaStream_t st);
void eu_launch_build_records(const EuGridDev& g, const int* owner_hf, const int* fid_of_hf, const unsigned char* slot_of_hf,
                             const int* slice_base, int2* rec, int2* desc, int* n_regular_slots, cudaStream_t st);
// Ga (3*n_local doubles, may be NULL): G of the faces in the axis planes as a separate array (box kernel);
// gmask_out (device int, may be NULL): bit a set when axis plane a has a non-zero G
void eu_launch_contract(const EuGridDev& g, const EuTablesDev& t, const int* owner_hf, const int* fid_of_hf,
                        const double gravity[3], int method_gravity, double* G, double* T, double* nn,
                        double* nn_maxdev, double* Ga, int* gmask_out, cudaStream_t st);
// tensor mobility in FAST mode: are all face normals axis-aligned (flag[0] |= 1 if not)?  axis per unique face
void eu_launch_contract_t3(const EuGridDev& g, const EuTablesDev& t, const int* owner_hf, const int* fid_of_hf,
                           const double gravity[3], int method_gravity, double* fv, long long stride, cudaStream_t st);
void eu_launch_axis_check(const EuGridDev& g, int* flag, cudaStream_t st);
void eu_launch_face_axis(const EuGridDev& g, const int* owner_hf, const int* fid_of_hf, unsigned char* axis8, cudaStream_t st);
void eu_launch_pcscale(const EuGridDev& g, const EuTablesDev& t, double* pcscale, unsigned char* rock8, double* inv_porevol, cudaStream_t st);
// CFL terms (CflCalculator.hpp); results are block minima reduced to out[0]
// qa (3*n_local doubles, may be NULL): the flux of the faces in the axis planes as a separate array (box kernel)
void eu_launch_cfl_velocity_compact(const EuGridDev& g, double cfl_factor, const double* hf_flux, const int* fid_of_hf,
                                    double* q, double* qa, double* block_min, int* zero_flag, double* out, cudaStream_t st);
void eu_launch_cfl_gravity(const EuGridDev& g, const EuTablesDev& t, double cfl_factor, const double gravity[3],
                           double* block_min, double* out, cudaStream_t st);
void eu_launch_cfl_capillary(const EuGridDev& g, double cfl_factor, double* block_min, double* out, cudaStream_t st);
int eu_cfl_blocks(int n_cells);
// STRICT substep
void eu_launch_strict_pc(const EuGridDev& g, const EuTablesDev& t, const double* S, double* pc, cudaStream_t st);
void eu_launch_strict_step(const EuGridDev& g, const EuTablesDev& t, const EuStrictDev& s, const double* hf_flux,
                           const EuStepArgs& a, cudaStream_t st);
// halo exchange (eu_setup.cu): peer-to-peer stores + epoch flags
void eu_launch_halo_push(const int* send_src, const int* send_dst, int n, const double* S_local, double* S_peer,
                         const double* pc_local, double* pc_peer, unsigned* block_counter, unsigned* peer_flag,
                         unsigned epoch, cudaStream_t st);
void eu_launch_halo_wait(const unsigned* my_flags, const int* wait_ranks, int n_wait, unsigned epoch,
                         long long timeout_cycles, int* err_flag, cudaStream_t st);
// ---- eu_diag.cu (-fmad=false): diagnostics on resident data (common/SimulatorUtilities.hpp) -------------
void eu_launch_cell_velocity(const EuGridDev& g, const double* hf_flux, double* out, cudaStream_t st);
void eu_launch_fractional_flow(const EuGridDev& g, const EuTablesDev& t, const double* S, double* out, cudaStream_t st);
void eu_launch_phase_velocities(const EuGridDev& g, const EuTablesDev& t, const double* S, const double* cell_v,
                                double* vw, double* vo, cudaStream_t st);
// ---- eu_fast.cu ------------------------------------------------------------------------------
void eu_launch_fast_state(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const double* S, double* pc,
                          double2* lam, int lo, int hi, cudaStream_t st);
void eu_launch_fast_step(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                         const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, cudaStream_t st);
void eu_launch_fast_step_t3(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                            int slice_lo, int slice_hi, int n_sms, cudaStream_t st);
void eu_launch_ghost_adjacent(const EuGridDev& g, int* out4, cudaStream_t st);
size_t eu_fast_smem_bytes(const EuTablesDev& t);
// box kernel (eu_tile.cuh): plane sweep over tiles with TMA-staged operands, for local numberings that are a box
struct EuBoxPlan;
EuBoxPlan* eu_box_plan_create(int nx, int ny, int nz, int z_lo, int z_hi, double* S0, double* S1, double* pc0, double* pc1,
                              double* qa, double* Ga, double* T, const unsigned short* cmask, const int* irr_cells, int n_irr, double* acc_irr,
                              double* inv_porevol, int n_sms);
void eu_box_plan_destroy(EuBoxPlan* p);
void eu_box_plan_set_gravity_mask(EuBoxPlan* p, int mask);    // bit a: axis plane a has a non-zero G somewhere
void eu_box_plan_info(const EuBoxPlan* p, int out[6]);      // tile x, tile y, units, boundary units A, B, threads per block
// builds (or reuses) the work units for `bnd_lo` / `bnd_hi` boundary planes at the two ends of the own range; info as above
int eu_box_plan_units(EuBoxPlan* p, int bnd_lo, int bnd_hi, bool capillary, int info[6]);
int eu_launch_box_step(EuBoxPlan* p, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                       const EuHaloDev& halo, int cur, int slice_lo, int slice_hi, int bnd_lo, int bnd_hi, cudaStream_t st);
int eu_fast_warps_per_sm(bool capillary);   // resident warps per SM of the substep kernel variant in use
bool eu_fast_uses_stored_lam();      // build-time choice of eu_fast.cu: per-cell mobility pairs kept in HBM

#endif
