// Host side of the C ABI (include/euler_b200.h).
//
// Mirrors, call by call, what opm-porsol's EulerUpstream does on the host:
//   eu_set_params            EulerUpstream::init(param)            euler/EulerUpstream_impl.hpp:95-108
//   eu_grid_* / eu_set_fluid EulerUpstream::initObj(g, r, b)       :119-127 (+ Residual_impl.hpp:391-432)
//   eu_transport_solve       EulerUpstream::transportSolve         :151-218
//       CFL terms            computeCflTime                        :263-331
//       step count           :163-172, retry loop :181-213, range check :336-349
// The arithmetic runs in the kernels of eu_setup.cu (STRICT, setup, CFL) and eu_fast.cu (FAST).
#include "eu_internal.h"

#include <nvtx3/nvToolsExt.h>     // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, no-ops otherwise

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <unistd.h>
#include <limits>
#include <map>
#include <algorithm>
#include <string>
#include <vector>

namespace {

// NVTX range over a scope (the reference has its own StopWatch prints around the same phases,
// EulerUpstream_impl.hpp:185-186,214-217)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

std::string g_create_error;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count*sizeof(T));
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct Range { int first, count, local; };

} // namespace

struct eu_solver {
    eu_config cfg;
    eu_params par;
    int mode;                 // resolved: EU_MODE_STRICT or EU_MODE_FAST
    int n_sms;
    cudaStream_t st;
    std::string err;

    // ---- upload bookkeeping
    bool grid_open = false, grid_ready = false, fluid_set = false;
    int n_global = 0, n_local_expected = 0, n_local = 0;
    long long H_expected = 0, H = 0;
    std::vector<Range> ranges;
    std::vector<int> h_hf_offset;
    std::vector<int> b_hf, b_kind, b_pcell, b_pface;
    std::vector<double> b_sat;
    bool any_rock_ids = false;
    int own_lo = 0, own_hi = 0;

    // ---- fat static
    DevBuf<int> d_hf_offset, d_hf_nbr, d_rock, d_bnd_kind, d_bnd_phf, d_bnd_pcell, d_l2g;
    DevBuf<double> d_hf_area, d_hf_normal, d_hf_centroid, d_cell_volume, d_cell_centroid, d_poro, d_perm, d_bnd_sat;
    DevBuf<int> d_range_first, d_range_count, d_range_local;
    // ---- fluid
    eu_fluid fluid;
    std::vector<int> h_tab_offset;
    std::vector<double> h_tab_s, h_tab_cols[7];
    DevBuf<int> d_tab_offset;
    DevBuf<unsigned char> d_fbucket;
    DevBuf<double> d_tab_s, d_tab_cols[7], d_fcoef, d_fjcoef, d_fxb;
    int n_buckets = 0;
    bool fast_tables_ok = true;
    EuTablesDev tab;
    // FAST curve sets: one per rock (scalar mobility) or three per rock, x/y/z (diagonal tensor mobility on grids with
    // axis-aligned face normals); tabf is `tab` with the set count / offsets of the FAST tables
    EuTablesDev tabf;
    DevBuf<int> d_tab_offset_fast;
    int tensor_fast = 0;               // 0: scalar class; 1: tensor class, axis-aligned normals; 2: tensor class, general normals
    DevBuf<unsigned char> d_axis8;
    DevBuf<double> d_fv;               // tensor_fast == 2: per-face axis vectors (k_contract_t3), 9 x (F + 1)
    // ---- derived
    DevBuf<int> d_owner_hf, d_fid_of_hf, d_slice_base, d_flags;
    DevBuf<int2> d_strict_list, d_rec, d_desc;
    double regular_fraction = 0.0;
    // work items of the FAST kernel (slice classes + marches), built for the slice range [items_lo, items_hi)
    std::vector<int> h_slice_base;
    DevBuf<int2> d_items;
    DevBuf<EuSliceClass> d_classes;
    int n_items = 0, n_classes = 0, items_lo = -1, items_hi = -1, items_variant = -1;
    double class_fraction = 0.0;       // share of the own slices that belong to a slice class
    int plan_max_len = 0;
    double plan_mean_len = 0.0;
    DevBuf<double> d_porevol, d_inv_porevol, d_pcscale, d_T, d_nn;
    DevBuf<double2> d_lam[2];          // per cell {lambda_w, lambda_o} of the state in d_S[k] (FAST)
    DevBuf<double2> d_qg;              // per unique face {flux of the current transportSolve, G}
    DevBuf<double> d_qa, d_Ga;         // the axis planes' q and G as separate arrays [3][n_local] (box kernel: G only where needed)
    DevBuf<unsigned char> d_rock8;
    long long F = 0;
    int axis[3] = { 0, 0, 0 };         // neighbour offsets of the axis planes of the face arrays (0 = none)
    int max_slots = 0;                 // widest slice (canonical slots)
    bool box_ok = false;               // local numbering is a box of whole planes (see eu_grid_end)
    EuBoxPlan* box = nullptr;          // box kernel (eu_tile.cuh): tiles, tensor maps, work units; nullptr = slice-class kernel
    bool ran_box = false;              // the last FAST substep ran the box kernel
    int box_enabled = 1;               // EU_BOX=0 keeps the slice-class kernel (tuning / A-B knob)
    DevBuf<unsigned short> d_cmask;    // per cell: record slots holding faces outside the axis planes
    DevBuf<int> d_irr_cells;           // own cells with such faces (work list of k_box_irregular)
    DevBuf<double> d_acc_irr;          // per cell: their summed contribution in the current substep
    int n_slices = 0;
    bool use_nn = false;
    int prefetch = 1;                  // EU_PREFETCH=<march steps ahead>, 0 turns the L2 prefetch of the marches off (tuning knob)
    int l2_hint = 1;                   // EU_L2_HINT=0 turns the evict-first L2 policy of the streaming arrays off (tuning knob)
    bool contracted = false;
    double contracted_gravity[3] = { 0, 0, 0 };
    int contracted_mg = -1;
    // ---- state
    DevBuf<double> d_S[2], d_pc[2], d_S_init, d_hf_flux, d_residual, d_block_min, d_scalars;
    DevBuf<double> d_diag;             // scratch of the diagnostics (eu_diag.cu), allocated on first use
    // host buffers of the caller registered (page-locked) on first use, see include/euler_b200.h
    struct Pinned { const void* p; size_t bytes; unsigned long long stamp; };
    std::vector<Pinned> pinned;
    unsigned long long pin_clock = 0;
    bool pin_cache = false;            // opt-in (EU_PIN_CACHE=1): the caller must not free a registered buffer while the solver lives
    DevBuf<unsigned long long> d_fail_key;
    DevBuf<int> d_src_cell;
    DevBuf<double> d_src_rate;
    int cur = 0;
    bool state_ready = false;
    // cached CFL terms
    bool cfl_cap_valid = false, cfl_grav_valid = false;
    double cfl_cap = 1e100, cfl_grav = 1e100, cfl_grav_gravity[3] = { 0, 0, 0 };
    cudaEvent_t ev0, ev1;
    // comm (one process per GPU): ghost saturations are pushed into the neighbours' buffers
    eu_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    struct Peer {
        int rank = -1;
        bool recv = false;                 // this peer pushes into my ghosts
        int n_send = 0;
        double* S[2] = { nullptr, nullptr };
        double* pc[2] = { nullptr, nullptr };
        unsigned* flags = nullptr;         // the peer's flag array (I write entry [my rank])
        void* opened[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
        DevBuf<int> src, dst;
        DevBuf<unsigned> counter;
    };
    std::vector<Peer*> peers;
    // fused exchange (FAST mode): boundary slice ranges at the two ends of the own range
    bool fused_ok = false;
    int fused_a_hi = 0, fused_b_lo = 0;
    int fused_down_max = -1, fused_up_min = INT_MAX;       // extreme local cells of the two boundary ranges
    int fused_peer[2] = { -1, -1 };        // index into peers
    unsigned fused_total[2] = { 0, 0 };
    DevBuf<int> d_fused_dst[2];
    DevBuf<unsigned> d_fused_dummy;
    DevBuf<unsigned> d_comm_flags;         // [world_size], written by the peers
    DevBuf<int> d_wait_ranks;
    int n_wait = 0;
    unsigned epoch = 0;
    bool comm_ready = false;
    std::vector<int> ghost_global, ghost_local;

    EuGridDev grid() const
    {
        EuGridDev g;
        g.n_local = n_local; g.own_lo = own_lo; g.own_hi = own_hi;
        g.cell_global0 = ranges.size() == 1 ? ranges[0].first : -1;
        g.H = H;
        g.hf_offset = d_hf_offset.p; g.hf_nbr = d_hf_nbr.p; g.hf_area = d_hf_area.p;
        g.hf_normal = d_hf_normal.p; g.hf_centroid = d_hf_centroid.p;
        g.cell_volume = d_cell_volume.p; g.cell_centroid = d_cell_centroid.p;
        g.poro = d_poro.p; g.perm = d_perm.p; g.rock = d_rock.p;
        g.bnd_kind = d_bnd_kind.p; g.bnd_sat = d_bnd_sat.p; g.bnd_partner_hf = d_bnd_phf.p;
        g.bnd_partner_cell = d_bnd_pcell.p; g.local_to_global = d_l2g.p;
        return g;
    }
    EuFastDev fast() const
    {
        EuFastDev f;
        f.n_slices = n_slices; f.n_local = n_local; f.items = d_items.p; f.n_items = n_items;
        f.classes = d_classes.p; f.n_classes = n_classes; f.slice_base = d_slice_base.p; f.rec = d_rec.p; f.desc = d_desc.p;
        f.qg = d_qg.p; f.T = d_T.p; f.nn = use_nn ? d_nn.p : nullptr;
        f.inv_porevol = d_inv_porevol.p; f.pcscale = d_pcscale.p; f.rock8 = d_rock8.p; f.F = F;
        f.axis8 = tensor_fast == 1 ? d_axis8.p : nullptr;
        f.fv = tensor_fast == 2 ? d_fv.p : nullptr;
        f.fv_stride = (long long)(d_fv.n/9);
        f.prefetch = prefetch;
        f.l2_hint = l2_hint;
        return f;
    }
    int global_to_local(int gcell) const
    {
        for (const Range& r : ranges) {
            if (gcell >= r.first && gcell < r.first + r.count) return r.local + (gcell - r.first);
        }
        return -1;
    }
    int local_to_global(int lcell) const
    {
        for (const Range& r : ranges) {
            if (lcell >= r.local && lcell < r.local + r.count) return r.first + (lcell - r.local);
        }
        return -1;
    }
};

namespace {

int fail(eu_handle h, int code, const std::string& msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define EU_CUDA(h, call)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            return fail(h, EU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
        }                                                                                      \
    } while (0)

// Page-lock a caller buffer that keeps coming back (include/euler_b200.h, "flux hand-off").  Best effort: any
// failure leaves the buffer pageable and the copy correct.
void unpin_all(eu_handle h)
{
    for (const eu_solver::Pinned& e : h->pinned) cudaHostUnregister(const_cast<void*>(e.p));
    h->pinned.clear();
    cudaGetLastError();
}

void pin_host(eu_handle h, const void* p, size_t bytes)
{
    if (!h->pin_cache || !p || bytes < (size_t(1) << 20)) return;
    for (eu_solver::Pinned& e : h->pinned) {
        if (e.p == p && e.bytes >= bytes) { e.stamp = ++h->pin_clock; return; }
    }
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) return;   // pinned already
    cudaGetLastError();
    // a stale registration that overlaps this buffer (the application reallocated its vector) must go first
    for (size_t i = 0; i < h->pinned.size();) {
        const char* a = static_cast<const char*>(h->pinned[i].p);
        const char* b = static_cast<const char*>(p);
        if (a < b + bytes && b < a + h->pinned[i].bytes) {
            cudaHostUnregister(const_cast<void*>(h->pinned[i].p));
            h->pinned.erase(h->pinned.begin() + long(i));
        } else {
            ++i;
        }
    }
    if (h->pinned.size() >= 4) {
        size_t oldest = 0;
        for (size_t i = 1; i < h->pinned.size(); ++i) if (h->pinned[i].stamp < h->pinned[oldest].stamp) oldest = i;
        cudaHostUnregister(const_cast<void*>(h->pinned[oldest].p));
        h->pinned.erase(h->pinned.begin() + long(oldest));
    }
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable) == cudaSuccess)
        h->pinned.push_back(eu_solver::Pinned{ p, bytes, ++h->pin_clock });
    cudaGetLastError();
}

template <class T>
int upload(eu_handle h, DevBuf<T>& dst, size_t offset, const T* src, size_t count)
{
    if (count == 0) return EU_OK;
    if (offset + count > dst.n) return fail(h, EU_ERR_ARG, "upload past the announced size");
    EU_CUDA(h, cudaMemcpyAsync(dst.p + offset, src, count*sizeof(T), cudaMemcpyHostToDevice, h->st));
    return EU_OK;
}

template <class T>
int upload_vec(eu_handle h, DevBuf<T>& dst, const std::vector<T>& v)
{
    EU_CUDA(h, dst.alloc(v.size()));
    if (v.empty()) return EU_OK;
    EU_CUDA(h, cudaMemcpyAsync(dst.p, v.data(), v.size()*sizeof(T), cudaMemcpyHostToDevice, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    return EU_OK;
}

// opm-core tableIndex for a bucket edge (host twin of eu_strict_math.cuh / the oracle shim)
int table_index_host(const double* x, int size, double v)
{
    int n = size - 1;
    if (n < 2) return 0;
    int jl = 0, ju = n;
    const bool ascend = x[n] > x[0];
    while (ju - jl > 1) {
        int jm = (ju + jl)/2;
        if ((v >= x[jm]) == ascend) jl = jm; else ju = jm;
    }
    return jl;
}

void set_tables_struct(eu_handle h)
{
    EuTablesDev& t = h->tab;
    t.kind = h->fluid.mobility_kind;
    t.n_rocks = h->fluid.n_rocks;
    t.n_nodes_total = int(h->h_tab_s.size());
    t.use_j = h->fluid.use_jfunction_scaling;
    t.sigma_cos_theta = h->fluid.sigma_cos_theta;
    t.visc[0] = h->fluid.viscosity[0]; t.visc[1] = h->fluid.viscosity[1];
    t.delta_rho = h->fluid.density[0] - h->fluid.density[1];       // densityDifference()
    t.offset = h->d_tab_offset.p; t.s = h->d_tab_s.p;
    for (int k = 0; k < 7; ++k) t.cols[k] = h->d_tab_cols[k].p;
    t.inv_visc[0] = 1.0/h->fluid.viscosity[0]; t.inv_visc[1] = 1.0/h->fluid.viscosity[1];
    t.n_buckets = h->n_buckets;
    t.fcoef = h->d_fcoef.p; t.fjcoef = h->d_fjcoef.p; t.fxb = h->d_fxb.p; t.fbucket = h->d_fbucket.p;
    // the view the FAST kernels get: curve sets instead of rocks
    h->tabf = t;
    const int reps = h->fluid.mobility_kind == EU_MOB_DIAGONAL ? 3 : 1;
    h->tabf.n_rocks = h->fluid.n_rocks*reps;
    h->tabf.n_nodes_total = int(h->d_fxb.n);
    h->tabf.offset = h->d_tab_offset_fast.p;
}

int ensure_contracted(eu_handle h, const double gravity[3])
{
    if (h->mode != EU_MODE_FAST) return EU_OK;
    const int mg = h->par.method_gravity ? 1 : 0;
    if (h->contracted && h->contracted_mg == mg && std::memcmp(h->contracted_gravity, gravity, 3*sizeof(double)) == 0)
        return EU_OK;
    eu_launch_contract(h->grid(), h->tab, h->d_owner_hf.p, h->d_fid_of_hf.p, gravity, mg, reinterpret_cast<double*>(h->d_qg.p) + 1, h->d_T.p, h->d_nn.p,
                       h->d_scalars.p + 8, h->d_Ga.p, h->d_Ga.p ? h->d_flags.p + 2 : nullptr, h->st);
    double maxdev = 0.0;
    int gmask = 7;
    EU_CUDA(h, cudaMemcpyAsync(&maxdev, h->d_scalars.p + 8, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (h->d_Ga.p) EU_CUDA(h, cudaMemcpyAsync(&gmask, h->d_flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    eu_box_plan_set_gravity_mask(h->box, gmask);
    EU_CUDA(h, cudaGetLastError());
    h->use_nn = maxdev > 1e-13;     // non-unit normals: keep the n.n factor of the viscous term
    if (h->use_nn && h->d_nn.n == 0) {
        EU_CUDA(h, h->d_nn.alloc(size_t(std::max<long long>(h->F, 1)) + 1));
        EU_CUDA(h, cudaMemsetAsync(h->d_nn.p, 0, h->d_nn.n*sizeof(double), h->st));
        eu_launch_contract(h->grid(), h->tab, h->d_owner_hf.p, h->d_fid_of_hf.p, gravity, mg, reinterpret_cast<double*>(h->d_qg.p) + 1, h->d_T.p, h->d_nn.p,
                           h->d_scalars.p + 8, nullptr, nullptr, h->st);
        EU_CUDA(h, cudaStreamSynchronize(h->st));
    }
    if (h->tensor_fast == 2)
        eu_launch_contract_t3(h->grid(), h->tab, h->d_owner_hf.p, h->d_fid_of_hf.p, gravity, mg, h->d_fv.p, (long long)(h->d_fv.n/9), h->st);
    h->contracted = true;
    h->contracted_mg = mg;
    std::memcpy(h->contracted_gravity, gravity, 3*sizeof(double));
    return EU_OK;
}

int upload_sources(eu_handle h, int n_src, const int* src_cell, const double* src_rate, int* n_local_src)
{
    std::vector<int> lc;
    std::vector<double> lr;
    for (int i = 0; i < n_src; ++i) {
        if (i > 0 && src_cell[i] <= src_cell[i - 1]) return fail(h, EU_ERR_ARG, "source cells must be strictly ascending");
        const int l = h->global_to_local(src_cell[i]);
        if (l >= h->own_lo && l < h->own_hi) { lc.push_back(l); lr.push_back(src_rate[i]); }
    }
    *n_local_src = int(lc.size());
    if (lc.empty()) return EU_OK;
    if (h->d_src_cell.n < lc.size()) {
        EU_CUDA(h, h->d_src_cell.alloc(lc.size()));
        EU_CUDA(h, h->d_src_rate.alloc(lc.size()));
    }
    EU_CUDA(h, cudaMemcpyAsync(h->d_src_cell.p, lc.data(), lc.size()*sizeof(int), cudaMemcpyHostToDevice, h->st));
    EU_CUDA(h, cudaMemcpyAsync(h->d_src_rate.p, lr.data(), lr.size()*sizeof(double), cudaMemcpyHostToDevice, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));   // lc/lr go out of scope
    return EU_OK;
}

EuStepArgs step_args(eu_handle h, double dt, const double gravity[3], int n_src, int substep)
{
    EuStepArgs a;
    a.dt = dt;
    a.method_viscous = h->par.method_viscous; a.method_gravity = h->par.method_gravity;
    a.method_capillary = h->par.method_capillary;
    a.check_sat = h->par.check_sat; a.clamp_sat = h->par.clamp_sat;
    a.substep = substep;
    a.n_src = n_src; a.src_cell = h->d_src_cell.p; a.src_rate = h->d_src_rate.p;
    a.S_in = h->d_S[h->cur].p; a.S_out = h->d_S[h->cur ^ 1].p;
    a.pc_in = h->d_pc[h->cur].p; a.pc_out = h->d_pc[h->cur ^ 1].p;
    a.lam_in = h->d_lam[h->cur].p; a.lam_out = h->d_lam[h->cur ^ 1].p;
    a.residual_out = nullptr;
    a.fail_key = h->d_fail_key.p;
    a.gravity[0] = gravity[0]; a.gravity[1] = gravity[1]; a.gravity[2] = gravity[2];
    return a;
}

bool fused_halo(eu_handle h)
{
    return h->cfg.world_size > 1 && h->comm_ready && h->mode == EU_MODE_FAST && h->fused_ok && h->tensor_fast != 2;
}

// ---- work items of the FAST kernel -----------------------------------------------------------------------
// Classifies the slices of [lo, hi) (EuSliceClass in eu_internal.h), chains slices of the same class along the
// march direction into items of at most `lmax` slices and orders the items so that the eight warps of a block
// work on eight neighbouring grid rows.  Everything is derived from the CSR adjacency as uploaded (through the
// per-slot descriptors of eu_setup.cu); a slice that fits no class stays a generic item.
struct ClassKey {
    int v[20];
    bool operator<(const ClassKey& o) const { return std::memcmp(v, o.v, sizeof(v)) < 0; }
};

int build_items(eu_handle h, int lo, int hi)
{
    // the plan depends on the slice range and on the kernel variant (resident warps per SM differ with the capillary term)
    const int variant_key = h->par.method_capillary != 0;
    if (h->items_lo == lo && h->items_hi == hi && h->items_variant == variant_key) return EU_OK;
    const std::vector<int>& base = h->h_slice_base;
    const size_t n_desc = size_t(base.back()/EU_SLICE);
    std::vector<int2> desc(n_desc + 1);
    EU_CUDA(h, cudaMemcpy(desc.data(), h->d_desc.p, n_desc*sizeof(int2), cudaMemcpyDeviceToHost));
    const int zero_face = int(std::max<long long>(h->F, 1));
    std::map<ClassKey, int> class_of_key;
    std::vector<EuSliceClass> classes;
    std::vector<unsigned short> cls(size_t(std::max(hi - lo, 0)), (unsigned short)EU_ITEM_GENERIC);
    const char* env_noclass = getenv("EU_NO_CLASSES");
    // the marches read the neighbours' stored mobilities, which exist for own cells only: slice classes are used when
    // no item touches a ghost cell (one rank, or the fused exchange whose boundary ranges take the generic path)
    const bool use_classes = !(env_noclass && atoi(env_noclass) != 0) && (h->cfg.world_size <= 1 || fused_halo(h)) && !h->tensor_fast;
    for (int s = lo; s < hi && use_classes; ++s) {
        if (s*EU_SLICE < h->own_lo || (s + 1)*EU_SLICE > h->own_hi) continue;
        const int width = (base[size_t(s) + 1] - base[size_t(s)])/EU_SLICE;
        if (width < 6 || width > 30) continue;
        const int2* d = &desc[size_t(base[size_t(s)]/EU_SLICE)];
        // regular slots in (-d, +d) pairs
        int mag[3] = { 0, 0, 0 }, neg[3] = { -1, -1, -1 }, pos[3] = { -1, -1, -1 }, npairs = 0;
        bool ok = true;
        for (int j = 0; j < 6 && ok; ++j) {
            if (d[j].y < 0) continue;
            const int m = std::abs(d[j].x);
            int p = 0;
            while (p < npairs && mag[p] != m) ++p;
            if (p == npairs) {
                if (npairs == 3) { ok = false; break; }
                mag[npairs++] = m;
            }
            int& member = d[j].x < 0 ? neg[p] : pos[p];
            if (member >= 0) ok = false;
            member = j;
        }
        if (!ok) continue;
        // march pair: the largest offset that is a whole number of slices; the other pairs by ascending offset
        int order[3] = { 0, 1, 2 };
        std::sort(order, order + npairs, [&](int a, int b) { return mag[a] < mag[b]; });
        int march = -1;
        for (int k = npairs - 1; k >= 0; --k) if (mag[order[k]] % EU_SLICE == 0) { march = order[k]; break; }
        int slot_of_pos[6] = { -1, -1, -1, -1, -1, -1 };      // position -> original regular slot
        int npos = 0;
        for (int k = 0; k < npairs; ++k) {
            const int p = order[k];
            if (p == march) continue;
            slot_of_pos[2*npos] = neg[p];
            slot_of_pos[2*npos + 1] = pos[p];
            ++npos;
        }
        if (march >= 0) { slot_of_pos[4] = neg[march]; slot_of_pos[5] = pos[march]; }
        ClassKey key;
        std::memset(&key, 0, sizeof(key));
        EuSliceClass c;
        std::memset(&c, 0, sizeof(c));
        for (int q = 0; q < 6; ++q) {
            const int j = slot_of_pos[q];
            if (j >= 0) {
                c.nb_off[q] = d[j].x;
                c.fid_mul[q] = 1;
                const long long off = (long long)d[j].y*h->n_local + (d[j].x > 0 ? 0 : d[j].x);
                c.fid_off[q] = int(off);
            } else {
                c.nb_off[q] = 0; c.fid_mul[q] = 0; c.fid_off[q] = zero_face;
            }
            key.v[3*q] = c.nb_off[q]; key.v[3*q + 1] = c.fid_mul[q]; key.v[3*q + 2] = c.fid_off[q];
        }
        for (int j = 0; j < 6; ++j) if (d[j].y == -1) c.rec_mask |= 1 << j;
        for (int j = 6; j < width; ++j) if (d[j].y != -2) c.rec_mask |= 1 << j;      // extra slots: always from the records
        c.D = (march >= 0 && pos[march] >= 0) ? mag[march] : 0;
        key.v[18] = c.rec_mask; key.v[19] = c.D;
        auto it = class_of_key.find(key);
        int id;
        if (it != class_of_key.end()) {
            id = it->second;
        } else {
            if (classes.size() >= EU_MAX_CLASSES) continue;
            id = int(classes.size());
            classes.push_back(c);
            class_of_key[key] = id;
        }
        cls[size_t(s - lo)] = (unsigned short)id;
    }
    // Chains along the march direction, cut into work items.  The persistent grid hands item v to warp v mod #warps, so
    // the kernel ends when the warp with the most march steps ends: the piece length L is chosen to minimise
    //     ceil(#items / #warps) * (mean piece length + head)
    // (head = the extra face and loads of the first cell of a march, about a third of a step) over L = 4..64, and every
    // chain is cut into ceil(n/L) pieces of nearly equal length.  A fixed L = 32 left 8 % of the warps' time idle on a
    // 128-plane slab (9.2 items per warp -> 10).  EU_MARCH_LEN overrides L.  Measured (profiles/README.md): +5-7 % on
    // single-rank grids of 32-128 planes, neutral at 256; on a 2-rank run the long pieces this rule picks there (63
    // steps) were 10 % slower than L = 32 -- plausibly because the warps of a block drift apart in z over a long march and
    // stop sharing their neighbours' lines in L1 -- so decomposed runs use the exact simulation below with short pieces.
    const int n_warps = h->n_sms*32;
    auto chain_length = [&](int s, int id, const std::vector<char>& taken) {
        int n = 1;
        const int step = classes[size_t(id)].D/EU_SLICE;
        if (step <= 0) return 1;
        for (int nxt = s + step; n < 65535 && nxt < hi && !taken[size_t(nxt - lo)] && cls[size_t(nxt - lo)] == id; nxt += step) ++n;
        return n;
    };
    int lmax = 32;
    {
        // EU_ITEM_RULE = fixed | sim | auto overrides the choice below (tuning knob)
        const char* rule = getenv("EU_ITEM_RULE");
        // decomposed runs keep the fixed rule (the one measured at 2 and 8 GPUs) unless EU_ITEM_RULE=sim asks otherwise
        const bool decomposed = h->cfg.world_size > 1;
        const bool rule_fixed = (rule && std::strcmp(rule, "fixed") == 0) || (!rule && decomposed);
        const bool sim_short = rule && std::strcmp(rule, "simshort") == 0;      // short pieces (6..16) + 4 % guard
        const bool rule_sim = (rule && std::strcmp(rule, "sim") == 0) || sim_short;
        const char* e = getenv("EU_MARCH_LEN");
        if (e && atoi(e) > 0) {
            lmax = std::min(atoi(e), 4096);
        } else if (rule_fixed) {
            while (lmax > 2 && (hi - lo)/lmax < 6*n_warps) lmax /= 2;
        } else if (rule_sim) {
            // L by simulating the round-robin hand-out exactly -- piece lengths in list order, load of a warp = sum over
            // its items of (length + head steps) -- and minimising the busiest warp's load (tools/item_balance.py is
            // the same model).  The rule of decomposed runs, restricted there to short pieces (6..16 steps): the fixed
            // rule leaves 8 % of the warps' time idle on 2, 4 and 8 ranks of the bench grid (8192 column slices over
            // 3552 warps), pieces of 10-14 steps 1-2 %; short pieces measured +5 % on a single-rank 32-plane slab, long
            // ones (63 steps) -10 % on 2 ranks.  EU_ITEM_RULE=sim applies it to any run with L = 4..48.
            std::vector<int> chain_n;
            std::vector<char> seen(cls.size(), 0);
            long long singles = 0;
            for (int s = lo; s < hi; ++s) {
                if (seen[size_t(s - lo)]) continue;
                const int id = cls[size_t(s - lo)];
                if (id == EU_ITEM_GENERIC) { ++singles; continue; }
                const int n = chain_length(s, id, seen);
                const int step = std::max(classes[size_t(id)].D/EU_SLICE, 1);
                for (int k = 0; k < n; ++k) seen[size_t(s + k*step - lo)] = 1;
                chain_n.push_back(n);
            }
            const char* eh = getenv("EU_ITEM_HEAD");
            // start of a march in steps: two pieces of 15 beat four of 8 by 5 % at equal balance on a 32-plane slab -> ~0.75
            const double head = eh ? atof(eh) : 0.75;
            const int L_lo = sim_short ? 6 : 4, L_hi = sim_short ? 16 : 48;
            double best = 1e300;
            // the warps the persistent grid really has (3 blocks of 8 per SM, 2 with the capillary term)
            const int n_warps = h->n_sms*eu_fast_warps_per_sm(h->par.method_capillary != 0);
            std::vector<double> load(size_t(n_warps), 0.0);
            // busiest warp's load for piece length L
            auto simulate = [&](int L) {
                // pieces of all chains, plane group by plane group (the order of the scan below): the k-th piece of
                // every chain, then the (k+1)-th
                std::fill(load.begin(), load.end(), 0.0);
                long long v = 0;
                for (long long q = 0; q < singles; ++q, ++v) load[size_t(v % n_warps)] += 1.0 + head;
                std::map<int, std::vector<int> > pieces_of;          // chain length -> piece lengths (few distinct lengths)
                size_t max_pieces = 0;
                for (int n : chain_n) {
                    std::vector<int>& pl = pieces_of[n];
                    if (pl.empty()) {
                        for (int left = n; left > 0;) {
                            const int pcs = (left + L - 1)/L;
                            const int len = (left + pcs - 1)/pcs;
                            pl.push_back(len);
                            left -= len;
                        }
                    }
                    max_pieces = std::max(max_pieces, pl.size());
                }
                for (size_t k = 0; k < max_pieces; ++k) {
                    for (int n : chain_n) {
                        const std::vector<int>& pl = pieces_of[n];
                        if (k < pl.size()) { load[size_t(v % n_warps)] += pl[k] + head; ++v; }
                    }
                }
                return *std::max_element(load.begin(), load.end());
            };
            for (int L = L_lo; L <= L_hi; ++L) {
                const double mx = simulate(L);
                if (mx < best) { best = mx; lmax = L; }
            }
            if (sim_short) {
                // Decomposed runs leave the fixed rule (the one measured at 2 and 8 GPUs) only where the model promises
                // at least 4 %: a shorter piece also means more march heads, whose real cost the model only estimates.
                int lfixed = 32;
                while (lfixed > 2 && (hi - lo)/lfixed < 6*(h->n_sms*32)) lfixed /= 2;
                if (best > 0.96*simulate(lfixed)) lmax = lfixed;
            }
        } else {
            // chain lengths (a dry run of the scan below with unlimited pieces)
            std::vector<int> chain_n;
            long long n_generic = 0;
            {
                std::vector<char> seen(cls.size(), 0);
                for (int s = lo; s < hi; ++s) {
                    if (seen[size_t(s - lo)]) continue;
                    const int id = cls[size_t(s - lo)];
                    if (id == EU_ITEM_GENERIC) { ++n_generic; continue; }
                    const int n = chain_length(s, id, seen);
                    const int step = std::max(classes[size_t(id)].D/EU_SLICE, 1);
                    for (int k = 0; k < n; ++k) seen[size_t(s + k*step - lo)] = 1;
                    chain_n.push_back(n);
                }
            }
            double best = 1e300;
            for (int L = 4; L <= 64; ++L) {
                long long pieces = 0, steps = 0;
                for (int n : chain_n) { pieces += (n + L - 1)/L; steps += n; }
                if (pieces == 0) break;
                const long long items_total = pieces + n_generic;
                const double per_warp = double((items_total + n_warps - 1)/n_warps);
                const double cost = per_warp*(double(steps)/double(pieces) + 0.35);
                if (cost <= best) { best = cost; lmax = L; }
            }
        }
    }
    std::vector<int2> items;
    std::vector<char> taken(cls.size(), 0);
    long long n_class_slices = 0;
    for (int s = lo; s < hi; ++s) {
        if (taken[size_t(s - lo)]) continue;
        const int id = cls[size_t(s - lo)];
        int len = 1;
        if (id != EU_ITEM_GENERIC) {
            const int step = classes[size_t(id)].D/EU_SLICE;
            if (step > 0) {
                // this piece: an equal share of what is left of the chain
                const int left = chain_length(s, id, taken);
                const int pieces = (left + lmax - 1)/lmax;
                len = (left + pieces - 1)/pieces;
                for (int k = 1; k < len; ++k) taken[size_t(s + k*step - lo)] = 1;
            }
            n_class_slices += len;
        }
        taken[size_t(s - lo)] = 1;
        items.push_back(make_int2(s, len | (id << 16)));
    }
    // row tiles: where 8*R consecutive items start at consecutive slices (R = slices per grid row, the second
    // largest slice-aligned offset of the class), reorder them as R groups of 8 rows so that the warps of a block
    // share their y-neighbour lines in L1
    {
        const char* e = getenv("EU_ROW_TILES");
        const bool tiles = !(e && atoi(e) == 0);
        size_t i = 0;
        std::vector<int2> tmp;
        while (tiles && i < items.size()) {
            const int id = int(unsigned(items[i].y) >> 16);
            int R = 0;
            if (id != EU_ITEM_GENERIC) {
                const EuSliceClass& c = classes[size_t(id)];
                for (int q = 0; q < 4; ++q)
                    if (c.fid_mul[q] && c.nb_off[q] > EU_SLICE && c.nb_off[q] % EU_SLICE == 0 && c.nb_off[q] != c.D)
                        R = (R == 0) ? c.nb_off[q]/EU_SLICE : std::min(R, c.nb_off[q]/EU_SLICE);
            }
            const size_t group = size_t(8)*size_t(R);
            bool uniform = R >= 2 && i + group <= items.size();
            for (size_t k = 1; uniform && k < group; ++k)
                uniform = items[i + k].x == items[i].x + int(k) && int(unsigned(items[i + k].y) >> 16) != EU_ITEM_GENERIC;
            if (!uniform) { ++i; continue; }
            tmp.assign(items.begin() + long(i), items.begin() + long(i + group));
            for (size_t p = 0; p < group; ++p) items[i + p] = tmp[(p & 7)*size_t(R) + (p >> 3)];
            i += group;
        }
    }
    h->n_items = int(items.size());
    h->n_classes = int(classes.size());
    h->plan_max_len = 0;
    {
        long long n_class_items = 0;
        for (const int2& it : items) {
            if (int(unsigned(it.y) >> 16) == EU_ITEM_GENERIC) continue;
            h->plan_max_len = std::max(h->plan_max_len, it.y & 0xffff);
            ++n_class_items;
        }
        h->plan_mean_len = n_class_items ? double(n_class_slices)/double(n_class_items) : 0.0;
    }
    h->class_fraction = hi > lo ? double(n_class_slices)/double(hi - lo) : 0.0;
    if (items.empty()) items.push_back(make_int2(0, 0));
    if (classes.empty()) { EuSliceClass c; std::memset(&c, 0, sizeof(c)); classes.push_back(c); }
    int rc;
    if ((rc = upload_vec(h, h->d_items, items))) return rc;
    if ((rc = upload_vec(h, h->d_classes, classes))) return rc;
    h->items_lo = lo;
    h->items_hi = hi;
    h->items_variant = variant_key;
    return EU_OK;
}

// one substep on the resident state; returns the number of kernels launched.  In FAST mode with several
// ranks the halo exchange is part of the kernel (`exchange`: this substep takes part in the lockstep).
int launch_substep(eu_handle h, const EuStepArgs& a, bool exchange)
{
    const EuGridDev g = h->grid();
    if (h->mode == EU_MODE_FAST && h->tensor_fast == 2) {
        eu_launch_fast_step_t3(g, h->tabf, h->fast(), a, h->own_lo/EU_SLICE, (h->own_hi + EU_SLICE - 1)/EU_SLICE, h->n_sms, h->st);
        return 1;
    }
    if (h->mode == EU_MODE_FAST) {
        const int slice_lo = h->own_lo/EU_SLICE;
        const int slice_hi = (h->own_hi + EU_SLICE - 1)/EU_SLICE;
        EuHaloDev halo;
        std::memset(&halo, 0, sizeof(halo));
        const bool fused = exchange && fused_halo(h);
        bool box = h->box && !h->use_nn && a.method_viscous;             // (the box kernel has the viscous term built in)
        int bnd_lo = 0, bnd_hi = 0, box_info[6] = { 0, 0, 0, 0, 0, 0 };
        if (box && fused) {
            // own planes that hold cells a neighbour rank keeps as ghosts, or that read ghosts
            const int D = h->axis[2];
            if (h->fused_down_max >= 0) bnd_lo = (h->fused_down_max - h->own_lo)/D + 1;
            if (h->fused_up_min < INT_MAX) bnd_hi = (h->own_hi - 1 - h->fused_up_min)/D + 1;
            if (eu_box_plan_units(h->box, bnd_lo, bnd_hi, a.method_capillary != 0, box_info)) box = false;      // slab too thin
        }
        if (fused) {
            const int out = h->cur ^ 1;
            halo.enabled = 1;
            halo.a_hi = h->fused_a_hi;
            halo.b_lo = h->fused_b_lo;
            halo.epoch = ++h->epoch;
            for (int r = 0; r < 2; ++r) {
                halo.dst[r] = h->d_fused_dst[r].p;
                if (h->fused_peer[r] >= 0) {
                    eu_solver::Peer* p = h->peers[size_t(h->fused_peer[r])];
                    halo.peer_S[r] = p->S[out];
                    halo.peer_pc[r] = a.method_capillary ? p->pc[out] : nullptr;
                    halo.counter[r] = p->counter.p;
                    halo.peer_flag[r] = p->flags + h->cfg.rank;
                } else {
                    halo.counter[r] = h->d_fused_dummy.p;
                    halo.peer_flag[r] = h->d_fused_dummy.p + 1;
                }
                halo.total[r] = h->fused_total[r];
            }
            halo.my_flags = h->d_comm_flags.p;
            halo.n_wait = 0;
            for (eu_solver::Peer* p : h->peers) if (p->recv && halo.n_wait < 2) halo.wait_rank[halo.n_wait++] = p->rank;
            halo.timeout_cycles = 20000000000LL;
            halo.err_flag = h->d_flags.p + 3;
        }
        if (box) {
            // box numbering: plane sweep over tiles with TMA-staged operands (eu_tile.cuh).  With the fused exchange the
            // own planes that hold cells a neighbour rank keeps as ghosts (or that read ghosts) are swept first, as
            // short work units of their own, and pushed; the kernel counts finished units instead of slices.
            if (fused) {
                const unsigned nA = unsigned(box_info[3]), nB = unsigned(box_info[4]);
                const bool same_peer = h->fused_peer[0] >= 0 && h->fused_peer[0] == h->fused_peer[1];
                halo.total[0] = same_peer ? nA + nB : (h->fused_peer[0] >= 0 ? nA : 0xffffffffu);
                halo.total[1] = same_peer ? nA + nB : (h->fused_peer[1] >= 0 ? nB : 0xffffffffu);
            }
            const int nl = eu_launch_box_step(h->box, g, h->tabf, h->fast(), a, halo, h->cur, slice_lo, slice_hi, bnd_lo, bnd_hi, h->st);
            h->ran_box = nl >= 0;
            return nl < 0 ? -1 : nl;
        }
        h->ran_box = false;
        // the items cover the slices that are not handled as slab-boundary ranges
        if (build_items(h, fused ? h->fused_a_hi : slice_lo, fused ? h->fused_b_lo : slice_hi) != EU_OK) return -1;
        eu_launch_fast_step(g, h->tabf, h->fast(), a, halo, slice_lo, slice_hi, h->n_sms, h->st);
        return 1;
    }
    int launches = 1;
    if (a.method_capillary) {
        eu_launch_strict_pc(g, h->tab, a.S_in, const_cast<double*>(a.pc_in), h->st);
        ++launches;
    }
    EuStrictDev s;
    s.list = h->d_strict_list.p;
    s.porevol = h->d_porevol.p;
    eu_launch_strict_step(g, h->tab, s, h->d_hf_flux.p, a, h->st);
    return launches;
}

// after a substep wrote buffer `out_buf`: push my boundary cells to the peers' ghosts, then make the
// stream wait for the peers' pushes of the same epoch
int halo_exchange(eu_handle h, int out_buf, bool with_pc, int* launches)
{
    if (h->cfg.world_size <= 1) return EU_OK;
    if (!h->comm_ready) return fail(h, EU_ERR_COMM, "world_size > 1 needs eu_comm_connect");
    if (fused_halo(h)) return EU_OK;           // done inside k_fast_step
    ++h->epoch;
    for (eu_solver::Peer* p : h->peers) {
        if (p->n_send == 0) continue;
        eu_launch_halo_push(p->src.p, p->dst.p, p->n_send, h->d_S[out_buf].p, p->S[out_buf],
                            with_pc ? h->d_pc[out_buf].p : nullptr, with_pc ? p->pc[out_buf] : nullptr,
                            p->counter.p, p->flags + h->cfg.rank, h->epoch, h->st);
        ++*launches;
    }
    if (h->n_wait > 0) {
        eu_launch_halo_wait(h->d_comm_flags.p, h->d_wait_ranks.p, h->n_wait, h->epoch, 20000000000LL, h->d_flags.p + 3, h->st);
        ++*launches;
    }
    return EU_OK;
}

int compute_cfl(eu_handle h, const double gravity[3], bool want_v, bool want_g, bool want_c, double out[3],
                int* zero_flag, int* launches)
{
    const EuGridDev g = h->grid();
    out[0] = out[1] = out[2] = 1e99;
    *zero_flag = 0;
    // The velocity pass also compacts the half-face fluxes for the FAST kernel, so it always runs.
    EU_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 4*sizeof(int), h->st));
    eu_launch_cfl_velocity_compact(g, h->fluid.cfl_factor[0], h->d_hf_flux.p, h->d_fid_of_hf.p,
                                   h->mode == EU_MODE_FAST ? reinterpret_cast<double*>(h->d_qg.p) : nullptr, h->d_qa.p, h->d_block_min.p, h->d_flags.p,
                                   h->d_scalars.p + 0, h->st);
    *launches += 2;
    const bool grav_cached = h->cfl_grav_valid && std::memcmp(h->cfl_grav_gravity, gravity, 3*sizeof(double)) == 0;
    if (want_g && !grav_cached) {
        eu_launch_cfl_gravity(g, h->tab, h->fluid.cfl_factor[1], gravity, h->d_block_min.p, h->d_scalars.p + 1, h->st);
        *launches += 2;
    }
    if (want_c && !h->cfl_cap_valid) {
        eu_launch_cfl_capillary(g, h->fluid.cfl_factor[2], h->d_block_min.p, h->d_scalars.p + 2, h->st);
        *launches += 2;
    }
    double host[3];
    int flags[4];
    EU_CUDA(h, cudaMemcpyAsync(host, h->d_scalars.p, 3*sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaMemcpyAsync(flags, h->d_flags.p, 4*sizeof(int), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    if (want_g && !grav_cached) {
        h->cfl_grav = host[1];
        h->cfl_grav_valid = true;
        std::memcpy(h->cfl_grav_gravity, gravity, 3*sizeof(double));
    }
    if (want_c && !h->cfl_cap_valid) { h->cfl_cap = host[2]; h->cfl_cap_valid = true; }
    double vals[4] = { host[0], want_g ? h->cfl_grav : 1e100, want_c ? h->cfl_cap : 1e100, 0.0 };
    if (h->cfg.world_size > 1) {
        if (!h->allreduce) return fail(h, EU_ERR_COMM, "world_size > 1 needs eu_comm_set_allreduce");
        h->allreduce(h->allreduce_user, vals, 3, 0);
        double z = flags[0] ? 1.0 : 0.0;
        h->allreduce(h->allreduce_user, &z, 1, 1);
        flags[0] = z != 0.0;
    }
    if (want_v) out[0] = vals[0];
    if (want_g) out[1] = vals[1];
    if (want_c) out[2] = vals[2];
    *zero_flag = flags[0];
    return EU_OK;
}

} // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int eu_abi_version(void) { return EU_ABI_VERSION; }

void eu_default_params(eu_params* p)
{
    p->courant_number = 0.5;
    p->method_viscous = p->method_gravity = p->method_capillary = 1;
    p->use_cfl_viscous = p->use_cfl_gravity = p->use_cfl_capillary = 1;
    p->minimum_small_steps = 1;
    p->maximum_small_steps = 10000;
    p->check_sat = 1;
    p->clamp_sat = 0;
}

const char* eu_last_error(eu_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int eu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int eu_create(const eu_config* cfg, eu_handle* out)
{
    if (!cfg || !out) return fail(nullptr, EU_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != EU_ABI_VERSION) return fail(nullptr, EU_ERR_ARG, "ABI version mismatch");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, EU_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, EU_ERR_ARG, "bad device ordinal");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, EU_ERR_CUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) return fail(nullptr, EU_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, EU_ERR_CUDA, "device is not sm_100-class: the kernels are built for sm_100a only");
    eu_solver* h = new eu_solver;
    h->cfg = *cfg;
    if (h->cfg.world_size < 1) h->cfg.world_size = 1;
    eu_default_params(&h->par);
    h->mode = EU_MODE_STRICT;
    h->n_sms = prop.multiProcessorCount;
    { const char* e = getenv("EU_PREFETCH"); if (e) h->prefetch = std::min(std::max(atoi(e), 0), 8); }
    { const char* e = getenv("EU_L2_HINT"); if (e) h->l2_hint = atoi(e) != 0; }
    { const char* e = getenv("EU_PIN_CACHE"); if (e) h->pin_cache = atoi(e) != 0; }
    { const char* e = getenv("EU_BOX"); if (e) h->box_enabled = atoi(e) != 0; }
    std::memset(&h->fluid, 0, sizeof(h->fluid));
    std::memset(&h->tab, 0, sizeof(h->tab));
    if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess) {
        delete h;
        return fail(nullptr, EU_ERR_CUDA, std::string("stream/event creation failed: ") + cudaGetErrorString(e));
    }
    *out = h;
    return EU_OK;
}

void eu_destroy(eu_handle h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->st);
    unpin_all(h);
    eu_box_plan_destroy(h->box);
    h->box = nullptr;
    for (eu_solver::Peer* p : h->peers) {
        for (void* o : p->opened) if (o) cudaIpcCloseMemHandle(o);
        delete p;
    }
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->st);
    delete h;
}

int eu_set_params(eu_handle h, const eu_params* p)
{
    if (!h || !p) return EU_ERR_ARG;
    h->par = *p;
    return EU_OK;
}

int eu_grid_begin(eu_handle h, int n_cells_global, int n_local_cells, long long n_local_halffaces)
{
    if (!h) return EU_ERR_ARG;
    if (n_cells_global <= 0 || n_local_cells <= 0 || n_local_halffaces <= 0 || n_local_halffaces > INT_MAX)
        return fail(h, EU_ERR_ARG, "bad grid sizes (half-faces per rank must fit 32-bit)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    h->grid_open = true; h->grid_ready = false; h->state_ready = false; h->contracted = false;
    eu_box_plan_destroy(h->box);
    h->box = nullptr;
    h->d_qa.release();
    h->d_Ga.release();
    h->cfl_cap_valid = h->cfl_grav_valid = false;
    h->n_global = n_cells_global; h->n_local_expected = n_local_cells; h->H_expected = n_local_halffaces;
    h->n_local = 0; h->H = 0;
    h->ranges.clear();
    h->h_hf_offset.assign(1, 0);
    h->h_hf_offset.reserve(size_t(n_local_cells) + 1);
    h->b_hf.clear(); h->b_kind.clear(); h->b_pcell.clear(); h->b_pface.clear(); h->b_sat.clear();
    h->any_rock_ids = false;
    h->d_nn.release();
    const size_t n = size_t(n_local_cells), H = size_t(n_local_halffaces);
    EU_CUDA(h, h->d_hf_nbr.alloc(H));
    EU_CUDA(h, h->d_hf_area.alloc(H));
    EU_CUDA(h, h->d_hf_normal.alloc(3*H));
    EU_CUDA(h, h->d_hf_centroid.alloc(3*H));
    EU_CUDA(h, h->d_cell_volume.alloc(n));
    EU_CUDA(h, h->d_cell_centroid.alloc(3*n));
    EU_CUDA(h, h->d_poro.alloc(n));
    EU_CUDA(h, h->d_perm.alloc(9*n));
    EU_CUDA(h, h->d_rock.alloc(n));
    EU_CUDA(h, cudaMemsetAsync(h->d_rock.p, 0, n*sizeof(int), h->st));
    return EU_OK;
}

int eu_grid_append(eu_handle h, const eu_grid_chunk* c)
{
    if (!h || !c) return EU_ERR_ARG;
    if (!h->grid_open) return fail(h, EU_ERR_ARG, "eu_grid_append before eu_grid_begin");
    if (c->n_cells <= 0) return EU_OK;
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    if (!h->ranges.empty()) {
        const Range& last = h->ranges.back();
        if (c->first_cell < last.first + last.count) return fail(h, EU_ERR_ARG, "chunks must come in ascending cell order");
    }
    if (c->first_cell < 0 || c->first_cell + c->n_cells > h->n_global) return fail(h, EU_ERR_ARG, "chunk outside the grid");
    if (h->n_local + c->n_cells > h->n_local_expected) return fail(h, EU_ERR_ARG, "more cells than announced");
    long long nhf = 0;
    for (int i = 0; i < c->n_cells; ++i) {
        if (c->hf_count[i] < 0) return fail(h, EU_ERR_ARG, "negative half-face count");
        nhf += c->hf_count[i];
        h->h_hf_offset.push_back(int(h->H + nhf));
    }
    if (h->H + nhf > h->H_expected) return fail(h, EU_ERR_ARG, "more half-faces than announced");
    if (!h->ranges.empty() && h->ranges.back().first + h->ranges.back().count == c->first_cell) {
        h->ranges.back().count += c->n_cells;
    } else {
        Range r = { c->first_cell, c->n_cells, h->n_local };
        h->ranges.push_back(r);
    }
    // neighbour array with the boundary half-faces tagged -2-(boundary index)
    std::vector<int> nbr(c->hf_neighbour, c->hf_neighbour + nhf);
    for (int i = 0; i < c->n_bnd; ++i) {
        const int rel = c->bnd_hf[i];
        if (rel < 0 || rel >= nhf) return fail(h, EU_ERR_ARG, "boundary half-face index out of range");
        if (nbr[rel] >= 0) return fail(h, EU_ERR_ARG, "boundary condition given for an interior half-face");
        const int kind = c->bnd_kind[i];
        if (kind != EU_HF_DIRICHLET && kind != EU_HF_PERIODIC) return fail(h, EU_ERR_ARG, "bad boundary kind");
        nbr[rel] = -2 - int(h->b_hf.size());
        h->b_hf.push_back(int(h->H + rel));
        h->b_kind.push_back(kind);
        h->b_sat.push_back(c->bnd_sat ? c->bnd_sat[i] : 1.0);
        h->b_pcell.push_back(kind == EU_HF_PERIODIC ? c->bnd_partner_cell[i] : -1);
        h->b_pface.push_back(kind == EU_HF_PERIODIC ? c->bnd_partner_face[i] : -1);
    }
    for (long long k = 0; k < nhf; ++k) {
        if (nbr[k] == -1) return fail(h, EU_ERR_ARG, "boundary half-face without a saturation boundary condition");
    }
    const size_t o = size_t(h->H), oc = size_t(h->n_local);
    int rc;
    if ((rc = upload(h, h->d_hf_nbr, o, nbr.data(), size_t(nhf)))) return rc;
    EU_CUDA(h, cudaStreamSynchronize(h->st));      // nbr is a temporary
    if ((rc = upload(h, h->d_hf_area, o, c->hf_area, size_t(nhf)))) return rc;
    if ((rc = upload(h, h->d_hf_normal, 3*o, c->hf_normal, size_t(3*nhf)))) return rc;
    if ((rc = upload(h, h->d_hf_centroid, 3*o, c->hf_centroid, size_t(3*nhf)))) return rc;
    if ((rc = upload(h, h->d_cell_volume, oc, c->cell_volume, size_t(c->n_cells)))) return rc;
    if ((rc = upload(h, h->d_cell_centroid, 3*oc, c->cell_centroid, size_t(3*c->n_cells)))) return rc;
    if ((rc = upload(h, h->d_poro, oc, c->porosity, size_t(c->n_cells)))) return rc;
    if ((rc = upload(h, h->d_perm, 9*oc, c->permeability, size_t(9*c->n_cells)))) return rc;
    if (c->rock_id) {
        h->any_rock_ids = true;
        if ((rc = upload(h, h->d_rock, oc, c->rock_id, size_t(c->n_cells)))) return rc;
    }
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    h->H += nhf;
    h->n_local += c->n_cells;
    return EU_OK;
}

int eu_set_fluid(eu_handle h, const eu_fluid* f)
{
    if (!h || !f) return EU_ERR_ARG;
    // the mode-dependent structures (strict list / records, pc scale, axis check) are built in eu_grid_end from the
    // fluid set before it: changing the fluid of a finished grid needs eu_grid_begin again (initObj re-flattens)
    if (h->grid_ready) return fail(h, EU_ERR_ARG, "eu_set_fluid after eu_grid_end: upload the grid again (eu_grid_begin)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    if (f->mobility_kind != EU_MOB_SCALAR && f->mobility_kind != EU_MOB_DIAGONAL) return fail(h, EU_ERR_ARG, "bad mobility kind");
    if (f->n_rocks < 0 || f->n_rocks > EU_MAX_ROCKS) return fail(h, EU_ERR_UNSUPPORTED, "at most 16 rock types");
    h->fluid = *f;
    const int ncol = f->mobility_kind == EU_MOB_SCALAR ? 3 : 7;
    h->h_tab_offset.assign(1, 0);
    h->h_tab_s.clear();
    for (int k = 0; k < 7; ++k) h->h_tab_cols[k].clear();
    if (f->n_rocks > 0) {
        h->h_tab_offset.assign(f->table_offset, f->table_offset + f->n_rocks + 1);
        const int nn = h->h_tab_offset.back();
        for (int r = 0; r < f->n_rocks; ++r) {
            const int cnt = h->h_tab_offset[r + 1] - h->h_tab_offset[r];
            if (cnt < 2) return fail(h, EU_ERR_ARG, "a rock table needs at least two nodes");
            if (cnt > 250) return fail(h, EU_ERR_UNSUPPORTED, "at most 250 nodes per rock table");
        }
        h->h_tab_s.assign(f->table_s, f->table_s + nn);
        for (int k = 0; k < ncol; ++k) h->h_tab_cols[k].assign(f->table_cols[k], f->table_cols[k] + nn);
    }
    int rc;
    if ((rc = upload_vec(h, h->d_tab_offset, h->h_tab_offset))) return rc;
    if ((rc = upload_vec(h, h->d_tab_s, h->h_tab_s))) return rc;
    for (int k = 0; k < 7; ++k) if ((rc = upload_vec(h, h->d_tab_cols[k], h->h_tab_cols[k]))) return rc;
    // FAST tables: mobility = kr/viscosity and J (or pc) per interval in intercept/slope form, one curve set per rock
    // (scalar mobility: krw, kro, J) or per rock and axis (diagonal tensor mobility: kr??_w, kr??_o, pc)
    h->fast_tables_ok = true;
    h->n_buckets = 0;
    const bool tensor = f->mobility_kind == EU_MOB_DIAGONAL;
    const int reps = tensor ? 3 : 1;
    std::vector<int> off_fast(1, 0);
    if (f->n_rocks > 0 && f->n_rocks*reps > EU_MAX_TABLES) h->fast_tables_ok = false;
    if (f->n_rocks > 0 && h->fast_tables_ok) {
        const double inf = std::numeric_limits<double>::infinity();
        std::vector<double> coef, jcoef, xb;
        for (int r = 0; r < f->n_rocks; ++r) {
            const int b = h->h_tab_offset[r], e = h->h_tab_offset[r + 1];
            for (int ax = 0; ax < reps; ++ax) {
                const std::vector<double>* col[3] = { &h->h_tab_cols[tensor ? 1 + ax : 0], &h->h_tab_cols[tensor ? 4 + ax : 1],
                                                      &h->h_tab_cols[tensor ? 0 : 2] };
                const size_t o = xb.size();
                coef.resize(4*(o + size_t(e - b)), 0.0);
                jcoef.resize(2*(o + size_t(e - b)), 0.0);
                for (int i = b; i < e; ++i) xb.push_back(h->h_tab_s[i]);
                for (int i = b; i + 1 < e; ++i) {
                    const double x0 = h->h_tab_s[i], dx = h->h_tab_s[i + 1] - x0;
                    if (!(dx > 0.0)) return fail(h, EU_ERR_ARG, "rock table saturations must be strictly increasing");
                    const size_t k = o + size_t(i - b);
                    for (int p = 0; p < 2; ++p) {
                        const double y0 = (*col[p])[i]/f->viscosity[p], y1 = (*col[p])[i + 1]/f->viscosity[p];
                        const double slope = (y1 - y0)/dx;
                        coef[4*k + 2*p] = y0 - slope*x0;
                        coef[4*k + 2*p + 1] = slope;
                    }
                    const double j0 = (*col[2])[i], slope = ((*col[2])[i + 1] - j0)/dx;
                    jcoef[2*k] = j0 - slope*x0;
                    jcoef[2*k + 1] = slope;
                }
                xb.back() = inf;
                off_fast.push_back(int(xb.size()));
            }
        }
        // buckets over [0,1): smallest power of two such that no bucket holds two interior nodes
        int nb = 64;
        for (; nb <= 4096; nb *= 2) {
            bool ok = true;
            for (int r = 0; r < f->n_rocks && ok; ++r) {
                const int b = h->h_tab_offset[r], e = h->h_tab_offset[r + 1];
                int prev_bucket = -1;
                for (int i = b + 1; i + 1 < e; ++i) {                   // interior nodes
                    int k = int(std::floor(h->h_tab_s[i]*nb));
                    k = std::min(std::max(k, 0), nb - 1);
                    if (k == prev_bucket) { ok = false; break; }
                    prev_bucket = k;
                }
            }
            if (ok) break;
        }
        if (nb > 4096) {
            h->fast_tables_ok = false;          // nodes closer than 1/4096: FAST search not applicable
        } else {
            h->n_buckets = nb;
            const int sets = f->n_rocks*reps;
            std::vector<unsigned char> bucket(size_t(sets)*nb, 0);
            for (int r = 0; r < f->n_rocks; ++r) {
                const int b = h->h_tab_offset[r], e = h->h_tab_offset[r + 1];
                for (int ax = 0; ax < reps; ++ax)
                    for (int k = 1; k < nb; ++k)
                        bucket[size_t(r*reps + ax)*nb + k] = (unsigned char)table_index_host(&h->h_tab_s[b], e - b, double(k)/nb);
            }
            if ((rc = upload_vec(h, h->d_fbucket, bucket))) return rc;
            if ((rc = upload_vec(h, h->d_fcoef, coef))) return rc;
            if ((rc = upload_vec(h, h->d_fjcoef, jcoef))) return rc;
            if ((rc = upload_vec(h, h->d_fxb, xb))) return rc;
        }
    }
    if ((rc = upload_vec(h, h->d_tab_offset_fast, off_fast))) return rc;
    set_tables_struct(h);
    h->fluid_set = true;
    h->contracted = false;
    h->cfl_cap_valid = h->cfl_grav_valid = false;
    // arithmetic mode.  Tensor mobility: eu_grid_end picks the FAST variant once the grid is there (axis-aligned
    // normals: scalar formula per face axis; oblique normals: k_fast_step_t3); without rock tables the tensor class is
    // isotropic (RockAnisotropicRelperm / ..._impl.hpp:86-101: the same quadratic curve in every direction) and runs
    // the scalar FAST path.
    h->tensor_fast = 0;
    if (h->cfg.mode == EU_MODE_STRICT) h->mode = EU_MODE_STRICT;
    else if (h->fast_tables_ok) { h->mode = EU_MODE_FAST; h->tensor_fast = (tensor && f->n_rocks > 0) ? 1 : 0; }
    else if (h->cfg.mode == EU_MODE_FAST)
        return fail(h, EU_ERR_UNSUPPORTED, "FAST mode needs rock-table nodes at least 1/4096 apart and at most 48 curve sets");
    else h->mode = EU_MODE_STRICT;
    return EU_OK;
}

int eu_grid_end(eu_handle h)
{
    if (!h) return EU_ERR_ARG;
    NvtxRange nvtx_end("eu_grid_end (structure building)");
    if (!h->grid_open) return fail(h, EU_ERR_ARG, "eu_grid_end before eu_grid_begin");
    if (!h->fluid_set) return fail(h, EU_ERR_ARG, "eu_set_fluid must precede eu_grid_end");
    if (h->n_local != h->n_local_expected || h->H != h->H_expected) return fail(h, EU_ERR_ARG, "fewer cells/half-faces than announced");
    if (h->fluid.n_rocks > 1 && !h->any_rock_ids) return fail(h, EU_ERR_ARG, "several rocks but no rock ids");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    int rc;
    // own range in local numbering
    {
        const int ob = h->cfg.world_size > 1 ? h->cfg.own_begin : 0;
        const int oe = h->cfg.world_size > 1 ? h->cfg.own_end : h->n_global;
        h->own_lo = h->global_to_local(ob);
        const int last = h->global_to_local(oe - 1);
        if (h->own_lo < 0 || last < 0 || last - h->own_lo != oe - 1 - ob) return fail(h, EU_ERR_ARG, "own cell range not fully uploaded");
        h->own_hi = last + 1;
    }
    if ((rc = upload_vec(h, h->d_hf_offset, h->h_hf_offset))) return rc;
    {
        std::vector<int> rf, rcn, rl, l2g(size_t(h->n_local));
        for (const Range& r : h->ranges) {
            rf.push_back(r.first); rcn.push_back(r.count); rl.push_back(r.local);
            for (int i = 0; i < r.count; ++i) l2g[size_t(r.local) + i] = r.first + i;
        }
        if ((rc = upload_vec(h, h->d_range_first, rf))) return rc;
        if ((rc = upload_vec(h, h->d_range_count, rcn))) return rc;
        if ((rc = upload_vec(h, h->d_range_local, rl))) return rc;
        if ((rc = upload_vec(h, h->d_l2g, l2g))) return rc;
        eu_launch_translate_nbr(h->d_hf_nbr.p, h->H, h->d_range_first.p, h->d_range_count.p, h->d_range_local.p, int(rf.size()), h->st);
    }
    // boundary tables; periodic partners -> local cell / half-face
    {
        const size_t nb = h->b_hf.size();
        std::vector<int> phf(std::max<size_t>(nb, 1), -1), pcl(std::max<size_t>(nb, 1), -1), kind(std::max<size_t>(nb, 1), 0);
        std::vector<double> sat(std::max<size_t>(nb, 1), 0.0);
        for (size_t i = 0; i < nb; ++i) {
            kind[i] = h->b_kind[i];
            sat[i] = h->b_sat[i];
            if (h->b_kind[i] == EU_HF_PERIODIC) {
                const int l = h->global_to_local(h->b_pcell[i]);
                if (l >= 0) {
                    const int cnt = h->h_hf_offset[l + 1] - h->h_hf_offset[l];
                    if (h->b_pface[i] < 0 || h->b_pface[i] >= cnt) return fail(h, EU_ERR_ARG, "periodic partner face out of range");
                    pcl[i] = l;
                    phf[i] = h->h_hf_offset[l] + h->b_pface[i];
                }
            }
        }
        if ((rc = upload_vec(h, h->d_bnd_kind, kind))) return rc;
        if ((rc = upload_vec(h, h->d_bnd_sat, sat))) return rc;
        if ((rc = upload_vec(h, h->d_bnd_phf, phf))) return rc;
        if ((rc = upload_vec(h, h->d_bnd_pcell, pcl))) return rc;
    }
    const size_t n = size_t(h->n_local), H = size_t(h->H);
    EU_CUDA(h, h->d_flags.alloc(4));
    EU_CUDA(h, h->d_scalars.alloc(16));
    EU_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 4*sizeof(int), h->st));
    EU_CUDA(h, h->d_owner_hf.alloc(H));
    EU_CUDA(h, h->d_porevol.alloc(n));
    const EuGridDev g = h->grid();
    eu_launch_owner(g, h->d_owner_hf.p, h->d_flags.p, h->st);
    eu_launch_porevol(g, h->d_porevol.p, h->st);
    int flags[4] = { 0, 0, 0, 0 };
    EU_CUDA(h, cudaMemcpyAsync(flags, h->d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    if (flags[0] == 1) return fail(h, EU_ERR_ARG, "connectivity is not symmetric (half-face without a matching half-face in the neighbour)");
    if (flags[0] == 2) return fail(h, EU_ERR_ARG, "periodic partner cell of an own cell was not uploaded");
    if (flags[0] == 3) return fail(h, EU_ERR_ARG, "neighbour (ghost) cell of an own cell was not uploaded");

    if (h->tensor_fast) {
        // FAST with diagonal tensor mobility needs axis-aligned face normals
        int not_aligned = 0;
        EU_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 4*sizeof(int), h->st));
        eu_launch_axis_check(g, h->d_flags.p, h->st);
        EU_CUDA(h, cudaMemcpyAsync(&not_aligned, h->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaStreamSynchronize(h->st));
        EU_CUDA(h, cudaGetLastError());
        if (not_aligned) h->tensor_fast = 2;       // oblique normals: the three-component variant (k_fast_step_t3)
    }
    EU_CUDA(h, h->d_fid_of_hf.alloc(H));
    if (h->mode == EU_MODE_STRICT) {
        EU_CUDA(h, h->d_strict_list.alloc(H));
        eu_launch_strict_list(g, h->d_owner_hf.p, h->d_strict_list.p, h->st);
        EU_CUDA(h, cudaMemsetAsync(h->d_fid_of_hf.p, 0xff, H*sizeof(int), h->st));
    } else {
        h->n_slices = (h->n_local + EU_SLICE - 1)/EU_SLICE;
        DevBuf<int> d_width;
        DevBuf<unsigned char> d_slot;          // canonical slot of every half-face (setup only)
        EU_CUDA(h, d_width.alloc(size_t(h->n_slices)));
        EU_CUDA(h, d_slot.alloc(std::max<size_t>(H, 1)));
        eu_launch_canonical_slots(g, d_slot.p, d_width.p, h->st);
        std::vector<int> width(size_t(h->n_slices));
        EU_CUDA(h, cudaMemcpyAsync(width.data(), d_width.p, width.size()*sizeof(int), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaStreamSynchronize(h->st));
        std::vector<int> base(size_t(h->n_slices) + 1);
        long long rec_total = 0;
        int planes = 1;
        for (int s = 0; s < h->n_slices; ++s) {
            base[size_t(s)] = int(rec_total);
            rec_total += (long long)width[size_t(s)]*EU_SLICE;
            planes = std::max(planes, width[size_t(s)]);
            if (rec_total > INT_MAX) return fail(h, EU_ERR_UNSUPPORTED, "too many half-face records for one GPU (32-bit)");
        }
        base[size_t(h->n_slices)] = int(rec_total);
        // unique-face arrays: three axis planes + one plane per local face slot, indexed plane*n_local + owner cell
        planes += 3;
        h->max_slots = planes - 3;
        h->F = (long long)planes*h->n_local;
        if (h->F > INT_MAX) return fail(h, EU_ERR_UNSUPPORTED, "face planes exceed 32-bit indexing on one GPU");
        if ((rc = upload_vec(h, h->d_slice_base, base))) return rc;
        EU_CUDA(h, h->d_rec.alloc(size_t(rec_total)));
        EU_CUDA(h, h->d_desc.alloc(size_t(rec_total/EU_SLICE) + 1));
        EU_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 4*sizeof(int), h->st));
        // axis planes: the three most common positive neighbour offsets (candidates: the offsets of the first, the
        // middle and the last-but-one-plane own cell), each shared by at least a quarter of the cells
        {
            std::vector<int> cand;
            const int samples[3] = { h->own_lo, h->own_lo + (h->own_hi - h->own_lo)/2, h->own_lo + (h->own_hi - h->own_lo)/3 };
            for (int sc : samples) {
                const int b = h->h_hf_offset[size_t(sc)], e = h->h_hf_offset[size_t(sc) + 1];
                std::vector<int> nb(size_t(std::max(e - b, 1)));
                if (e > b) EU_CUDA(h, cudaMemcpy(nb.data(), h->d_hf_nbr.p + b, size_t(e - b)*sizeof(int), cudaMemcpyDeviceToHost));
                for (int k = 0; k < e - b; ++k) {
                    const int d = nb[size_t(k)] - sc;
                    if (nb[size_t(k)] > sc && cand.size() < 16 && std::find(cand.begin(), cand.end(), d) == cand.end()) cand.push_back(d);
                }
            }
            h->axis[0] = h->axis[1] = h->axis[2] = 0;
            if (!cand.empty()) {
                DevBuf<int> d_cand;
                DevBuf<unsigned long long> d_votes;
                if ((rc = upload_vec(h, d_cand, cand))) return rc;
                EU_CUDA(h, d_votes.alloc(16));
                EU_CUDA(h, cudaMemsetAsync(d_votes.p, 0, 16*sizeof(unsigned long long), h->st));
                eu_launch_offset_votes(g, d_cand.p, int(cand.size()), d_votes.p, h->st);
                unsigned long long votes[16];
                EU_CUDA(h, cudaMemcpyAsync(votes, d_votes.p, sizeof(votes), cudaMemcpyDeviceToHost, h->st));
                EU_CUDA(h, cudaStreamSynchronize(h->st));
                std::vector<int> order(cand.size());
                for (size_t i = 0; i < order.size(); ++i) order[i] = int(i);
                std::sort(order.begin(), order.end(), [&](int a, int b) { return votes[a] > votes[b]; });
                std::vector<int> top;
                for (size_t i = 0; i < order.size() && top.size() < 3; ++i)
                    if (votes[order[i]]*4ULL >= (unsigned long long)h->n_local) top.push_back(cand[size_t(order[i])]);
                std::sort(top.begin(), top.end());
                // box numbering c = x + ax1*(y + ...): each axis offset must be a multiple of the one below
                bool nested = true;
                for (size_t i = 1; i < top.size(); ++i) nested = nested && top[i] % top[i - 1] == 0;
                if (!top.empty() && top[0] == 1 && nested)
                    for (size_t i = 0; i < top.size(); ++i) h->axis[i] = top[i];
            }
        }
        // box numbering: local cell c = x + nx*(y + ny*z) over whole planes, own cells = whole planes (z-slab ranks with
        // ghost planes, or a single rank); what the box kernel (eu_tile.cuh) needs
        h->box_ok = h->axis[2] > 0 && h->n_local % h->axis[2] == 0 && h->own_lo % h->axis[2] == 0 && h->own_hi % h->axis[2] == 0;
        eu_launch_assign_fid(g, h->d_owner_hf.p, d_slot.p, h->axis, h->box_ok ? 1 : 0, h->d_fid_of_hf.p, h->st);
        eu_launch_build_records(g, h->d_owner_hf.p, h->d_fid_of_hf.p, d_slot.p, h->d_slice_base.p, h->d_rec.p, h->d_desc.p,
                                h->d_flags.p + 1, h->st);
        int nreg[4] = { 0, 0, 0, 0 };
        EU_CUDA(h, cudaMemcpyAsync(nreg, h->d_flags.p, sizeof(nreg), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaStreamSynchronize(h->st));
        h->regular_fraction = rec_total > 0 ? double(nreg[1])/double(rec_total/EU_SLICE) : 0.0;
        h->h_slice_base = base;
        h->items_lo = h->items_hi = -1;
        // one extra face (id F) with zero flux, G and T: the dummy face of the explicit slots of a slice class
        const size_t F = size_t(std::max<long long>(h->F, 1)) + 1;
        EU_CUDA(h, h->d_qg.alloc(F));
        EU_CUDA(h, h->d_T.alloc(F));
        EU_CUDA(h, cudaMemsetAsync(h->d_qg.p, 0, F*sizeof(double2), h->st));
        EU_CUDA(h, cudaMemsetAsync(h->d_T.p, 0, F*sizeof(double), h->st));
        if (h->tensor_fast == 2) {
            EU_CUDA(h, h->d_fv.alloc(9*F));
            EU_CUDA(h, cudaMemsetAsync(h->d_fv.p, 0, 9*F*sizeof(double), h->st));
        }
        if (h->tensor_fast == 1) {
            EU_CUDA(h, h->d_axis8.alloc(F));
            EU_CUDA(h, cudaMemsetAsync(h->d_axis8.p, 0, F, h->st));
            eu_launch_face_axis(g, h->d_owner_hf.p, h->d_fid_of_hf.p, h->d_axis8.p, h->st);
        }
        EU_CUDA(h, h->d_pcscale.alloc(n));
        EU_CUDA(h, h->d_rock8.alloc(n));
        EU_CUDA(h, h->d_inv_porevol.alloc(n));
        for (int k = 0; k < 2 && eu_fast_uses_stored_lam(); ++k) {
            EU_CUDA(h, h->d_lam[k].alloc(n));
            EU_CUDA(h, cudaMemsetAsync(h->d_lam[k].p, 0, n*sizeof(double2), h->st));
        }
        eu_launch_pcscale(g, h->tab, h->d_pcscale.p, h->d_rock8.p, h->d_inv_porevol.p, h->st);
        EU_CUDA(h, cudaStreamSynchronize(h->st));
    }
    // state
    for (int k = 0; k < 2; ++k) {
        EU_CUDA(h, h->d_S[k].alloc(n));
        EU_CUDA(h, h->d_pc[k].alloc(n));
        EU_CUDA(h, cudaMemsetAsync(h->d_S[k].p, 0, n*sizeof(double), h->st));
        EU_CUDA(h, cudaMemsetAsync(h->d_pc[k].p, 0, n*sizeof(double), h->st));
    }
    EU_CUDA(h, h->d_S_init.alloc(n));
    EU_CUDA(h, h->d_hf_flux.alloc(H));
    EU_CUDA(h, h->d_residual.alloc(n));
    EU_CUDA(h, h->d_block_min.alloc(size_t(eu_cfl_blocks(h->n_local))));
    EU_CUDA(h, h->d_fail_key.alloc(1));
    if (h->mode == EU_MODE_FAST && !h->tensor_fast && h->box_ok && h->box_enabled && h->max_slots <= 16) {
        // box kernel: per-cell mask of the faces outside the axis planes, tensor maps over the state and face arrays
        EU_CUDA(h, h->d_cmask.alloc(n));
        EU_CUDA(h, h->d_irr_cells.alloc(n));
        EU_CUDA(h, h->d_acc_irr.alloc(n));
        EU_CUDA(h, cudaMemsetAsync(h->d_acc_irr.p, 0, n*sizeof(double), h->st));
        EU_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 4*sizeof(int), h->st));
        eu_launch_cell_mask(h->grid(), h->d_slice_base.p, h->d_rec.p, h->d_cmask.p, h->d_irr_cells.p, h->d_flags.p, h->st);
        int n_irr = 0;
        EU_CUDA(h, cudaMemcpyAsync(&n_irr, h->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaStreamSynchronize(h->st));
        const int nx = h->axis[1], ny = h->axis[2]/h->axis[1], nz = h->n_local/h->axis[2];
        EU_CUDA(h, h->d_qa.alloc(3*n));
        EU_CUDA(h, h->d_Ga.alloc(3*n));
        EU_CUDA(h, cudaMemsetAsync(h->d_qa.p, 0, 3*n*sizeof(double), h->st));
        EU_CUDA(h, cudaMemsetAsync(h->d_Ga.p, 0, 3*n*sizeof(double), h->st));
        h->box = eu_box_plan_create(nx, ny, nz, h->own_lo/h->axis[2], h->own_hi/h->axis[2], h->d_S[0].p, h->d_S[1].p, h->d_pc[0].p,
                                    h->d_pc[1].p, h->d_qa.p, h->d_Ga.p, h->d_T.p, h->d_cmask.p, h->d_irr_cells.p, n_irr, h->d_acc_irr.p, h->d_inv_porevol.p, h->n_sms);
    }
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    h->grid_open = false;
    h->grid_ready = true;
    h->cur = 0;
    return EU_OK;
}

int eu_local_cells(eu_handle h) { return h ? h->n_local : 0; }
int eu_resolved_mode(eu_handle h) { return (h && h->grid_ready) ? h->mode : EU_MODE_AUTO; }
double eu_regular_fraction(eu_handle h) { return h ? h->regular_fraction : 0.0; }
long long eu_local_halffaces(eu_handle h) { return h ? h->H : 0; }
int eu_work_plan(eu_handle h, double out[5])
{
    if (!h || !out) return EU_ERR_ARG;
    out[4] = 0.0;
    if (h->mode == EU_MODE_FAST && h->ran_box && h->box) {
        // box kernel: every own cell is swept by a tile; items = work units, march = planes per unit
        int info[6];
        eu_box_plan_info(h->box, info);
        out[0] = 1.0;
        out[1] = double(info[2]);
        const int planes = (h->own_hi - h->own_lo)/h->axis[2];
        const int tiles = ((h->axis[1] + info[0] - 1)/info[0])*((h->axis[2]/h->axis[1] + info[1] - 1)/info[1]);
        out[3] = info[2] > 0 ? double(planes)*tiles/double(info[2]) : 0.0;
        out[2] = std::ceil(out[3]);
        out[4] = 1.0;
        return EU_OK;
    }
    const bool have = h->mode == EU_MODE_FAST && h->items_lo >= 0;
    out[0] = have ? h->class_fraction : 0.0;
    out[1] = have ? double(h->n_items) : 0.0;
    out[2] = have ? double(h->plan_max_len) : 0.0;
    out[3] = have ? h->plan_mean_len : 0.0;
    return EU_OK;
}

int eu_upload_saturation(eu_handle h, const double* saturation)
{
    if (!h || !saturation) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    pin_host(h, saturation, size_t(h->n_local)*sizeof(double));
    EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur].p, saturation, size_t(h->n_local)*sizeof(double), cudaMemcpyHostToDevice, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    return EU_OK;
}

int eu_upload_state(eu_handle h, const double* saturation, const double* hf_flux)
{
    if (!h || !hf_flux) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    if (!saturation && h->cur != 0) {
        // the resident state stays (NULL saturation): after an odd number of substeps it lives in buffer 1
        EU_CUDA(h, cudaMemcpyAsync(h->d_S[0].p, h->d_S[h->cur].p, size_t(h->n_local)*sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    }
    h->cur = 0;                                  // every rank keeps the same buffer parity
    pin_host(h, hf_flux, size_t(h->H)*sizeof(double));
    if (saturation) pin_host(h, saturation, size_t(h->n_local)*sizeof(double));
    EU_CUDA(h, cudaMemcpyAsync(h->d_hf_flux.p, hf_flux, size_t(h->H)*sizeof(double), cudaMemcpyHostToDevice, h->st));
    if (saturation)
        EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur].p, saturation, size_t(h->n_local)*sizeof(double), cudaMemcpyHostToDevice, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    h->state_ready = true;
    return EU_OK;
}

int eu_download_saturation(eu_handle h, double* saturation)
{
    if (!h || !saturation) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    pin_host(h, saturation, size_t(h->n_local)*sizeof(double));
    EU_CUDA(h, cudaMemcpyAsync(saturation, h->d_S[h->cur].p, size_t(h->n_local)*sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    return EU_OK;
}

int eu_cfl_times(eu_handle h, const double gravity[3], double out[3])
{
    if (!h || !gravity || !out) return EU_ERR_ARG;
    if (!h->state_ready) return fail(h, EU_ERR_ARG, "no resident state (eu_upload_state)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    int zero = 0, launches = 0;
    int rc = compute_cfl(h, gravity, true, true, true, out, &zero, &launches);
    if (rc) return rc;
    if (zero) return fail(h, EU_ERR_CFL_ZERO, "Cfl computation gave dt = 0.0");
    return EU_OK;
}

int eu_small_step(eu_handle h, double dt, const double gravity[3], int n_src, const int* src_cell, const double* src_rate,
                  double* residual_out, int* bad_cell, double* bad_value)
{
    if (!h || !gravity) return EU_ERR_ARG;
    if (!h->state_ready) return fail(h, EU_ERR_ARG, "no resident state (eu_upload_state)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    int rc, nls = 0;
    if ((rc = upload_sources(h, n_src, src_cell, src_rate, &nls))) return rc;
    int zero = 0, launches = 0;
    double cfl[3];
    if ((rc = compute_cfl(h, gravity, false, false, false, cfl, &zero, &launches))) return rc;   // compacts the fluxes
    if ((rc = ensure_contracted(h, gravity))) return rc;
    const unsigned long long none = ~0ULL;
    EU_CUDA(h, cudaMemcpyAsync(h->d_fail_key.p, &none, sizeof(none), cudaMemcpyHostToDevice, h->st));
    if (h->mode == EU_MODE_FAST)
        eu_launch_fast_state(h->grid(), h->tabf, h->fast(), h->d_S[h->cur].p, h->par.method_capillary ? h->d_pc[h->cur].p : nullptr,
                             h->d_lam[h->cur].p, 0, h->n_local, h->st);
    EuStepArgs a = step_args(h, dt, gravity, nls, 0);
    a.residual_out = h->d_residual.p;
    // ghost entries of the new state keep the old values (single substep, no exchange)
    EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur ^ 1].p, h->d_S[h->cur].p, size_t(h->n_local)*sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    if (h->cfg.world_size > 1) return fail(h, EU_ERR_UNSUPPORTED, "eu_small_step is a single-rank debugging entry point");
    if (launch_substep(h, a, false) < 0) return EU_ERR_CUDA;
    h->cur ^= 1;
    unsigned long long key = none;
    EU_CUDA(h, cudaMemcpyAsync(&key, h->d_fail_key.p, sizeof(key), cudaMemcpyDeviceToHost, h->st));
    if (residual_out)
        EU_CUDA(h, cudaMemcpyAsync(residual_out, h->d_residual.p + h->own_lo, size_t(h->own_hi - h->own_lo)*sizeof(double),
                                   cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    if (bad_cell) *bad_cell = -1;
    if (key != none) {
        const int lc = int(key & 0xffffffffu);
        double v = 0.0;
        EU_CUDA(h, cudaMemcpy(&v, h->d_S[h->cur].p + lc, sizeof(double), cudaMemcpyDeviceToHost));
        if (bad_cell) *bad_cell = h->local_to_global(lc);
        if (bad_value) *bad_value = v;
        return EU_ERR_SAT_RANGE;
    }
    return EU_OK;
}

// EulerUpstreamResidual::computeResidual as an operator (Residual_impl.hpp:472-505).  Runs the substep kernel of
// the current mode with dt = 0 and the range check off on a scratch copy of the state: the residual it writes is
// exactly what smallTimeStep would have used; the resident saturation and the solver parameters are untouched.
int eu_compute_residual(eu_handle h, const double* saturation, const double gravity[3], const double* hf_flux,
                        int n_src, const int* src_cell, const double* src_rate,
                        int method_viscous, int method_gravity, int method_capillary, double* sat_delta)
{
    if (!h || !saturation || !gravity || !sat_delta) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    if (!hf_flux && !h->state_ready) return fail(h, EU_ERR_ARG, "no resident fluxes: pass hf_flux or call eu_upload_state first");
    if (h->cfg.world_size > 1) return fail(h, EU_ERR_UNSUPPORTED, "eu_compute_residual is a single-rank entry point");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t nbytes = size_t(h->n_local)*sizeof(double);
    const int in = h->cur ^ 1;                   // the buffer that does not hold the resident state
    pin_host(h, saturation, nbytes);
    if (hf_flux) pin_host(h, hf_flux, size_t(h->H)*sizeof(double));
    // The resident state is parked in d_S_init while both ping-pong buffers serve as scratch.
    EU_CUDA(h, cudaMemcpyAsync(h->d_S_init.p, h->d_S[h->cur].p, nbytes, cudaMemcpyDeviceToDevice, h->st));
    EU_CUDA(h, cudaMemcpyAsync(h->d_S[in].p, saturation, nbytes, cudaMemcpyHostToDevice, h->st));
    if (hf_flux) {
        EU_CUDA(h, cudaMemcpyAsync(h->d_hf_flux.p, hf_flux, size_t(h->H)*sizeof(double), cudaMemcpyHostToDevice, h->st));
        h->state_ready = true;
    }
    int rc, nls = 0, zero = 0, launches = 0;
    if ((rc = upload_sources(h, n_src, src_cell, src_rate, &nls))) return rc;
    const eu_params saved = h->par;
    h->par.method_viscous = method_viscous != 0;
    h->par.method_gravity = method_gravity != 0;
    h->par.method_capillary = method_capillary != 0;
    h->par.check_sat = 0;
    h->par.clamp_sat = 0;
    double cfl[3];
    rc = compute_cfl(h, gravity, false, false, false, cfl, &zero, &launches);     // compacts the fluxes (FAST)
    if (!rc) rc = ensure_contracted(h, gravity);
    if (!rc) {
        const unsigned long long none = ~0ULL;
        cudaMemcpyAsync(h->d_fail_key.p, &none, sizeof(none), cudaMemcpyHostToDevice, h->st);
        h->cur = in;
        if (h->mode == EU_MODE_FAST)
            eu_launch_fast_state(h->grid(), h->tabf, h->fast(), h->d_S[in].p, method_capillary ? h->d_pc[in].p : nullptr,
                                 h->d_lam[in].p, 0, h->n_local, h->st);
        EuStepArgs a = step_args(h, 0.0, gravity, nls, 0);
        a.residual_out = h->d_residual.p;
        if (launch_substep(h, a, false) < 0) rc = EU_ERR_CUDA;
        h->cur = in ^ 1;
    }
    h->par = saved;
    if (rc) return rc;
    EU_CUDA(h, cudaMemcpyAsync(sat_delta, h->d_residual.p + h->own_lo, size_t(h->own_hi - h->own_lo)*sizeof(double),
                               cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur].p, h->d_S_init.p, nbytes, cudaMemcpyDeviceToDevice, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    return EU_OK;
}

// EulerUpstreamResidual::computeCapPressures (Residual_impl.hpp:459-467) / computeCapPressure
// (SimulatorUtilities.hpp:219-230): the bit-exact kernel in every mode.
int eu_compute_cap_pressures(eu_handle h, const double* saturation, double* cap_pressures)
{
    if (!h || !saturation || !cap_pressures) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t nbytes = size_t(h->n_local)*sizeof(double);
    const int scratch = h->cur ^ 1;
    EU_CUDA(h, cudaMemcpyAsync(h->d_S[scratch].p, saturation, nbytes, cudaMemcpyHostToDevice, h->st));
    eu_launch_strict_pc(h->grid(), h->tab, h->d_S[scratch].p, h->d_residual.p, h->st);
    EU_CUDA(h, cudaMemcpyAsync(cap_pressures, h->d_residual.p, nbytes, cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    return EU_OK;
}

// ---- diagnostics (eu_diag.cu) ------------------------------------------------------------------------------
namespace {
// scratch for the diagnostics: 3 doubles per own cell and field, allocated on first use
int diag_scratch(eu_handle h, int fields)
{
    const size_t need = size_t(fields)*3*size_t(std::max(1, h->own_hi - h->own_lo));
    if (h->d_diag.n < need) EU_CUDA(h, h->d_diag.alloc(need));
    return EU_OK;
}
}

int eu_cell_velocity(eu_handle h, double* cell_velocity)
{
    if (!h || !cell_velocity) return EU_ERR_ARG;
    if (!h->state_ready) return fail(h, EU_ERR_ARG, "no resident fluxes (eu_upload_state)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    int rc;
    if ((rc = diag_scratch(h, 3))) return rc;
    const size_t n3 = 3*size_t(h->own_hi - h->own_lo);
    eu_launch_cell_velocity(h->grid(), h->d_hf_flux.p, h->d_diag.p, h->st);
    EU_CUDA(h, cudaMemcpyAsync(cell_velocity, h->d_diag.p, n3*sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    return EU_OK;
}

int eu_phase_velocities(eu_handle h, const double* saturation, const double* cell_velocity,
                        double* water_velocity, double* oil_velocity)
{
    if (!h || !water_velocity || !oil_velocity) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    if (!saturation && !h->state_ready) return fail(h, EU_ERR_ARG, "no resident state (eu_upload_state)");
    if (!cell_velocity && !h->state_ready) return fail(h, EU_ERR_ARG, "no resident fluxes (eu_upload_state)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    int rc;
    if ((rc = diag_scratch(h, 3))) return rc;
    const size_t n3 = 3*size_t(h->own_hi - h->own_lo);
    double* cv = h->d_diag.p;
    double* vw = cv + n3;
    double* vo = vw + n3;
    const double* S = h->d_S[h->cur].p;                   // NULL saturation: the resident state
    if (saturation) {
        EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur ^ 1].p, saturation, size_t(h->n_local)*sizeof(double), cudaMemcpyHostToDevice, h->st));
        S = h->d_S[h->cur ^ 1].p;
    }
    if (cell_velocity) EU_CUDA(h, cudaMemcpyAsync(cv, cell_velocity, n3*sizeof(double), cudaMemcpyHostToDevice, h->st));
    else eu_launch_cell_velocity(h->grid(), h->d_hf_flux.p, cv, h->st);       // NULL: from the resident fluxes
    eu_launch_phase_velocities(h->grid(), h->tab, S, cv, vw, vo, h->st);
    EU_CUDA(h, cudaMemcpyAsync(water_velocity, vw, n3*sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaMemcpyAsync(oil_velocity, vo, n3*sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    return EU_OK;
}

int eu_fractional_flow(eu_handle h, const double* saturation, double* frac_flow)
{
    if (!h || !frac_flow) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "grid not ready");
    if (!saturation && !h->state_ready) return fail(h, EU_ERR_ARG, "no resident state (eu_upload_state)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    const double* S = h->d_S[h->cur].p;
    if (saturation) {
        EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur ^ 1].p, saturation, size_t(h->n_local)*sizeof(double), cudaMemcpyHostToDevice, h->st));
        S = h->d_S[h->cur ^ 1].p;
    }
    eu_launch_fractional_flow(h->grid(), h->tab, S, h->d_residual.p, h->st);
    EU_CUDA(h, cudaMemcpyAsync(frac_flow, h->d_residual.p, size_t(h->own_hi - h->own_lo)*sizeof(double), cudaMemcpyDeviceToHost, h->st));
    EU_CUDA(h, cudaStreamSynchronize(h->st));
    EU_CUDA(h, cudaGetLastError());
    return EU_OK;
}

int eu_transport_solve_resident(eu_handle h, double time, const double gravity[3], int n_src, const int* src_cell,
                                const double* src_rate, eu_report* rep)
{
    if (!h || !gravity || !rep) return EU_ERR_ARG;
    if (!h->state_ready) return fail(h, EU_ERR_ARG, "no resident state (eu_upload_state)");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    std::memset(rep, 0, sizeof(*rep));
    rep->bad_cell = -1;
    const eu_params& p = h->par;
    int rc, nls = 0, launches = 0;
    if ((rc = upload_sources(h, n_src, src_cell, src_rate, &nls))) return rc;

    // saturation_initial (:181) and the second ping-pong buffer.  These copies must be complete on every rank
    // before any neighbour starts pushing ghosts; the reductions inside compute_cfl are that barrier.
    const size_t nbytes = size_t(h->n_local)*sizeof(double);
    EU_CUDA(h, cudaMemcpyAsync(h->d_S_init.p, h->d_S[h->cur].p, nbytes, cudaMemcpyDeviceToDevice, h->st));
    EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur ^ 1].p, h->d_S[h->cur].p, nbytes, cudaMemcpyDeviceToDevice, h->st));

    // ---- computeCflTime (EulerUpstream_impl.hpp:263-331)
    int zero = 0;
    double cfl[3];
    NvtxRange nvtx_solve("eu_transport_solve_resident");
    nvtxRangePushA("eu: computeCflTime + flux compaction");
    rc = compute_cfl(h, gravity, p.method_viscous && p.use_cfl_viscous, p.method_gravity && p.use_cfl_gravity,
                     p.method_capillary && p.use_cfl_capillary, cfl, &zero, &launches);
    nvtxRangePop();
    if (rc) return rc;
    rep->cfl_dt[0] = cfl[0]; rep->cfl_dt[1] = cfl[1]; rep->cfl_dt[2] = cfl[2];
    if (zero && p.method_viscous && p.use_cfl_viscous) {
        rep->status = EU_ERR_CFL_ZERO;
        rep->kernel_launches = launches;
        return fail(h, EU_ERR_CFL_ZERO, "Cfl computation gave dt = 0.0");
    }
    double cfl_dt = std::min(std::min(cfl[0], cfl[1]), cfl[2]);
    cfl_dt *= p.courant_number;

    // ---- number of small steps (:163-172)
    int nsteps;
    if (cfl_dt > time) {
        nsteps = p.minimum_small_steps;
    } else {
        double steps = std::min<double>(std::ceil(time/cfl_dt), std::numeric_limits<int>::max());
        nsteps = (steps != steps) ? std::numeric_limits<int>::min() : int(steps);
        nsteps = std::max(nsteps, p.minimum_small_steps);
        nsteps = std::min(nsteps, p.maximum_small_steps);
    }
    double dt = time/nsteps;

    if ((rc = ensure_contracted(h, gravity))) return rc;

    const unsigned long long none = ~0ULL;
    int repeats = 0;
    const int max_repeats = 10;
    bool finished = false;
    while (!finished) {
        ++rep->attempts;
        EU_CUDA(h, cudaMemcpyAsync(h->d_fail_key.p, &none, sizeof(none), cudaMemcpyHostToDevice, h->st));
        if (h->mode == EU_MODE_FAST) {
            eu_launch_fast_state(h->grid(), h->tabf, h->fast(), h->d_S[h->cur].p, p.method_capillary ? h->d_pc[h->cur].p : nullptr,
                                 h->d_lam[h->cur].p, 0, h->n_local, h->st);
            ++launches;
        }
        const int start = h->cur;
        if (h->mode == EU_MODE_FAST && !(h->box && !h->use_nn && p.method_viscous)) {       // work items of the slice-class kernel (built once per grid / decomposition)
            const bool fused = fused_halo(h);
            if ((rc = build_items(h, fused ? h->fused_a_hi : h->own_lo/EU_SLICE,
                                  fused ? h->fused_b_lo : (h->own_hi + EU_SLICE - 1)/EU_SLICE))) return rc;
        }
        NvtxRange nvtx_attempt("eu: substep loop (one attempt)");
        EU_CUDA(h, cudaEventRecord(h->ev0, h->st));
        for (int q = 0; q < nsteps; ++q) {
            EuStepArgs a = step_args(h, dt, gravity, nls, q);
            const int nl = launch_substep(h, a, true);
            if (nl < 0) return EU_ERR_CUDA;
            launches += nl;
            if ((rc = halo_exchange(h, h->cur ^ 1, h->mode == EU_MODE_FAST && p.method_capillary, &launches))) return rc;
            h->cur ^= 1;
        }
        EU_CUDA(h, cudaEventRecord(h->ev1, h->st));
        if (h->cfg.world_size > 1 && h->n_wait > 0) {
            // the neighbours' last pushes must have landed before anything else touches the ghosts
            eu_launch_halo_wait(h->d_comm_flags.p, h->d_wait_ranks.p, h->n_wait, h->epoch, 20000000000LL, h->d_flags.p + 3, h->st);
            ++launches;
        }
        unsigned long long key = none;
        int comm_err = 0;
        EU_CUDA(h, cudaMemcpyAsync(&key, h->d_fail_key.p, sizeof(key), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaMemcpyAsync(&comm_err, h->d_flags.p + 3, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaStreamSynchronize(h->st));
        EU_CUDA(h, cudaGetLastError());
        if (comm_err) return fail(h, EU_ERR_COMM, "halo exchange timed out waiting for a neighbour rank");
        rep->substeps_executed += nsteps;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        rep->device_ms = ms;
        if (h->cfg.world_size > 1) {
            double k = key == none ? -1.0 : double(key >> 32);   // first failing substep on any rank
            // encode as max over (large - substep) so that the earliest failure wins
            double enc = key == none ? 0.0 : 4294967296.0 - k;
            h->allreduce(h->allreduce_user, &enc, 1, 1);
            // the attempt failed in substep `first` somewhere; a rank whose own first failure came later (or never) has
            // no cell to report for it -- the reference stops at the lowest failing cell of that substep
            if (enc != 0.0) {
                const unsigned long long first = (unsigned long long)(4294967296.0 - enc);
                if (key == none || (key >> 32) > first) key = first << 32 | 0xffffffffu;
            }
        }
        if (key == none) {
            finished = true;
        } else {
            // "Saturation out of range in EulerUpstream: Cell <c>   sat <s>" (:344-346)
            const int lc = int(key & 0xffffffffu);
            const int fs = int(key >> 32);
            if (lc != -1 && unsigned(lc) != 0xffffffffu) {
                // buffer holding the output of substep fs: fs+1 swaps away from the attempt's start buffer
                const int buf = start ^ ((fs + 1) & 1);
                double v = 0.0;
                EU_CUDA(h, cudaMemcpy(&v, h->d_S[buf].p + lc, sizeof(double), cudaMemcpyDeviceToHost));
                rep->bad_cell = h->local_to_global(lc);
                rep->bad_value = v;
            }
            ++repeats;
            if (repeats > max_repeats) {
                rep->status = EU_ERR_SAT_RANGE;
                break;
            }
            nsteps *= 2;                       // :209-211
            dt = time/nsteps;
            EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur].p, h->d_S_init.p, nbytes, cudaMemcpyDeviceToDevice, h->st));
            EU_CUDA(h, cudaMemcpyAsync(h->d_S[h->cur ^ 1].p, h->d_S_init.p, nbytes, cudaMemcpyDeviceToDevice, h->st));
            if (h->cfg.world_size > 1) {
                // every rank must have restored its buffers before a neighbour pushes into them again
                EU_CUDA(h, cudaStreamSynchronize(h->st));
                double dummy = 0.0;
                h->allreduce(h->allreduce_user, &dummy, 1, 1);
            }
        }
    }
    rep->nsteps = nsteps;
    rep->dt = dt;
    rep->kernel_launches = launches;
    if (rep->status == EU_ERR_SAT_RANGE) {
        char buf[160];
        std::snprintf(buf, sizeof(buf), "Saturation out of range in EulerUpstream: Cell %d   sat %.17g", rep->bad_cell, rep->bad_value);
        return fail(h, EU_ERR_SAT_RANGE, buf);
    }
    if (finished) { rep->bad_cell = -1; rep->bad_value = 0.0; }
    return EU_OK;
}

int eu_transport_solve(eu_handle h, double* saturation, double time, const double gravity[3], const double* hf_flux,
                       int n_src, const int* src_cell, const double* src_rate, eu_report* report)
{
    if (!h || !saturation || !hf_flux || !report) return EU_ERR_ARG;
    int rc;
    NvtxRange nvtx_all("eu_transport_solve (H2D + solve + D2H)");
    if ((rc = eu_upload_state(h, saturation, hf_flux))) return rc;
    rc = eu_transport_solve_resident(h, time, gravity, n_src, src_cell, src_rate, report);
    if (rc != EU_OK && rc != EU_ERR_SAT_RANGE) return rc;
    int rc2 = eu_download_saturation(h, saturation);
    return rc != EU_OK ? rc : rc2;
}

void* eu_host_alloc(unsigned long long bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, size_t(bytes ? bytes : 1), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void eu_host_free(void* p) { if (p) cudaFreeHost(p); }
void eu_host_unpin_all(eu_handle h) { if (h) { cudaSetDevice(h->cfg.device); cudaStreamSynchronize(h->st); unpin_all(h); } }

struct EuBlobHeader {
    int magic, rank, world, own_begin, own_end, n_ghost;
    cudaIpcMemHandle_t mem[5];       // S[0], S[1], pc[0], pc[1], flags
    // ranks of ONE process (one solver per GPU, the C++ drop-in with several devices): CUDA IPC does not open a handle in
    // the process that exported it; peers there use the device pointers directly, with peer access enabled
    long long pid;
    int device, pad;
    void* raw[5];
};

int eu_comm_plan_sends(int own_begin, int own_end, int n_ghost, const int* ghost_global, const int* ghost_local,
                       int* send_global, int* send_peer_local)
{
    int n = 0;
    for (int i = 0; i < n_ghost; ++i) {
        if (ghost_global[i] >= own_begin && ghost_global[i] < own_end) {
            if (send_global) send_global[n] = ghost_global[i];
            if (send_peer_local) send_peer_local[n] = ghost_local[i];
            ++n;
        }
    }
    return n;
}

static void collect_ghosts(eu_handle h)
{
    h->ghost_global.clear();
    h->ghost_local.clear();
    for (int l = 0; l < h->n_local; ++l) {
        if (l >= h->own_lo && l < h->own_hi) { l = h->own_hi - 1; continue; }
        h->ghost_global.push_back(h->local_to_global(l));
        h->ghost_local.push_back(l);
    }
}

int eu_comm_blob_size(eu_handle h)
{
    if (!h || !h->grid_ready) return 0;
    collect_ghosts(h);
    return int(sizeof(EuBlobHeader) + 2*sizeof(int)*h->ghost_global.size());
}

int eu_comm_export(eu_handle h, void* blob)
{
    if (!h || !blob) return EU_ERR_ARG;
    if (!h->grid_ready) return fail(h, EU_ERR_ARG, "eu_comm_export before eu_grid_end");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    collect_ghosts(h);
    if (h->d_comm_flags.n == 0) {
        EU_CUDA(h, h->d_comm_flags.alloc(size_t(std::max(h->cfg.world_size, 1))));
        EU_CUDA(h, cudaMemset(h->d_comm_flags.p, 0, h->d_comm_flags.n*sizeof(unsigned)));
    }
    EuBlobHeader hd;
    std::memset(&hd, 0, sizeof(hd));
    hd.magic = 0x45553031; hd.rank = h->cfg.rank; hd.world = h->cfg.world_size;
    hd.own_begin = h->cfg.own_begin; hd.own_end = h->cfg.own_end; hd.n_ghost = int(h->ghost_global.size());
    void* ptrs[5] = { h->d_S[0].p, h->d_S[1].p, h->d_pc[0].p, h->d_pc[1].p, h->d_comm_flags.p };
    for (int k = 0; k < 5; ++k) EU_CUDA(h, cudaIpcGetMemHandle(&hd.mem[k], ptrs[k]));
    hd.pid = (long long)getpid();
    hd.device = h->cfg.device;
    for (int k = 0; k < 5; ++k) hd.raw[k] = ptrs[k];
    char* out = static_cast<char*>(blob);
    std::memcpy(out, &hd, sizeof(hd));
    std::memcpy(out + sizeof(hd), h->ghost_global.data(), sizeof(int)*h->ghost_global.size());
    std::memcpy(out + sizeof(hd) + sizeof(int)*h->ghost_global.size(), h->ghost_local.data(), sizeof(int)*h->ghost_local.size());
    return EU_OK;
}

int eu_comm_connect(eu_handle h, int n_blobs, const void* const* blobs, const int* blob_sizes)
{
    if (!h || !blobs || !blob_sizes) return EU_ERR_ARG;
    if (n_blobs != h->cfg.world_size) return fail(h, EU_ERR_ARG, "one blob per rank expected");
    if (h->d_comm_flags.n == 0) return fail(h, EU_ERR_ARG, "eu_comm_export must precede eu_comm_connect");
    EU_CUDA(h, cudaSetDevice(h->cfg.device));
    for (eu_solver::Peer* p : h->peers) {
        for (void* o : p->opened) if (o) cudaIpcCloseMemHandle(o);
        delete p;
    }
    h->peers.clear();
    std::vector<int> wait_ranks;
    for (int r = 0; r < n_blobs; ++r) {
        if (r == h->cfg.rank) continue;
        if (blob_sizes[r] < int(sizeof(EuBlobHeader))) return fail(h, EU_ERR_ARG, "short blob");
        EuBlobHeader hd;
        std::memcpy(&hd, blobs[r], sizeof(hd));
        if (hd.magic != 0x45553031 || hd.rank != r || hd.world != n_blobs) return fail(h, EU_ERR_ARG, "bad blob");
        if (blob_sizes[r] != int(sizeof(hd) + 2*sizeof(int)*size_t(hd.n_ghost))) return fail(h, EU_ERR_ARG, "blob size mismatch");
        const int* gg = reinterpret_cast<const int*>(static_cast<const char*>(blobs[r]) + sizeof(hd));
        const int* gl = gg + hd.n_ghost;
        std::vector<int> sg(static_cast<size_t>(std::max(hd.n_ghost, 1)), 0), sl(static_cast<size_t>(std::max(hd.n_ghost, 1)), 0);
        const int n_send = eu_comm_plan_sends(h->cfg.own_begin, h->cfg.own_end, hd.n_ghost, gg, gl, sg.data(), sl.data());
        bool recv = false;
        for (int g : h->ghost_global) if (g >= hd.own_begin && g < hd.own_end) { recv = true; break; }
        if (n_send == 0 && !recv) continue;
        eu_solver::Peer* p = new eu_solver::Peer;
        h->peers.push_back(p);
        p->rank = r; p->recv = recv; p->n_send = n_send;
        void* mapped[5];
        if (hd.pid == (long long)getpid()) {
            // same process: the peer's buffers are plain device pointers once peer access is on
            if (hd.device != h->cfg.device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, h->cfg.device, hd.device);
                if (!can) return fail(h, EU_ERR_COMM, "no peer-to-peer access between the devices of this process");
                cudaError_t e = cudaDeviceEnablePeerAccess(hd.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(h, EU_ERR_COMM, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            for (int k = 0; k < 5; ++k) mapped[k] = hd.raw[k];
        } else {
            for (int k = 0; k < 5; ++k) {
                cudaError_t e = cudaIpcOpenMemHandle(&p->opened[k], hd.mem[k], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess)
                    return fail(h, EU_ERR_COMM, std::string("cudaIpcOpenMemHandle (peer-to-peer access to the neighbour rank): ") + cudaGetErrorString(e));
                mapped[k] = p->opened[k];
            }
        }
        p->S[0] = static_cast<double*>(mapped[0]); p->S[1] = static_cast<double*>(mapped[1]);
        p->pc[0] = static_cast<double*>(mapped[2]); p->pc[1] = static_cast<double*>(mapped[3]);
        p->flags = static_cast<unsigned*>(mapped[4]);
        if (n_send > 0) {
            std::vector<int> src(static_cast<size_t>(n_send), 0);
            std::vector<int> dst(sl.begin(), sl.begin() + n_send);
            for (int i = 0; i < n_send; ++i) src[size_t(i)] = h->global_to_local(sg[size_t(i)]);
            int rc;
            if ((rc = upload_vec(h, p->src, src))) return rc;
            if ((rc = upload_vec(h, p->dst, dst))) return rc;
            EU_CUDA(h, p->counter.alloc(1));
            EU_CUDA(h, cudaMemset(p->counter.p, 0, sizeof(unsigned)));
        }
        if (recv) wait_ranks.push_back(r);
    }
    if (wait_ranks.size() > 32) return fail(h, EU_ERR_UNSUPPORTED, "more than 32 neighbour ranks");
    h->n_wait = int(wait_ranks.size());
    if (h->n_wait > 0) {
        int rc;
        if ((rc = upload_vec(h, h->d_wait_ranks, wait_ranks))) return rc;
    }
    h->epoch = 0;
    // ---- plan of the fused exchange (FAST mode): boundary slice ranges and per-cell ghost slots
    h->fused_ok = false;
    if (h->mode == EU_MODE_FAST) {
        int init[4] = { -1, INT_MAX, 0, 0 };
        DevBuf<int> d_adj;
        EU_CUDA(h, d_adj.alloc(4));
        EU_CUDA(h, cudaMemcpyAsync(d_adj.p, init, sizeof(init), cudaMemcpyHostToDevice, h->st));
        eu_launch_ghost_adjacent(h->grid(), d_adj.p, h->st);
        int adj[4];
        EU_CUDA(h, cudaMemcpyAsync(adj, d_adj.p, sizeof(adj), cudaMemcpyDeviceToHost, h->st));
        EU_CUDA(h, cudaStreamSynchronize(h->st));
        int down_max = adj[0], up_min = adj[1];
        const int mid = h->own_lo + (h->own_hi - h->own_lo)/2;
        std::vector<std::vector<int> > srcs(h->peers.size()), dsts(h->peers.size());
        for (size_t k = 0; k < h->peers.size(); ++k) {
            eu_solver::Peer* p = h->peers[k];
            if (p->n_send == 0) continue;
            srcs[k].resize(size_t(p->n_send)); dsts[k].resize(size_t(p->n_send));
            EU_CUDA(h, cudaMemcpy(srcs[k].data(), p->src.p, sizeof(int)*size_t(p->n_send), cudaMemcpyDeviceToHost));
            EU_CUDA(h, cudaMemcpy(dsts[k].data(), p->dst.p, sizeof(int)*size_t(p->n_send), cudaMemcpyDeviceToHost));
            for (int c : srcs[k]) {
                if (c < mid) down_max = std::max(down_max, c); else up_min = std::min(up_min, c);
            }
        }
        const int slice_lo = h->own_lo/EU_SLICE, slice_hi = (h->own_hi + EU_SLICE - 1)/EU_SLICE;
        const int a_hi = down_max >= 0 ? down_max/EU_SLICE + 1 : slice_lo;
        const int b_lo = up_min < INT_MAX ? up_min/EU_SLICE : slice_hi;
        bool ok = a_hi <= b_lo && h->n_wait <= 2;
        std::vector<int> dst[2];
        dst[0].assign(size_t(std::max(1, (a_hi - slice_lo)*EU_SLICE)), -1);
        dst[1].assign(size_t(std::max(1, (slice_hi - b_lo)*EU_SLICE)), -1);
        int range_peer[2] = { -1, -1 };
        for (size_t k = 0; k < h->peers.size() && ok; ++k) {
            for (size_t i = 0; i < srcs[k].size(); ++i) {
                const int c = srcs[k][i], sl = c/EU_SLICE;
                const int r = sl < a_hi ? 0 : (sl >= b_lo ? 1 : -1);
                if (r < 0 || (range_peer[r] >= 0 && range_peer[r] != int(k))) { ok = false; break; }
                range_peer[r] = int(k);
                int& slot = dst[r][size_t(c - (r == 0 ? slice_lo : b_lo)*EU_SLICE)];
                if (slot >= 0) { ok = false; break; }          // one cell, two ghost slots
                slot = dsts[k][i];
            }
        }
        if (ok) {
            int rc;
            for (int r = 0; r < 2; ++r) if ((rc = upload_vec(h, h->d_fused_dst[r], dst[r]))) return rc;
            EU_CUDA(h, h->d_fused_dummy.alloc(2));
            EU_CUDA(h, cudaMemset(h->d_fused_dummy.p, 0, 2*sizeof(unsigned)));
            const unsigned nA = unsigned(a_hi - slice_lo), nB = unsigned(slice_hi - b_lo);
            h->fused_a_hi = a_hi; h->fused_b_lo = b_lo;
            h->fused_down_max = down_max; h->fused_up_min = up_min;
            h->fused_peer[0] = range_peer[0]; h->fused_peer[1] = range_peer[1];
            if (range_peer[0] >= 0 && range_peer[0] == range_peer[1]) {
                h->fused_total[0] = h->fused_total[1] = nA + nB;       // one neighbour on both sides (periodic, 2 ranks)
            } else {
                h->fused_total[0] = range_peer[0] >= 0 ? nA : 0xffffffffu;
                h->fused_total[1] = range_peer[1] >= 0 ? nB : 0xffffffffu;
            }
            // every neighbour that expects pushes must be fed by a range
            for (size_t k = 0; k < h->peers.size(); ++k)
                if (h->peers[k]->n_send > 0 && range_peer[0] != int(k) && range_peer[1] != int(k)) ok = false;
        }
        h->fused_ok = ok;
    }
    h->comm_ready = true;
    return EU_OK;
}

int eu_comm_set_allreduce(eu_handle h, eu_allreduce_fn fn, void* user)
{
    if (!h) return EU_ERR_ARG;
    h->allreduce = fn;
    h->allreduce_user = user;
    return EU_OK;
}

} // extern "C"
