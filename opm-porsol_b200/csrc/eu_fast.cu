// FAST substep kernel: owner-computes gather over the SELL-32 records, static per-face
// quantities pre-contracted to scalars (G, T), rock curves as mobility tables in shared
// memory, fused explicit update + range check/clamp + capillary pressure of the new state.
// FMA contraction is allowed here; the result differs from the reference by ~1e-16 per
// substep (gate 1e-12), never in a step count (CFL is computed in eu_setup.cu, bit-exact).
//
// Reference semantics reproduced (euler/EulerUpstreamResidual_impl.hpp:100-300):
//   face evaluated with (lo, hi) = (lower, higher) cell index and the lo cell's flux q;
//   triv phase = water if G >= 0; upstream cell of the triv phase by sign of q; upstream cell
//   of the other phase by sign of q + sign*lambda_triv*G; viscous lambda_w/(lambda_w+lambda_o) q;
//   gravity (not on Dirichlet faces) lambda_w lambda_o/(lambda_t) G; capillary with mobilities
//   at the average saturation (averaged over the two rocks) times T (pc_hi - pc_lo);
//   residual[lo] -= dS, residual[hi] += dS; source; S += dt*residual/porevol; check/clamp.
// Per cell the kernel first loads all its records, then issues every gather (neighbour S, pc, face q/G/T)
// before any arithmetic, so a warp has 6-8 independent loads per lane in flight instead of a chain.
#include "eu_internal.h"
#include "eu_box_units.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// Neighbour mobilities in the marches: recomputed from the neighbours' saturations (five rock-curve lookups per cell,
// the default), or loaded as {lambda_w, lambda_o} pairs the previous substep stored (one lookup per cell, but 32 more
// bytes of HBM traffic per cell and substep: -DEU_STORED_LAM).  Build-time choice, measured both ways on B200
// (DESIGN.md section 5): recomputing is 10 % faster on the V+G bench grid and 1-5 % with the capillary term.
#ifdef EU_STORED_LAM
constexpr bool kStoredLam = true;
#else
constexpr bool kStoredLam = false;
#endif

#ifdef EU_DYNAMIC_ITEMS
__device__ unsigned g_item_counter;          // experiment: next unassigned work item of the running launch
#endif

constexpr int kWarpsPerBlock = 8;
constexpr int kBlock = kWarpsPerBlock*32;

// Rock curves in shared memory (FAST mode).  For every rock and table interval j the mobilities and
// J are kept in intercept/slope form  y(s) = a_j + b_j s  (a_j = y_j - b_j x_j):
//   coef[j]  = { a_w, b_w, a_o, b_o }   one 32-byte row, two LDS.128
//   jcoef[j] = { a_J, b_J }
//   xb[j]    = x_j, with the last node of each rock replaced by +inf (upper bound of the last interval)
//   bucket[rock*NB + k] = interval containing k/NB (NB chosen on the host so that a bucket holds at most
//                         one interior node: the interval is bucket or bucket+1)
// The interval found is exactly the one the reference's binary search returns (first/last interval
// outside the table).
extern __shared__ __align__(16) unsigned char eu_smem[];

struct TabLayout {
    int nn;          // nodes over all rocks
    int nb;          // buckets per rock
    double nbd;      // nb as a double (hoisted conversion)
    __device__ __forceinline__ const double4* coef() const { return reinterpret_cast<const double4*>(eu_smem); }
    __device__ __forceinline__ const double2* jcoef() const { return reinterpret_cast<const double2*>(eu_smem + size_t(32)*nn); }
    __device__ __forceinline__ const double* xb() const { return reinterpret_cast<const double*>(eu_smem + size_t(48)*nn); }
    __device__ __forceinline__ const int* offset() const { return reinterpret_cast<const int*>(eu_smem + size_t(56)*nn); }
    __device__ __forceinline__ const unsigned char* bucket() const { return eu_smem + size_t(56)*nn + 4*(EU_MAX_TABLES + 2); }
};

__device__ __forceinline__ void tables_to_smem(const EuTablesDev& t)
{
    const int nn = t.n_nodes_total;
    double4* coef = reinterpret_cast<double4*>(eu_smem);
    double2* jc = reinterpret_cast<double2*>(eu_smem + size_t(32)*nn);
    double* xb = reinterpret_cast<double*>(eu_smem + size_t(48)*nn);
    int* off = reinterpret_cast<int*>(eu_smem + size_t(56)*nn);
    unsigned char* bucket = eu_smem + size_t(56)*nn + 4*(EU_MAX_TABLES + 2);
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        coef[i] = make_double4(t.fcoef[4*i], t.fcoef[4*i + 1], t.fcoef[4*i + 2], t.fcoef[4*i + 3]);
        jc[i] = make_double2(t.fjcoef[2*i], t.fjcoef[2*i + 1]);
        xb[i] = t.fxb[i];
    }
    for (int i = threadIdx.x; i <= t.n_rocks; i += blockDim.x) off[i] = t.offset[i];
    for (int i = threadIdx.x; i < t.n_rocks*t.n_buckets; i += blockDim.x) bucket[i] = t.fbucket[i];
}

template <bool MULTIROCK>
__device__ __forceinline__ int interval(const TabLayout& L, int rock, double sat)
{
    int k = __double2int_rd(sat*L.nbd);
    k = min(max(k, 0), L.nb - 1);
    const int b = MULTIROCK ? L.offset()[rock] : 0;
    int j = b + L.bucket()[(MULTIROCK ? rock*L.nb : 0) + k];
    j += (sat >= L.xb()[j + 1]) ? 1 : 0;
    return j;
}

template <bool ROCKS, bool MULTIROCK>
struct Mob {
    // mobilities of both phases at (rock, sat)
    static __device__ __forceinline__ void both(const TabLayout& L, const EuTablesDev& t, int rock, double sat,
                                                double& lw, double& lo)
    {
        if (ROCKS) {
            const double4 c = L.coef()[interval<MULTIROCK>(L, rock, sat)];
            lw = fma(c.y, sat, c.x);
            lo = fma(c.w, sat, c.z);
        } else {
            lw = sat*sat*t.inv_visc[0];
            lo = (1.0 - sat)*(1.0 - sat)*t.inv_visc[1];
        }
    }
    // mobilities and capillary pressure at the same point: one interval search
    static __device__ __forceinline__ void both_and_pc(const TabLayout& L, const EuTablesDev& t, int rock, double sat, double scale,
                                                       double& lw, double& lo, double& pcv)
    {
        if (ROCKS) {
            const int j = interval<MULTIROCK>(L, rock, sat);
            const double4 c = L.coef()[j];
            const double2 cj = L.jcoef()[j];
            lw = fma(c.y, sat, c.x);
            lo = fma(c.w, sat, c.z);
            pcv = fma(cj.y, sat, cj.x)*scale;
        } else {
            lw = sat*sat*t.inv_visc[0];
            lo = (1.0 - sat)*(1.0 - sat)*t.inv_visc[1];
            pcv = 1e5*(1.0 - sat);
        }
    }
    static __device__ __forceinline__ double pc(const TabLayout& L, int rock, double sat, double scale)
    {
        if (ROCKS) {
            const double2 c = L.jcoef()[interval<MULTIROCK>(L, rock, sat)];
            return fma(c.y, sat, c.x)*scale;
        } else {
            return 1e5*(1.0 - sat);
        }
    }
};

// x/d for a positive, normal-range denominator (sum of two mobilities): reciprocal seed (>= 19 good bits), one
// Newton step (>= 38 bits) and a residual correction of the quotient, which squares the error once more
// (2^-76 + one rounding); no special-case branch.
__device__ __forceinline__ double div_pos(double x, double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    const double qv = x*r;
    const double rem = fma(-d, qv, x);
    return fma(rem, r, qv);
}

// One face of the gather, from the point of view of cell "self" (mobilities lw0/lo0) against the cell or
// boundary value on the other side (lw1/lo1).  own: self is the lower-index ("lo") cell, whose frame q, G
// and T are expressed in.  Returns the contribution to residual[self].  Everything decided at run time:
// used for the slots read from SELL records (boundaries, faults, irregular slices).
template <bool CAP>
__device__ __forceinline__ double face_contribution(bool own, bool interior, double q, double qq, double G,
                                                    double lw0, double lo0, double lw1, double lo1,
                                                    int method_viscous, int method_gravity,
                                                    double cap_coef /* lam_w lam_o / lam_t at the average saturation */,
                                                    double Tdpc /* T*(pc_hi - pc_lo) */)
{
    // straight-line code (selects only) so that the faces of a batch can be interleaved by the scheduler
    const bool triv_w = G >= 0.0;
    const double t0 = triv_w ? lw0 : lo0, t1 = triv_w ? lw1 : lo1;
    const double n0 = triv_w ? lo0 : lw0, n1 = triv_w ? lo1 : lw1;
    const bool u_self = (q >= 0.0) == own;                 // upstream cell of the trivial phase is self
    const double lam_t = u_self ? t0 : t1;
    const bool u2_self = (fma(lam_t, -fabs(G), q) >= 0.0) == own;
    const double lam_n = u2_self ? n0 : n1;
    const double lw = triv_w ? lam_t : lam_n;
    // lam_w (q + lam_o G)/(lam_w + lam_o), the gravity part only on interior / periodic faces
    double num = method_viscous ? lw*qq : 0.0;
    num = (method_gravity && interior) ? fma(lam_t*lam_n, G, num) : num;
    double dS = div_pos(num, lam_t + lam_n);
    if (CAP) dS = interior ? fma(cap_coef, Tdpc, dS) : dS;
    return own ? -dS : dS;
}

// The same face when the slot is regular: interior, and whether self is the lo cell is a compile-time
// property of the slot (odd slots of a slice class hold the positive offsets).  Returns dS in the lo cell's
// frame: the contribution is -dS to the lo cell and +dS to the hi cell.  G is zero when gravity is off.
template <bool OWN, bool CAP>
__device__ __forceinline__ double face_regular(double q, double qq, double G, double lw0, double lo0, double lw1, double lo1,
                                               int method_viscous, double cap_coef, double Tdpc)
{
    const double lwa = OWN ? lw0 : lw1, lwb = OWN ? lw1 : lw0;      // a = lo cell, b = hi cell
    const double loa = OWN ? lo0 : lo1, lob = OWN ? lo1 : lo0;
    const bool triv_w = G >= 0.0;
    const bool up_a = q >= 0.0;
    const double cw = up_a ? lwa : lwb, co = up_a ? loa : lob;
    const double lam_t = triv_w ? cw : co;
    const bool up2 = fma(lam_t, -fabs(G), q) >= 0.0;
    const double oa = triv_w ? loa : lwa, ob = triv_w ? lob : lwb;
    const double lam_n = up2 ? oa : ob;
    const double lw = triv_w ? lam_t : lam_n;
    double num = lw*qq;
    if (!method_viscous) num = 0.0;
    num = fma(lam_t*lam_n, G, num);
    double dS = div_pos(num, lam_t + lam_n);
    if (CAP) dS = fma(cap_coef, Tdpc, dS);
    return dS;
}

template <bool ROCKS, bool MULTIROCK>
__device__ __forceinline__ double cap_coefficient(const TabLayout& L, const EuTablesDev& t, int rock0, int rock1, double S0, double S1)
{
    const double Sa = 0.5*(S0 + S1);
    double lwa, loa;
    Mob<ROCKS, MULTIROCK>::both(L, t, rock0, Sa, lwa, loa);
    if (MULTIROCK && rock1 != rock0) {
        double lwb, lob;
        Mob<ROCKS, MULTIROCK>::both(L, t, rock1, Sa, lwb, lob);
        lwa = 0.5*(lwa + lwb);
        loa = 0.5*(loa + lob);
    }
    return div_pos(lwa*loa, lwa + loa);
}

// Diagonal tensor mobility (TENSOR, ReservoirPropertyCapillaryAnisotropicRelperm): FAST mode covers grids whose face
// normals are axis-aligned.  Then every term of the face flux (Residual_impl.hpp:205-262) reduces to the scalar
// formula with the mobility component of the face's axis: n.(M x) = n_a m_a x_a.  The curves of (rock r, axis a) are
// stored as table 3 r + a, the own cell's mobilities come per axis (OwnMob), the face's axis from f.axis8.
template <bool TENSOR>
struct OwnMob {
    double lw[TENSOR ? 3 : 1], lo[TENSOR ? 3 : 1];
    __device__ __forceinline__ double w(int ax) const { return TENSOR ? (ax == 0 ? lw[0] : (ax == 1 ? lw[TENSOR ? 1 : 0] : lw[TENSOR ? 2 : 0])) : lw[0]; }
    __device__ __forceinline__ double o(int ax) const { return TENSOR ? (ax == 0 ? lo[0] : (ax == 1 ? lo[TENSOR ? 1 : 0] : lo[TENSOR ? 2 : 0])) : lo[0]; }
};

template <bool ROCKS, bool MULTIROCK, bool TENSOR>
__device__ __forceinline__ OwnMob<TENSOR> own_mobilities(const TabLayout& L, const EuTablesDev& t, int rock0, double S0)
{
    OwnMob<TENSOR> m;
    if (TENSOR) {
#pragma unroll
        for (int ax = 0; ax < (TENSOR ? 3 : 1); ++ax) Mob<ROCKS, MULTIROCK>::both(L, t, 3*rock0 + ax, S0, m.lw[ax], m.lo[ax]);
    } else {
        Mob<ROCKS, MULTIROCK>::both(L, t, rock0, S0, m.lw[0], m.lo[0]);
    }
    return m;
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int W, int B, bool TENSOR>
__device__ __forceinline__ double gather_cell(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t,
                                              const EuFastDev& f, const EuStepArgs& a, const int2* __restrict__ recp,
                                              const int2* __restrict__ dscp, int width, int c, double S0, int rock0, double pc0,
                                              const OwnMob<TENSOR>& own0)
{
    // phase A: the cell's records.  Regular slots come from the 8-byte slice descriptor (neighbour = c + d,
    // face id affine in c): no per-cell record is read; irregular slots (boundaries, faults) load theirs.
    // All descriptors are fetched before any of them is looked at (they are warp-uniform broadcast loads).
    int2 dsc[W];
#pragma unroll
    for (int j = 0; j < W; ++j) dsc[j] = __ldg(dscp + (j < width ? j : 0));
    int2 r[W];
    bool any_explicit = false;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int nb = c + dsc[j].x;
        const bool regular = (j < width) && dsc[j].y >= 0;
        r[j].x = regular ? nb : EU_REC_PAD;
        r[j].y = dsc[j].y*f.n_local + (dsc[j].x > 0 ? c : nb);
        any_explicit |= (j < width) && dsc[j].y == -1;
    }
    if (any_explicit) {                       // warp-uniform: only slices at boundaries / faults
#pragma unroll
        for (int j = 0; j < W; ++j)
            if (j < width && dsc[j].y == -1) r[j] = __ldg(recp + j*EU_SLICE);
    }
    double acc = 0.0;
#pragma unroll
    for (int j0 = 0; j0 < W; j0 += B) {
        // phase B: every gather of the batch in flight at once
        double S1[B], T[CAP ? B : 1], pc1[CAP ? B : 1], nn[NN ? B : 1];
        double2 qg[B];
        int rk[MULTIROCK ? B : 1];
        int ax[TENSOR ? B : 1];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int j = j0 + k;
            S1[k] = 0.0; qg[k] = make_double2(0.0, 0.0);
            if (CAP) { T[k] = 0.0; pc1[k] = 0.0; }
            if (NN) nn[k] = 1.0;
            if (MULTIROCK) rk[k] = rock0;
            if (TENSOR) ax[k] = 0;
            if (j < W && r[j].x != EU_REC_PAD) {
                qg[k] = __ldg(f.qg + r[j].y);
                if (NN) nn[k] = __ldg(f.nn + r[j].y);
                if (TENSOR) ax[k] = __ldg(f.axis8 + r[j].y);
                if (r[j].x >= 0) {
                    S1[k] = __ldg(a.S_in + r[j].x);
                    if (MULTIROCK) rk[k] = __ldg(f.rock8 + r[j].x);
                    if (CAP) { T[k] = __ldg(f.T + r[j].y); pc1[k] = __ldg(a.pc_in + r[j].x); }
                } else {
                    S1[k] = __ldg(g.bnd_sat + (-2 - r[j].x));
                }
            }
        }
        // phase C: arithmetic, branch-free per face (an absent face contributes an exact zero)
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int j = j0 + k;
            if (j >= W) continue;                              // compile-time
            const bool valid = r[j].x != EU_REC_PAD;
            const bool interior = r[j].x >= 0;
            const bool own = !interior || c < r[j].x;
            const int axk = TENSOR ? ax[k] : 0;
            const int rk1 = MULTIROCK ? (TENSOR ? 3*rk[k] + axk : rk[k]) : 0;
            const int rk0 = TENSOR ? 3*rock0 + axk : rock0;
            double lw1, lo1;
            Mob<ROCKS, MULTIROCK>::both(L, t, rk1, S1[k], lw1, lo1);
            double cap_coef = 0.0, Tdpc = 0.0;
            if (CAP) {
                cap_coef = cap_coefficient<ROCKS, MULTIROCK>(L, t, rk0, rk1, S0, S1[k]);
                Tdpc = T[k]*(own ? (pc1[k] - pc0) : (pc0 - pc1[k]));
            }
            const double contrib = face_contribution<CAP>(own, interior, qg[k].x, NN ? qg[k].x*nn[k] : qg[k].x, qg[k].y,
                                                          own0.w(axk), own0.o(axk), lw1, lo1, a.method_viscous, a.method_gravity,
                                                          cap_coef, Tdpc);
            acc += valid ? contrib : 0.0;
        }
    }
    return acc;
}

// One face at a time from the SELL records: cells with more than 8 faces (slot_mask = all), and the explicit
// slots of a slice class (slot_mask = those slots).  Same arithmetic as gather_cell.
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, bool TENSOR>
__device__ __noinline__ double gather_cell_loop(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t,
                                                const EuFastDev& f, const EuStepArgs& a, const int2* __restrict__ recp,
                                                int width, unsigned slot_mask, int c, double S0, int rock0, double pc0,
                                                const OwnMob<TENSOR>& own0)
{
    double acc = 0.0;
    for (int j = 0; j < width; ++j) {
        if (j < 32 && !((slot_mask >> j) & 1u)) continue;
        const int2 r = recp[j*EU_SLICE];
        if (r.x == EU_REC_PAD) continue;
        const bool interior = r.x >= 0;
        const bool own = !interior || c < r.x;
        const double2 qg = f.qg[r.y];
        const double q = qg.x, G = qg.y;
        const double nn = NN ? f.nn[r.y] : 1.0;
        const double S1 = interior ? a.S_in[r.x] : g.bnd_sat[-2 - r.x];
        const int axk = TENSOR ? f.axis8[r.y] : 0;
        int rk = (MULTIROCK && interior) ? f.rock8[r.x] : rock0;
        int rk0 = rock0;
        if (TENSOR) { rk = 3*rk + axk; rk0 = 3*rock0 + axk; }
        double lw1, lo1;
        Mob<ROCKS, MULTIROCK>::both(L, t, rk, S1, lw1, lo1);
        double cap_coef = 0.0, Tdpc = 0.0;
        if (CAP && interior) {
            cap_coef = cap_coefficient<ROCKS, MULTIROCK>(L, t, rk0, rk, S0, S1);
            const double pc1 = a.pc_in[r.x];
            Tdpc = f.T[r.y]*(own ? (pc1 - pc0) : (pc0 - pc1));
        }
        acc += face_contribution<CAP>(own, interior, q, q*nn, G, own0.w(axk), own0.o(axk), lw1, lo1, a.method_viscous,
                                      a.method_gravity, cap_coef, Tdpc);
    }
    return acc;
}

// source term (:293-299), explicit update (EulerUpstream_impl.hpp:371-374), range check / clamp (:336-349) and the
// capillary pressure of the new state (Residual_impl.hpp:459-467, fused); returns the new saturation
template <bool ROCKS, bool MULTIROCK, bool CAP, bool TENSOR>
__device__ __forceinline__ double finish_cell(const TabLayout& L, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                                              int c, double S0, int rock0, const OwnMob<TENSOR>& own0, double inv_pv, double acc,
                                              double& pcn, bool pcs_given = false, double pcs = 1.0)
{
    if (a.n_src > 0) {
        double rate = 0.0;
        int lo = 0, hi = a.n_src;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.src_cell[mid] < c) lo = mid + 1; else hi = mid; }
        if (lo < a.n_src && a.src_cell[lo] == c) rate = a.src_rate[lo];
        if (rate < 0.0) {
            if (TENSOR) {        // fractionalFlow of the tensor class: mean over the three directions (..AnisotropicRelperm_impl.hpp:57-72)
                double ff = 0.0;
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) ff += own0.w(ax)/(own0.w(ax) + own0.o(ax));
                rate *= ff/3.0;
            } else {
                rate *= own0.w(0)/(own0.w(0) + own0.o(0));
            }
        }
        acc += rate;
    }
    if (a.residual_out) a.residual_out[c] = acc;
    double sat = fma(a.dt*acc, inv_pv, S0);
    if (a.check_sat || a.clamp_sat) {
        if (sat > 1.0 || sat < 0.0) {
            if (a.clamp_sat) {
                sat = fmax(fmin(sat, 1.0), 0.0);
            } else if (sat > 1.001 || sat < -0.001) {
                atomicMin(a.fail_key, ((unsigned long long)(unsigned)a.substep << 32) | (unsigned)c);
            }
        }
    }
    a.S_out[c] = sat;
    // mobilities (and capillary pressure) of the new state, looked up once per cell: the marches of the next substep
    // read their neighbours' mobilities instead of evaluating the rock curves again
    double lwn, lon;
    pcn = 0.0;
    if (TENSOR) {                // no marches for this class: only the capillary pressure (table 3 r holds the pc column)
        if (CAP) {
            pcn = Mob<ROCKS, MULTIROCK>::pc(L, 3*rock0, sat, 1.0);
            a.pc_out[c] = pcn;
        }
        return sat;
    }
    if (!kStoredLam) {           // no pair arrays: only the capillary pressure of the new state
        if (CAP) {
            pcn = Mob<ROCKS, MULTIROCK>::pc(L, rock0, sat, ROCKS ? (pcs_given ? pcs : f.pcscale[c]) : 1.0);
            a.pc_out[c] = pcn;
        }
        return sat;
    }
    if (CAP) {
        Mob<ROCKS, MULTIROCK>::both_and_pc(L, t, rock0, sat, ROCKS ? f.pcscale[c] : 1.0, lwn, lon, pcn);
        a.pc_out[c] = pcn;
    } else {
        Mob<ROCKS, MULTIROCK>::both(L, t, rock0, sat, lwn, lon);
    }
    a.lam_out[c] = make_double2(lwn, lon);
    return sat;
}

// ---- march along a slice class ---------------------------------------------------------------------------
// A work item of a slice class is `len` slices s, s + D/32, ... of the same class (EuSliceClass in eu_internal.h).
// Per cell the regular code evaluates slots 0..3 and 5; slot 4 only for the first cell of the march: afterwards it
// is the previous cell's slot 5, carried in registers together with that neighbour's saturation, mobilities, rock
// and capillary pressure.  The neighbours' mobilities come from the per-cell {lambda_w, lambda_o} pairs the previous
// substep stored (finish_cell), so a step needs per face one 16-byte pair of the neighbour and one 16-byte {q, G}
// pair of the face, all issued as one batch of loads before any arithmetic.
__device__ __forceinline__ double2 ldg_pair(const double2* p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_f64(const double* p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
// L2 eviction policy for the arrays a substep streams through exactly once ({q, G} pairs, 1/porevol, T): evict-first,
// so that they do not displace the neighbours' mobility pairs other warps are about to read.
__device__ __forceinline__ unsigned long long stream_policy(bool evict_first)
{
    unsigned long long p;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else             asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double2 ldg_pair_stream(const double2* p, unsigned long long pol)
{
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ldg_f64_stream(const double* p, unsigned long long pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

// state carried along a march
struct MarchCarry {
    double S0, lw0, lo0, pc0, dS4;     // this cell; dS4 = flux of its slot-4 face in the lo cell's frame
    int rock0;
};

// one regular face: capillary coefficient (if any) and dS in the lo cell's frame
template <bool ROCKS, bool MULTIROCK, bool CAP, bool OWN>
__device__ __forceinline__ double regular_slot(const TabLayout& L, const EuTablesDev& t, const EuStepArgs& a, const MarchCarry& m,
                                               double2 lam1, double S1, int rk1, double2 qg, double nn, bool use_nn, double T,
                                               double pc1)
{
    double cap_coef = 0.0, Tdpc = 0.0;
    if (CAP) {
        cap_coef = cap_coefficient<ROCKS, MULTIROCK>(L, t, m.rock0, rk1, m.S0, S1);
        Tdpc = T*(OWN ? (pc1 - m.pc0) : (m.pc0 - pc1));
    }
    return face_regular<OWN, CAP>(qg.x, use_nn ? qg.x*nn : qg.x, qg.y, m.lw0, m.lo0, lam1.x, lam1.y, a.method_viscous, cap_coef, Tdpc);
}

// the first cell of a march: its own state and the slot-4 face
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN>
__device__ __forceinline__ void march_head(const TabLayout& L, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                                           const EuSliceClass* __restrict__ cl, int c, MarchCarry& m)
{
    const int nb = c + cl->nb_off[4];
    const int fid = c*cl->fid_mul[4] + cl->fid_off[4];
    double2 lam0, lam1;
    const double2 qg = ldg_pair(f.qg + fid);
    m.S0 = ldg_f64(a.S_in + c);
    const double S1 = (CAP || !kStoredLam) ? __ldg(a.S_in + nb) : 0.0;
    const double nn = NN ? __ldg(f.nn + fid) : 1.0;
    m.rock0 = MULTIROCK ? __ldg(f.rock8 + c) : 0;
    const int rk1 = (MULTIROCK && (CAP || !kStoredLam)) ? __ldg(f.rock8 + nb) : 0;
    if (kStoredLam) {
        lam0 = ldg_pair(a.lam_in + c);
        lam1 = ldg_pair(a.lam_in + nb);
    } else {
        Mob<ROCKS, MULTIROCK>::both(L, t, m.rock0, m.S0, lam0.x, lam0.y);
        Mob<ROCKS, MULTIROCK>::both(L, t, rk1, S1, lam1.x, lam1.y);
    }
    m.pc0 = CAP ? __ldg(a.pc_in + c) : 0.0;
    const double T = CAP ? __ldg(f.T + fid) : 0.0, pc1 = CAP ? __ldg(a.pc_in + nb) : 0.0;
    m.lw0 = lam0.x; m.lo0 = lam0.y;
    m.dS4 = regular_slot<ROCKS, MULTIROCK, CAP, false>(L, t, a, m, lam1, S1, rk1, qg, nn, NN, T, pc1);
}

// one cell of a march; advances the carry to the cell across slot 5
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, bool RECORDS>
__device__ __forceinline__ void march_step(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f,
                                           const EuStepArgs& a, const EuSliceClass* __restrict__ cl, int c, int lane, int ahead,
                                           MarchCarry& m)
{
    const unsigned long long pol = stream_policy(f.l2_hint != 0);     // (live only while the loads are issued)
    constexpr int kOrder[5] = { 5, 3, 2, 1, 0 };       // far neighbours first: their lines take longest to arrive
    int nb[6], fid[6];
    double2 lam[6], qg[6];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int j = kOrder[k];
        nb[j] = c + cl->nb_off[j];
        fid[j] = c*cl->fid_mul[j] + cl->fid_off[j];
        if (kStoredLam) lam[j] = ldg_pair(a.lam_in + nb[j]);
        qg[j] = ldg_pair_stream(f.qg + fid[j], pol);
    }
    const double S5 = ldg_f64(a.S_in + nb[5]);
    const double inv_pv = ldg_f64_stream(f.inv_porevol + c, pol);
    if (ahead > 0) {
        // the lines the next cell of the march (c + D) will load, requested into L2 now: its z+ and y neighbours'
        // mobility pairs (the x neighbours share the lines of lam[c + D], loaded above as slot 5), the {q, G} pairs of
        // its faces, its z+ neighbour's saturation and its 1/porevol.  No registers are tied up by these requests.
        const int D = cl->D*ahead;
        if (kStoredLam) {
            prefetch_l2(a.lam_in + nb[5] + D);
            prefetch_l2(a.lam_in + nb[3] + D);
            prefetch_l2(a.lam_in + nb[2] + D);
        } else {
            prefetch_l2(a.S_in + nb[3] + D);
            prefetch_l2(a.S_in + nb[2] + D);
        }
        prefetch_l2(f.qg + fid[5] + D*cl->fid_mul[5]);
        prefetch_l2(f.qg + fid[3] + D*cl->fid_mul[3]);
        prefetch_l2(f.qg + fid[2] + D*cl->fid_mul[2]);
        prefetch_l2(f.qg + fid[1] + D*cl->fid_mul[1]);
        prefetch_l2(a.S_in + nb[5] + D);
        prefetch_l2(f.inv_porevol + c + D);
        if (CAP) {
            prefetch_l2(a.pc_in + nb[5] + D);
            prefetch_l2(f.T + fid[5] + D*cl->fid_mul[5]);
            prefetch_l2(f.T + fid[3] + D*cl->fid_mul[3]);
            prefetch_l2(f.T + fid[1] + D*cl->fid_mul[1]);
            prefetch_l2(f.pcscale + c + D);
        }
    }
    double S1[6], T[6], pc1[6], nn[6];
    int rk[6];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int j = kOrder[k];
        S1[j] = ((CAP || !kStoredLam) && j != 5) ? __ldg(a.S_in + nb[j]) : S5;
        rk[j] = (MULTIROCK && (CAP || !kStoredLam || j == 5)) ? __ldg(f.rock8 + nb[j]) : 0;
        nn[j] = NN ? __ldg(f.nn + fid[j]) : 1.0;
        T[j] = CAP ? __ldg(f.T + fid[j]) : 0.0;
        pc1[j] = CAP ? __ldg(a.pc_in + nb[j]) : 0.0;
    }
    if (!kStoredLam) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int j = kOrder[k];
            Mob<ROCKS, MULTIROCK>::both(L, t, rk[j], S1[j], lam[j].x, lam[j].y);
        }
    }
    double acc = m.dS4;                 // self is the hi cell of the slot-4 face: +dS
#ifdef EU_EXP_NOARITH               // timing experiment: the memory access pattern without the face arithmetic
    m.dS4 = 0.0;
    for (int k = 0; k < 5; ++k) { const int j = kOrder[k]; acc += 1e-30*(lam[j].x + lam[j].y + qg[j].x + qg[j].y); }
    if (false)
#endif
    {
    m.dS4 = regular_slot<ROCKS, MULTIROCK, CAP, true>(L, t, a, m, lam[5], S1[5], rk[5], qg[5], nn[5], NN, T[5], pc1[5]);
    acc -= m.dS4;
    acc -= regular_slot<ROCKS, MULTIROCK, CAP, true>(L, t, a, m, lam[3], S1[3], rk[3], qg[3], nn[3], NN, T[3], pc1[3]);
    acc += regular_slot<ROCKS, MULTIROCK, CAP, false>(L, t, a, m, lam[2], S1[2], rk[2], qg[2], nn[2], NN, T[2], pc1[2]);
    acc -= regular_slot<ROCKS, MULTIROCK, CAP, true>(L, t, a, m, lam[1], S1[1], rk[1], qg[1], nn[1], NN, T[1], pc1[1]);
    acc += regular_slot<ROCKS, MULTIROCK, CAP, false>(L, t, a, m, lam[0], S1[0], rk[0], qg[0], nn[0], NN, T[0], pc1[0]);
    }
    if (RECORDS) {                      // boundary / fault faces of this class, from the SELL records
        const int base = f.slice_base[c >> 5];
        OwnMob<false> own0;
        own0.lw[0] = m.lw0; own0.lo[0] = m.lo0;
        acc += gather_cell_loop<ROCKS, MULTIROCK, CAP, NN, false>(L, g, t, f, a, f.rec + base + lane, 32 - __clz(cl->rec_mask),
                                                                  (unsigned)cl->rec_mask, c, m.S0, m.rock0, m.pc0, own0);
    }
    double pcn;
    {
        OwnMob<false> own0;
        own0.lw[0] = m.lw0; own0.lo[0] = m.lo0;
        finish_cell<ROCKS, MULTIROCK, CAP, false>(L, t, f, a, c, m.S0, m.rock0, own0, inv_pv, acc, pcn);
    }
    m.S0 = S5; m.lw0 = lam[5].x; m.lo0 = lam[5].y; m.rock0 = rk[5]; m.pc0 = pc1[5];
}

// Items of classes with explicit slots run the same march out of line, so that the call of the record loop does
// not constrain the register allocation of the common path.
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, bool RECORDS>
__device__ __forceinline__ void march_item_impl(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f,
                                                const EuStepArgs& a, const EuSliceClass* __restrict__ cl, int s, int len, int lane)
{
    int c = s*EU_SLICE + lane;
    const int D = cl->D;
    MarchCarry m;
    march_head<ROCKS, MULTIROCK, CAP, NN>(L, t, f, a, cl, c, m);
    for (int i = 0; i < len; ++i) {
        // prefetch distance in march steps (0 = off); never past the last cell of the item
        const int ahead = (i + f.prefetch < len) ? f.prefetch : 0;
        march_step<ROCKS, MULTIROCK, CAP, NN, RECORDS>(L, g, t, f, a, cl, c, lane, ahead, m);
        c += D;
    }
}
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN>
__device__ __noinline__ void march_item_records(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f,
                                                const EuStepArgs& a, const EuSliceClass* __restrict__ cl, int s, int len, int lane)
{
    march_item_impl<ROCKS, MULTIROCK, CAP, NN, true>(L, g, t, f, a, cl, s, len, lane);
}

// One slice through its descriptors / records: slices next to a slab boundary (with the halo push), slices that
// straddle the own range, irregular slices (faults, cells with other than six faces).
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int B6, int B8, bool TENSOR>
__device__ __forceinline__ void slice_generic(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f,
                                              const EuStepArgs& a, const EuHaloDev& halo, int s, int range, int lane,
                                              int slice_lo)
{
    const int c = s*EU_SLICE + lane;
    const bool active = (c >= g.own_lo) && (c < g.own_hi);
    if (active) {
        const int base = f.slice_base[s];
        const int width = (f.slice_base[s + 1] - base) >> 5;
        const double S0 = a.S_in[c];
        const int rock0 = MULTIROCK ? f.rock8[c] : 0;
        const double pc0 = CAP ? a.pc_in[c] : 0.0;
        const double inv_pv = f.inv_porevol[c];
        const int2* __restrict__ recp = f.rec + base + lane;
        const int2* __restrict__ dscp = f.desc + (base >> 5);
        double acc;
        const OwnMob<TENSOR> own0 = own_mobilities<ROCKS, MULTIROCK, TENSOR>(L, t, rock0, S0);
        if (width <= 6)      acc = gather_cell<ROCKS, MULTIROCK, CAP, NN, 6, B6, TENSOR>(L, g, t, f, a, recp, dscp, width, c, S0, rock0, pc0, own0);
        else if (width <= 8) acc = gather_cell<ROCKS, MULTIROCK, CAP, NN, 8, B8, TENSOR>(L, g, t, f, a, recp, dscp, width, c, S0, rock0, pc0, own0);
        else                 acc = gather_cell_loop<ROCKS, MULTIROCK, CAP, NN, TENSOR>(L, g, t, f, a, recp, width, 0xffffffffu, c, S0, rock0, pc0, own0);
        double pcn;
        const double sat = finish_cell<ROCKS, MULTIROCK, CAP, TENSOR>(L, t, f, a, c, S0, rock0, own0, inv_pv, acc, pcn);
        if (range >= 0) {
            // this cell is a ghost of the neighbour rank: store it into the neighbour's HBM as well (made visible by the
            // one system fence the warp issues after its last boundary slice, see k_fast_step)
            const int first = (range == 0 ? slice_lo : halo.b_lo)*EU_SLICE;
            const int d = halo.dst[range][c - first];
            if (d >= 0) {
                halo.peer_S[range][d] = sat;
                if (CAP && halo.peer_pc[range]) halo.peer_pc[range][d] = pcn;
            }
        }
    }
}

// Persistent grid.  Processing order: the two slice ranges next to slab boundaries (multi-GPU runs; their results
// are also pushed into the neighbours' ghost cells), then the work items of the interior, item v to warp v mod #warps.
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int B6, int B8, int MINB, bool TENSOR>
__global__ void __launch_bounds__(kBlock, MINB) k_fast_step(EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a,
                                                            EuHaloDev halo, int slice_lo, int slice_hi, int class_smem_offset)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.nbd = double(t.n_buckets);
    if (ROCKS) tables_to_smem(t);
    EuSliceClass* classes = reinterpret_cast<EuSliceClass*>(eu_smem + class_smem_offset);
    {
        const int n_words = f.n_classes*int(sizeof(EuSliceClass)/sizeof(int));
        const int* src = reinterpret_cast<const int*>(f.classes);
        int* dst = reinterpret_cast<int*>(classes);
        for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (!halo.enabled) {   // (with a halo exchange every rank keeps stepping so that the flags stay in lockstep)
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x*kWarpsPerBlock + (threadIdx.x >> 5);
    const int n_warps = gridDim.x*kWarpsPerBlock;
    const int nAB = halo.enabled ? (halo.a_hi - slice_lo) + (slice_hi - halo.b_lo) : 0;
    const int nA = halo.enabled ? halo.a_hi - slice_lo : 0;
    int v = warp_global;
    if (v < nAB) {
        // ghosts of the previous substep must have landed before a boundary slice reads them
        if (lane < halo.n_wait) {
            const volatile unsigned* fl = halo.my_flags + halo.wait_rank[lane];
            const long long t0 = clock64();
            while ((int)(*fl - (halo.epoch - 1u)) < 0) {
                __nanosleep(100);
                if (clock64() - t0 > halo.timeout_cycles) { atomicExch(halo.err_flag, 1); break; }
            }
            __threadfence_system();
        }
        __syncwarp();
        unsigned done[2] = { 0u, 0u };
        for (; v < nAB; v += n_warps) {
            const int range = v < nA ? 0 : 1;
            const int s = v < nA ? slice_lo + v : halo.b_lo + (v - nA);
            slice_generic<ROCKS, MULTIROCK, CAP, NN, B6, B8, TENSOR>(L, g, t, f, a, halo, s, range, lane, slice_lo);
            ++done[range];
        }
        // One system-scope fence per warp for all its pushes, then the finished-slice counters; the warp that completes
        // a neighbour's share publishes the epoch in that neighbour's flag word.
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (done[r] == 0u) continue;
                const unsigned before = atomicAdd(halo.counter[r], done[r]);
                if (before + done[r] == halo.total[r]) {
                    *halo.counter[r] = 0u;
                    __threadfence_system();
                    *(volatile unsigned*)halo.peer_flag[r] = halo.epoch;
                    __threadfence_system();
                }
            }
        }
    }
    // ---- interior items
    int vi = v - nAB;
    if (vi >= f.n_items) return;
    int2 item = __ldg(f.items + vi);
#ifdef EU_DYNAMIC_ITEMS
    // EXPERIMENT (not the default build, unmeasured): after its first item -- assigned statically, so that the warps of
    // a block still start on neighbouring grid rows -- a warp takes the next unassigned item from a per-launch counter
    // instead of item vi + #warps: no warp idles while items are left.  The host zeroes the counter before the launch.
    for (;;) {
        const int cls = int(unsigned(item.y) >> 16);
        if (TENSOR || cls == EU_ITEM_GENERIC) {
            slice_generic<ROCKS, MULTIROCK, CAP, NN, B6, B8, TENSOR>(L, g, t, f, a, halo, item.x, -1, lane, slice_lo);
        } else {
            const EuSliceClass* cl = classes + cls;
            if (cl->rec_mask != 0) march_item_records<ROCKS, MULTIROCK, CAP, NN>(L, g, t, f, a, cl, item.x, item.y & 0xffff, lane);
            else                   march_item_impl<ROCKS, MULTIROCK, CAP, NN, false>(L, g, t, f, a, cl, item.x, item.y & 0xffff, lane);
        }
        int next = 0;
        if (lane == 0) next = n_warps + int(atomicAdd(&g_item_counter, 1u));
        next = __shfl_sync(0xffffffffu, next, 0);
        if (next >= f.n_items) break;
        item = __ldg(f.items + next);
    }
#else
    for (;;) {
        // the next item of this warp is fetched before the current one is processed
        const bool has_next = vi + n_warps < f.n_items;
        int2 item_next = item;
        if (has_next) item_next = __ldg(f.items + vi + n_warps);
        const int cls = int(unsigned(item.y) >> 16);
        if (TENSOR || cls == EU_ITEM_GENERIC) {       // (no slice classes are built for the tensor class)
            slice_generic<ROCKS, MULTIROCK, CAP, NN, B6, B8, TENSOR>(L, g, t, f, a, halo, item.x, -1, lane, slice_lo);
        } else {
            const EuSliceClass* cl = classes + cls;
            if (cl->rec_mask != 0) march_item_records<ROCKS, MULTIROCK, CAP, NN>(L, g, t, f, a, cl, item.x, item.y & 0xffff, lane);
            else                   march_item_impl<ROCKS, MULTIROCK, CAP, NN, false>(L, g, t, f, a, cl, item.x, item.y & 0xffff, lane);
        }
        if (!has_next) break;
        item = item_next;
        vi += n_warps;
    }
#endif
}

// ---- diagonal tensor mobility on general (oblique) face normals ---------------------------------------------
// ReservoirPropertyCapillaryAnisotropicRelperm on corner-point geometry.  With M = diag(m_x, m_y, m_z) every
// mobility-weighted projection of the face flux (Residual_impl.hpp:205-262) is a sum over the axes,
//   n.(M x) = sum_k n_k m_k x_k,
// with the mobility-independent factors contracted per axis at upload (k_contract_t3: Gv, n_k^2, Tv):
//   upstream of the non-trivial phase:  q -+ sum_k lam_t,k Gv_k          (:216-221)
//   viscous   q sum_k n_k^2 lam_w,k / (lam_w,k + lam_o,k)                 (:241-249)
//   gravity   sum_k Gv_k lam_w,k lam_o,k / (lam_w,k + lam_o,k)            (:252-262)
//   capillary (pc_hi - pc_lo) sum_k Tv_k lam_w,k lam_o,k / (..) at the average saturation (:265-272)
// The trivial phase is decided by the sign of the scalar G of the reference (:208-212), which k_contract stores.
// One warp per slice, one face at a time from the SELL records (the structure of gather_cell_loop); the three
// curve sets of a rock share their nodes, so one interval search serves the three axes.
__device__ __forceinline__ void mob3(const TabLayout& L, int rock, double sat, double w[3], double o[3])
{
    const int s0 = 3*rock;
    const int jrel = interval<true>(L, s0, sat) - L.offset()[s0];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const double4 c = L.coef()[L.offset()[s0 + ax] + jrel];
        w[ax] = fma(c.y, sat, c.x);
        o[ax] = fma(c.w, sat, c.z);
    }
}

template <bool CAP>
__global__ void __launch_bounds__(kBlock, 2) k_fast_step_t3(EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a,
                                                            int slice_lo, int slice_hi)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.nbd = double(t.n_buckets);
    tables_to_smem(t);
    __syncthreads();
    {
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x*kWarpsPerBlock;
    for (int s = slice_lo + blockIdx.x*kWarpsPerBlock + (threadIdx.x >> 5); s < slice_hi; s += n_warps) {
        const int c = s*EU_SLICE + lane;
        if (c < g.own_lo || c >= g.own_hi) continue;
        const int base = f.slice_base[s];
        const int width = (f.slice_base[s + 1] - base) >> 5;
        const double S0 = a.S_in[c];
        const int rock0 = f.rock8[c];
        const double pc0 = CAP ? a.pc_in[c] : 0.0;
        double w0[3], o0[3];
        mob3(L, rock0, S0, w0, o0);
        double acc = 0.0;
        for (int j = 0; j < width; ++j) {
            const int2 r = f.rec[base + j*EU_SLICE + lane];
            if (r.x == EU_REC_PAD) continue;
            const bool interior = r.x >= 0;
            const bool own = !interior || c < r.x;
            const double2 qg = f.qg[r.y];
            const double q = qg.x, G = qg.y;
            double Gv[3], nn[3], Tv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Gv[k] = f.fv[(long long)k*f.fv_stride + r.y];
                nn[k] = f.fv[(long long)(3 + k)*f.fv_stride + r.y];
                Tv[k] = (CAP && interior) ? f.fv[(long long)(6 + k)*f.fv_stride + r.y] : 0.0;
            }
            const double S1 = interior ? a.S_in[r.x] : g.bnd_sat[-2 - r.x];
            const int rock1 = interior ? f.rock8[r.x] : rock0;
            double w1[3], o1[3];
            mob3(L, rock1, S1, w1, o1);
            // upstream mobilities: trivial phase by the sign of q, the other phase by the sign of q -+ sum lam_t Gv
            const bool triv_w = G >= 0.0;
            const bool u_self = (q >= 0.0) == own;
            double lt[3], gsum = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double t0 = triv_w ? w0[k] : o0[k], t1 = triv_w ? w1[k] : o1[k];
                lt[k] = u_self ? t0 : t1;
                gsum = fma(lt[k], Gv[k], gsum);
            }
            const bool u2_self = ((triv_w ? q - gsum : q + gsum) >= 0.0) == own;
            double visc = 0.0, grav = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double n0 = triv_w ? o0[k] : w0[k], n1 = triv_w ? o1[k] : w1[k];
                const double ln = u2_self ? n0 : n1;
                const double lw = triv_w ? lt[k] : ln;
                const double rtot = div_pos(1.0, lt[k] + ln);
                visc = fma(nn[k], lw*rtot, visc);
                grav = fma(Gv[k], lt[k]*ln*rtot, grav);
            }
            double dS = a.method_viscous ? q*visc : 0.0;
            if (a.method_gravity && interior) dS += grav;
            if (CAP && interior) {
                // mobilities at the average saturation, averaged over the two rocks (:224-239)
                const double Sa = 0.5*(S0 + S1);
                double wa[3], oa[3];
                mob3(L, rock0, Sa, wa, oa);
                if (rock1 != rock0) {
                    double wb[3], ob[3];
                    mob3(L, rock1, Sa, wb, ob);
#pragma unroll
                    for (int k = 0; k < 3; ++k) { wa[k] = 0.5*(wa[k] + wb[k]); oa[k] = 0.5*(oa[k] + ob[k]); }
                }
                double cap = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) cap = fma(Tv[k], div_pos(wa[k]*oa[k], wa[k] + oa[k]), cap);
                const double pc1 = a.pc_in[r.x];
                dS = fma(cap, own ? (pc1 - pc0) : (pc0 - pc1), dS);
            }
            acc += own ? -dS : dS;
        }
        OwnMob<true> own0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { own0.lw[k] = w0[k]; own0.lo[k] = o0[k]; }
        double pcn;
        finish_cell<true, true, CAP, true>(L, t, f, a, c, S0, rock0, own0, f.inv_porevol[c], acc, pcn);
    }
}

// capillary pressure and mobilities of a given state (start of an attempt; afterwards the substep kernel keeps
// them current for the cells it updates, and the halo exchange for the ghosts)
template <bool ROCKS, bool MULTIROCK>
__global__ void __launch_bounds__(kBlock) k_fast_state(EuGridDev g, EuTablesDev t, EuFastDev f,
                                                       const double* __restrict__ S, double* __restrict__ pc,
                                                       double2* __restrict__ lam, int lo, int hi)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.nbd = double(t.n_buckets);
    if (ROCKS) {
        tables_to_smem(t);
        __syncthreads();
    }
    for (int c = lo + blockIdx.x*blockDim.x + threadIdx.x; c < hi; c += gridDim.x*blockDim.x) {
        double lw, lo_, pcv;
        // (tensor mobility: table 3 r holds the pc column; the pairs are not used by that class)
        const int table = (MULTIROCK ? f.rock8[c] : 0)*((f.axis8 || f.fv) ? 3 : 1);
        Mob<ROCKS, MULTIROCK>::both_and_pc(L, t, table, S[c], ROCKS ? f.pcscale[c] : 1.0, lw, lo_, pcv);
        if (pc) pc[c] = pcv;
        if (lam) lam[c] = make_double2(lw, lo_);
    }
}

} // namespace

#include "eu_tile.cuh"

// ---- box kernel: host side -------------------------------------------------------------------------------------
struct EuBoxPlan {
    int nx = 0, ny = 0, nz = 0, z_lo = 0, z_hi = 0, n_local = 0;
    int tx = 0, ty = 0, threads = 0;
    CUtensorMap mapS[2], mapPc[2], mapQ, mapG, mapT, mapA, mapV;     // A: acc_irr, V: 1 / pore volume
    int g_mask = 7;                                  // axis planes with a gravity component (set after the contraction)
    int4* d_units = nullptr;
    int* d_unit_start = nullptr;                    // [n_blocks + 1] unit range of each block
    int four_blocks = -1;                           // without the capillary term: 1 = two stages, four blocks per SM (-1: not yet known)
    int n_units = 0, n_blocks = 0, n_bnd_units[2] = { 0, 0 }, n_flagged = 0;
    int units_key[5] = { -1, -1, -1, -1, -1 };      // (bnd planes lo, hi, grid blocks, lz override, cap) the unit list was built for
    const unsigned short* cmask = nullptr;
    int g_mask_seen = -1;
    size_t smem_set[32] = { 0 };
    int blocks_per_sm[32] = { 0 };
    const int* irr_cells = nullptr;
    int n_irr = 0;
    double* acc_irr = nullptr;
    int n_sms = 148;
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    }
    return fn;
}

// tensor map over doubles: dims / box innermost first, strides in bytes for dims 1..rank-1
bool make_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint32_t ones[5] = { 1, 1, 1, 1, 1 };
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, cuuint32_t(rank), base, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int r128(int v) { return (v + 127) & ~127; }

// shared-memory layout of the box kernel for a tile shape
struct BoxLayout { EuBoxDev b; size_t total; };
BoxLayout box_layout(const EuBoxPlan& p, bool cap, bool multirock, int stages, size_t tab_bytes, bool share = false)
{
    BoxLayout o;
    std::memset(&o, 0, sizeof(o));
    EuBoxDev& b = o.b;
    b.nx = p.nx; b.ny = p.ny; b.nz = p.nz; b.tx = p.tx; b.ty = p.ty;
    b.stages = stages;
    const int cells = (p.tx + 2)*(p.ty + 2);
    b.lam_bytes = cap ? 2*r128(cells*16) : r128(cells*16);      // {lw, lo} entries, then (capillary) {S, pc} entries
    b.rk_bytes = (cap && multirock) ? r128(cells) : 0;
    b.T_bytes = r128((p.tx + 2)*(p.ty + 1)*8);               // one box of q, G or T
    b.g_mask = p.g_mask & 7;
    const int n_G = ((b.g_mask >> 0) & 1) + ((b.g_mask >> 1) & 1) + ((b.g_mask >> 2) & 1);
    const int S_bytes = r128((p.tx + 4)*(p.ty + 2)*8);
    b.off_S = 0;
    b.off_pc = S_bytes;
    b.off_q = S_bytes*(cap ? 2 : 1);
    b.off_G = b.off_q + 3*b.T_bytes;
    b.off_T = b.off_G + n_G*b.T_bytes;
    b.A_bytes = r128(p.tx*p.ty*8);                           // tile boxes of acc_irr and 1 / pore volume
    b.off_A = b.off_T + (cap ? 3*b.T_bytes : 0);
    b.off_V = b.off_A + b.A_bytes;
    b.stage_bytes = b.off_V + b.A_bytes;
    b.off_bar = 0;
    b.off_lam = 128;
    b.off_rk = b.off_lam + 3*b.lam_bytes;
    b.fx_bytes = share ? r128(p.ty*(p.tx + 1)*8) : 0;
    b.fy_bytes = share ? r128((p.ty + 1)*p.tx*8) : 0;
    b.off_fx = b.off_rk + 3*b.rk_bytes;
    b.off_fy = b.off_fx + 2*b.fx_bytes;
    b.off_stage = b.off_fy + 2*b.fy_bytes;
    o.total = tab_bytes + 128 /* alignment slack */ + size_t(b.off_stage) + size_t(stages)*size_t(b.stage_bytes);
    return o;
}

} // namespace

void eu_box_plan_destroy(EuBoxPlan* p)
{
    if (!p) return;
    if (p->d_units) cudaFree(p->d_units);
    if (p->d_unit_start) cudaFree(p->d_unit_start);
    delete p;
}

// nullptr when the box kernel does not apply (then the slice-class kernel runs)
EuBoxPlan* eu_box_plan_create(int nx, int ny, int nz, int z_lo, int z_hi, double* S0, double* S1, double* pc0, double* pc1,
                              double* qa, double* Ga, double* T, const unsigned short* cmask, const int* irr_cells, int n_irr, double* acc_irr,
                              double* inv_porevol, int n_sms)
{
    // TMA: global strides are multiples of 16 bytes (nx even); coordinates fit the unit encoding
    if (nx < 2 || (nx & 1) || ny < 1 || nz < 1 || nx > 65535 || ny > 32767) return nullptr;
    if (!encode_tiled_fn()) return nullptr;
    EuBoxPlan* p = new EuBoxPlan;
    p->nx = nx; p->ny = ny; p->nz = nz; p->z_lo = z_lo; p->z_hi = z_hi; p->n_local = nx*ny*nz;
    p->cmask = cmask; p->irr_cells = irr_cells; p->n_irr = n_irr; p->acc_irr = acc_irr; p->n_sms = n_sms;
    // tile: tx even, 2(tx+1) <= 256 (TMA box limit), tx*ty <= 256 threads; minimise (halo overhead) x (idle threads)
    {
        const char* e = getenv("EU_BOX_TILE");          // "<tx>x<ty>" (tuning knob)
        int etx = 0, ety = 0;
        if (e && std::sscanf(e, "%dx%d", &etx, &ety) == 2 && etx >= 2 && !(etx & 1) && etx <= 126 && ety >= 1 && etx*ety <= 256) {
            p->tx = etx; p->ty = ety;
        } else {
            double best = 1e300;
            for (int tx = 2; tx <= 126 && tx <= ((nx + 1) & ~1); tx += 2) {
                for (int ty = 1; ty <= 32 && tx*ty <= 256 && ty <= ny; ++ty) {
                    const double halo = double((tx + 2)*(ty + 2))/double(tx*ty);
                    const double idle_x = double(((nx + tx - 1)/tx)*tx)/nx, idle_y = double(((ny + ty - 1)/ty)*ty)/ny;
                    const double fill = double(tx*ty)/double((tx*ty + 31)/32*32);
                    // fitted to tile sweeps on B200 (C2 at 100^3 / 200^3, C4 at 512^2 x 256; profiles/README.md): the halo
                    // cells cost about a third of an own cell, short rows lose a little (row segments of the TMA boxes
                    // and of the ring reads), idle lanes cost everything
                    const int thr = (tx*ty + 31)/32*32;
                    const double cost = (1.0 + 0.35*(halo - 1.0))*(1.0 + 1.5/tx)*idle_x*idle_y/fill*(1.0 + 0.5*(256 - thr)/256.0);   // (small blocks: fewer warps per SM)
                    if (cost < best - 1e-12) { best = cost; p->tx = tx; p->ty = ty; }
                }
            }
        }
        p->threads = (p->tx*p->ty + 31)/32*32;
    }
    const cuuint64_t n = cuuint64_t(p->n_local);
    bool ok = true;
    {
        cuuint64_t dims[3] = { cuuint64_t(nx), cuuint64_t(ny), cuuint64_t(nz) };
        cuuint64_t str[2] = { cuuint64_t(nx)*8, cuuint64_t(nx)*ny*8 };
        cuuint32_t box[3] = { cuuint32_t(p->tx + 4), cuuint32_t(p->ty + 2), 1 };      // starts at x0 - 2: even coordinate
        ok = ok && make_map(&p->mapS[0], S0, 3, dims, str, box) && make_map(&p->mapS[1], S1, 3, dims, str, box);
        ok = ok && make_map(&p->mapPc[0], pc0, 3, dims, str, box) && make_map(&p->mapPc[1], pc1, 3, dims, str, box);
        // per-cell operands of the plane being finished: the tile itself (tx even: 16-byte rows at 16-byte offsets)
        cuuint32_t tile[3] = { cuuint32_t(p->tx), cuuint32_t(p->ty), 1 };
        ok = ok && make_map(&p->mapA, acc_irr, 3, dims, str, tile) && make_map(&p->mapV, inv_porevol, 3, dims, str, tile);
    }
    {
        // the three axis planes of q, G and T: [3][nz][ny][nx] doubles, boxes of (tx+2) x (ty+1) (x boxes start at x0 - 2)
        cuuint64_t dims[4] = { cuuint64_t(nx), cuuint64_t(ny), cuuint64_t(nz), 3 };
        cuuint64_t str[3] = { cuuint64_t(nx)*8, cuuint64_t(nx)*ny*8, n*8 };
        cuuint32_t box[4] = { cuuint32_t(p->tx + 2), cuuint32_t(p->ty + 1), 1, 1 };
        ok = ok && make_map(&p->mapQ, qa, 4, dims, str, box) && make_map(&p->mapG, Ga, 4, dims, str, box) && make_map(&p->mapT, T, 4, dims, str, box);
    }
    if (!ok) { eu_box_plan_destroy(p); return nullptr; }
    return p;
}

void eu_box_plan_set_gravity_mask(EuBoxPlan* p, int mask) { if (p) p->g_mask = mask & 7; }

void eu_box_plan_info(const EuBoxPlan* p, int out[6])
{
    out[0] = p->tx; out[1] = p->ty; out[2] = p->n_units; out[3] = p->n_bnd_units[0]; out[4] = p->n_bnd_units[1]; out[5] = p->threads;
}

// Work units of the sweep: the host logic is eu_box_make_units (eu_host.cpp; tested without a device through
// eu_debug_box_units).  EU_BOX_UNITS=spans / EU_BOX_LZ=<planes> (tuning knobs, read at every plan build) choose the other
// partition / a fixed chunk length.
static int box_build_units(EuBoxPlan* p, int bnd_lo, int bnd_hi, int grid_blocks, bool cap)
{
    const char* e = getenv("EU_BOX_LZ");
    const int lz_env = e ? atoi(e) : 0;
    const char* eu = getenv("EU_BOX_UNITS");
    const int mode = (eu && std::strcmp(eu, "spans") == 0) ? 1 : 0;
    const int lz_key = lz_env + 100000*mode;
    if (p->units_key[0] == bnd_lo && p->units_key[1] == bnd_hi && p->units_key[2] == grid_blocks && p->units_key[3] == lz_key &&
        p->units_key[4] == int(cap) && p->d_units) return 0;
    static_assert(sizeof(EuBoxUnit) == sizeof(int4), "unit records are uploaded as int4");
    std::vector<EuBoxUnit> units;       // block after block, a block's flagged units first
    std::vector<int> start;             // [blocks + 1]
    if (eu_box_make_units(p->nx, p->ny, p->tx, p->ty, p->z_lo, p->z_hi, bnd_lo, bnd_hi, grid_blocks, mode, lz_env, units, start) < 0) return -1;
    p->n_bnd_units[0] = p->n_bnd_units[1] = 0;
    p->n_flagged = 0;
    for (const EuBoxUnit& un : units) {
        if (un.flags) ++p->n_flagged;
        if (un.flags & 1) ++p->n_bnd_units[0];
        if (un.flags & 2) ++p->n_bnd_units[1];
    }
    if (p->d_units) { cudaFree(p->d_units); p->d_units = nullptr; }
    if (p->d_unit_start) { cudaFree(p->d_unit_start); p->d_unit_start = nullptr; }
    p->n_units = int(units.size());
    p->n_blocks = start.empty() ? 0 : int(start.size()) - 1;
    if (units.empty()) return 0;
    if (cudaMalloc((void**)&p->d_units, units.size()*sizeof(int4)) != cudaSuccess) return -1;
    if (cudaMemcpy(p->d_units, units.data(), units.size()*sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMalloc((void**)&p->d_unit_start, start.size()*sizeof(int)) != cudaSuccess) return -1;
    if (cudaMemcpy(p->d_unit_start, start.data(), start.size()*sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    p->units_key[0] = bnd_lo; p->units_key[1] = bnd_hi; p->units_key[2] = grid_blocks; p->units_key[3] = lz_key; p->units_key[4] = int(cap);
    return 0;
}

// number of flagged units the plan will have for these boundary planes (info[3], info[4]); -1 when the slab is thinner than
// the boundary ranges
int eu_box_plan_units(EuBoxPlan* p, int bnd_lo, int bnd_hi, bool, int info[6])
{
    if (std::max(bnd_lo, bnd_hi) > p->z_hi - p->z_lo) return -1;
    const int tiles = ((p->nx + p->tx - 1)/p->tx)*((p->ny + p->ty - 1)/p->ty);
    eu_box_plan_info(p, info);
    info[3] = bnd_lo > 0 ? tiles : 0;
    info[4] = bnd_hi > 0 ? tiles : 0;
    return 0;
}

template <bool ROCKS, bool MULTIROCK, bool CAP>
static int launch_box(EuBoxPlan* p, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                      const EuHaloDev& halo, int cur, int slice_lo, int slice_hi, int bnd_lo, int bnd_hi, cudaStream_t st)
{
    const size_t tab_bytes = eu_fast_smem_bytes(t);
    static int stages_env = -1;
    if (stages_env < 0) { const char* e = getenv("EU_BOX_STAGES"); stages_env = e ? std::min(std::max(atoi(e), 2), 4) : 0; }
    // as many bundles in flight (2..4) as still leave the kernel's resident blocks per SM (3, or 2 with the capillary term)
    static int minb_pre = -1;
    if (minb_pre < 0) { const char* e = getenv("EU_BOX_MINB"); minb_pre = e ? atoi(e) : 0; }
    int want_blocks = minb_pre ? minb_pre : (CAP ? 2 : 3);
    int stages = stages_env ? stages_env : 3;
    if (p->g_mask_seen != p->g_mask) { std::memset(p->smem_set, 0, sizeof(p->smem_set)); std::memset(p->blocks_per_sm, 0, sizeof(p->blocks_per_sm)); p->g_mask_seen = p->g_mask; p->units_key[0] = -1; p->four_blocks = -1; }
    // EU_BOX_SHARE (tuning knob): lateral faces evaluated once per tile and exchanged through shared memory; default: with
    // the capillary term
    static int share_env = -1;
    if (share_env < 0) { const char* e = getenv("EU_BOX_SHARE"); share_env = e ? (atoi(e) != 0 ? 1 : 0) : 2; }
    const bool share = CAP && (share_env == 2 ? true : share_env == 1);
    if (!CAP && !minb_pre && !stages_env) {
        // without the capillary term the kernel needs 64 registers (no or one rock table): FOUR blocks per SM are resident
        // if their shared memory fits, which it does with two stages -- measured 6 % faster than three blocks with three
        // stages (profiles/README.md, r04d)
        if (p->four_blocks < 0) {
            cudaFuncAttributes fa;
            const BoxLayout l2 = box_layout(*p, CAP, MULTIROCK, 2, tab_bytes, share);
            p->four_blocks = (cudaFuncGetAttributes(&fa, k_box_step<ROCKS, MULTIROCK, CAP, 2, 3, false>) == cudaSuccess &&
                              fa.numRegs*p->threads*4 <= 65536 && l2.total <= size_t(227*1024)/4 - 1024) ? 1 : 0;
        }
        if (p->four_blocks == 1) want_blocks = 4;
    }
    BoxLayout lay = box_layout(*p, CAP, MULTIROCK, stages, tab_bytes, share);
    const size_t budget = size_t(227*1024)/want_blocks - 1024;
    while (!stages_env && stages > 2 && lay.total > budget) lay = box_layout(*p, CAP, MULTIROCK, --stages, tab_bytes, share);
    // EU_BOX_MINB (tuning knob): resident blocks per SM the kernel is compiled for (register budget 2: 128, 3: 80)
    static int minb_env = -1;
    if (minb_env < 0) { const char* e = getenv("EU_BOX_MINB"); minb_env = e ? atoi(e) : 0; }
    const bool two = minb_env ? minb_env == 2 : CAP;
    auto kern3 = stages == 2 ? k_box_step<ROCKS, MULTIROCK, CAP, 2, 3, false> : (stages == 3 ? k_box_step<ROCKS, MULTIROCK, CAP, 3, 3, false> : k_box_step<ROCKS, MULTIROCK, CAP, 4, 3, false>);
    auto kern2 = stages == 2 ? k_box_step<ROCKS, MULTIROCK, CAP, 2, 2, false> : (stages == 3 ? k_box_step<ROCKS, MULTIROCK, CAP, 3, 2, false> : k_box_step<ROCKS, MULTIROCK, CAP, 4, 2, false>);
    auto kern2s = stages == 2 ? k_box_step<ROCKS, MULTIROCK, CAP, 2, 2, CAP> : k_box_step<ROCKS, MULTIROCK, CAP, 3, 2, CAP>;
    auto kern4 = stages == 2 ? k_box_step<ROCKS, MULTIROCK, CAP, 2, 4, false> : k_box_step<ROCKS, MULTIROCK, CAP, 3, 4, false>;
    auto kern = share ? kern2s : (minb_env == 4 ? kern4 : (two ? kern2 : kern3));
    // function attributes are per device: a plan remembers what it set on ITS device (several solvers, one per GPU, may
    // live in one process)
    const int vkey = (share ? 16 : 0) + (minb_env == 4 ? 8 : 0) + (two ? 4 : 0) + stages - 2;
    if (lay.total > p->smem_set[vkey]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total) != cudaSuccess) return -1;
        p->smem_set[vkey] = lay.total;
        p->blocks_per_sm[vkey] = 0;
    }
    if (p->blocks_per_sm[vkey] == 0) {
        int bps = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, p->threads, lay.total);
        p->blocks_per_sm[vkey] = bps;
    }
    const int blocks_per_sm = p->blocks_per_sm[vkey];
    if (blocks_per_sm < 1) return -1;
    const int grid_full = p->n_sms*blocks_per_sm;
    if (box_build_units(p, bnd_lo, bnd_hi, grid_full, CAP)) return -1;
    if (p->n_units == 0) return 0;
    lay.b.units = p->d_units;
    lay.b.n_units = p->n_units;
    lay.b.n_flagged = p->n_flagged;
    lay.b.unit_start = p->d_unit_start;
    int launches = 1;
    // EU_PDL (tuning knob, default 0): programmatic dependent launch -- a kernel's blocks may take their SM slots and run
    // their prologue (rock tables to shared memory, mbarrier set-up) while the previous kernel of the stream drains;
    // griddepcontrol.wait in the kernels orders every access to substep data behind the previous kernel's completion.
    // Measured (profiles/README.md, r04b): correct (all GPU tests pass with it) but SLOWER -- the sweep without the
    // capillary term 1.12 -> 1.69 ms per substep at 512x512x256, as if one of its three blocks per SM came too late
    // (early blocks of the next pre-pass hold registers); with the capillary term (two blocks per SM) no change.  Off.
    static int pdl = -1;
    if (pdl < 0) { const char* e = getenv("EU_PDL"); pdl = e ? (atoi(e) != 0) : 0; }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p->n_irr > 0) {
        // pre-pass: faces outside the axis planes, summed per cell
        const int ib = std::min((p->n_irr + 255)/256, p->n_sms*8);
        cfg.gridDim = dim3(ib);
        cfg.dynamicSmemBytes = tab_bytes;
        const int* irr_cells = p->irr_cells;
        const unsigned short* cmask = p->cmask;
        if (cudaLaunchKernelEx(&cfg, k_box_irregular<ROCKS, MULTIROCK, CAP>, g, t, f, a, halo, irr_cells, p->n_irr, cmask, p->acc_irr) != cudaSuccess) return -1;
        ++launches;
    }
    cfg.gridDim = dim3(p->n_blocks);
    cfg.blockDim = dim3(p->threads);
    cfg.dynamicSmemBytes = lay.total;
    if (cudaLaunchKernelEx(&cfg, kern, p->mapS[cur], p->mapPc[cur], p->mapQ, p->mapG, p->mapT, p->mapA, p->mapV, g, t, f, a, halo, lay.b, slice_lo, slice_hi, (int)tab_bytes) != cudaSuccess) return -1;
    return launches;
}

// one substep of the own planes [z_lo, z_hi) with the box kernel; returns the number of launches, -1 on error
int eu_launch_box_step(EuBoxPlan* p, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                       const EuHaloDev& halo, int cur, int slice_lo, int slice_hi, int bnd_lo, int bnd_hi, cudaStream_t st)
{
    const bool cap = a.method_capillary != 0;
    if (t.n_rocks > 1) {
        return cap ? launch_box<true, true, true>(p, g, t, f, a, halo, cur, slice_lo, slice_hi, bnd_lo, bnd_hi, st)
                   : launch_box<true, true, false>(p, g, t, f, a, halo, cur, slice_lo, slice_hi, bnd_lo, bnd_hi, st);
    } else if (t.n_rocks == 1) {
        return cap ? launch_box<true, false, true>(p, g, t, f, a, halo, cur, slice_lo, slice_hi, bnd_lo, bnd_hi, st)
                   : launch_box<true, false, false>(p, g, t, f, a, halo, cur, slice_lo, slice_hi, bnd_lo, bnd_hi, st);
    }
    return cap ? launch_box<false, false, true>(p, g, t, f, a, halo, cur, slice_lo, slice_hi, bnd_lo, bnd_hi, st)
               : launch_box<false, false, false>(p, g, t, f, a, halo, cur, slice_lo, slice_hi, bnd_lo, bnd_hi, st);
}

bool eu_fast_uses_stored_lam() { return kStoredLam; }

size_t eu_fast_smem_bytes(const EuTablesDev& t)
{
    if (t.n_rocks == 0) return 0;
    size_t b = size_t(56)*t.n_nodes_total + 4*(EU_MAX_TABLES + 2) + size_t(t.n_rocks)*t.n_buckets;
    return (b + 15) & ~size_t(15);
}

void eu_launch_fast_state(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const double* S, double* pc,
                          double2* lam, int lo, int hi, cudaStream_t st)
{
    if (hi <= lo || (!pc && !lam)) return;
    const size_t smem = eu_fast_smem_bytes(t);
    int blocks = (hi - lo + kBlock - 1)/kBlock;
    if (blocks > 148*8) blocks = 148*8;
    if (t.n_rocks > 1)       k_fast_state<true, true><<<blocks, kBlock, smem, st>>>(g, t, f, S, pc, lam, lo, hi);
    else if (t.n_rocks == 1) k_fast_state<true, false><<<blocks, kBlock, smem, st>>>(g, t, f, S, pc, lam, lo, hi);
    else                     k_fast_state<false, false><<<blocks, kBlock, 0, st>>>(g, t, f, S, pc, lam, lo, hi);
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int B6, int B8, int MINB, bool TENSOR = false>
static void launch_variant(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                           const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, size_t smem_tables, cudaStream_t st)
{
    // per device (function attributes are per device; several solvers, one per GPU, may live in one process)
    static int blocks_per_sm_dev[64] = { 0 };
    static size_t smem_seen_dev[64] = { 0 };
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    int& blocks_per_sm = blocks_per_sm_dev[dev];
    size_t& smem_seen = smem_seen_dev[dev];
    auto kern = k_fast_step<ROCKS, MULTIROCK, CAP, NN, B6, B8, MINB, TENSOR>;
    // shared memory: rock tables | slice classes
    const size_t smem = smem_tables + sizeof(EuSliceClass)*EU_MAX_CLASSES;
    if (blocks_per_sm == 0 || smem != smem_seen) {      // (another solver in this process may have larger tables)
        if (smem > 48*1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        smem_seen = smem;
    }
    // persistent grid: every SM full, capped by the work available
    int n = f.n_items;
    if (halo.enabled) n += (halo.a_hi - slice_lo) + (slice_hi - halo.b_lo);
    int blocks = n_sms*blocks_per_sm;
    const int need = (n + kWarpsPerBlock - 1)/kWarpsPerBlock;
    if (blocks > need) blocks = need;
    if (blocks < 1) return;
#ifdef EU_DYNAMIC_ITEMS
    {
        void* counter = nullptr;
        cudaGetSymbolAddress(&counter, g_item_counter);
        cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
    }
#endif
    kern<<<blocks, kBlock, smem, st>>>(g, t, f, a, halo, slice_lo, slice_hi, (int)smem_tables);
}

// EU_FAST_VARIANT (tuning knob, read once): resident blocks per SM the kernel is compiled for.
//   0 (default)  by variant: 2 blocks (128 registers) with the capillary term -- at 80 registers that variant spills
//                128-1400 bytes per thread (-Xptxas -v) --, 3 blocks (80 registers) without
//   1 / 2 / 3    force 2 / 4 / 3 blocks per SM
static int fast_variant()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("EU_FAST_VARIANT"); v = e ? atoi(e) : 0; }
    return v;
}

// resident warps per SM of the variant launch_fast picks (8 warps per block): what the host's work-item model needs
int eu_fast_warps_per_sm(bool capillary)
{
    int v = fast_variant();
    if (v == 0) v = capillary ? 1 : 3;
    return kWarpsPerBlock*(v == 1 ? 2 : (v == 2 ? 4 : 3));
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN>
static void launch_fast(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                        const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, size_t smem, cudaStream_t st)
{
    int v = fast_variant();
    if (v == 0) v = CAP ? 1 : 3;
    switch (v) {
    case 1:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 3, 4, 2>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    case 2:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 3, 4, 4>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    default: launch_variant<ROCKS, MULTIROCK, CAP, NN, 3, 4, 3>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    }
}

template <bool ROCKS, bool MULTIROCK>
static void launch_fast2(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                         const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, size_t smem, cudaStream_t st)
{
    const bool cap = a.method_capillary != 0;
    const bool nn = f.nn != nullptr;
    if (ROCKS && MULTIROCK && f.axis8 != nullptr) {
        // diagonal tensor mobility: generic gather only, 2 blocks per SM (128 registers)
        if (cap && nn)   launch_variant<ROCKS, MULTIROCK, true, true, 3, 4, 2, ROCKS && MULTIROCK>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
        else if (cap)    launch_variant<ROCKS, MULTIROCK, true, false, 3, 4, 2, ROCKS && MULTIROCK>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
        else if (nn)     launch_variant<ROCKS, MULTIROCK, false, true, 3, 4, 2, ROCKS && MULTIROCK>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
        else             launch_variant<ROCKS, MULTIROCK, false, false, 3, 4, 2, ROCKS && MULTIROCK>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
        return;
    }
    if (cap && nn)       launch_fast<ROCKS, MULTIROCK, true, true>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else if (cap)        launch_fast<ROCKS, MULTIROCK, true, false>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else if (nn)         launch_fast<ROCKS, MULTIROCK, false, true>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else                 launch_fast<ROCKS, MULTIROCK, false, false>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
}

void eu_launch_fast_step_t3(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                            int slice_lo, int slice_hi, int n_sms, cudaStream_t st)
{
    if (slice_hi <= slice_lo) return;
    const size_t smem = eu_fast_smem_bytes(t);
    int blocks = n_sms*2;
    const int need = (slice_hi - slice_lo + kWarpsPerBlock - 1)/kWarpsPerBlock;
    if (blocks > need) blocks = need;
    if (a.method_capillary) {
        if (smem > 48*1024) cudaFuncSetAttribute(k_fast_step_t3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_fast_step_t3<true><<<blocks, kBlock, smem, st>>>(g, t, f, a, slice_lo, slice_hi);
    } else {
        if (smem > 48*1024) cudaFuncSetAttribute(k_fast_step_t3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_fast_step_t3<false><<<blocks, kBlock, smem, st>>>(g, t, f, a, slice_lo, slice_hi);
    }
}

void eu_launch_fast_step(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                         const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, cudaStream_t st)
{
    if (slice_hi <= slice_lo) return;
    const size_t smem = eu_fast_smem_bytes(t);
    if (t.n_rocks > 1)       launch_fast2<true, true>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else if (t.n_rocks == 1) launch_fast2<true, false>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else                     launch_fast2<false, false>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
}
