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

#include <cstdlib>

namespace {

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
    int shift;       // unused
    __device__ __forceinline__ const double4* coef() const { return reinterpret_cast<const double4*>(eu_smem); }
    __device__ __forceinline__ const double2* jcoef() const { return reinterpret_cast<const double2*>(eu_smem + size_t(32)*nn); }
    __device__ __forceinline__ const double* xb() const { return reinterpret_cast<const double*>(eu_smem + size_t(48)*nn); }
    __device__ __forceinline__ const int* offset() const { return reinterpret_cast<const int*>(eu_smem + size_t(56)*nn); }
    __device__ __forceinline__ const unsigned char* bucket() const { return eu_smem + size_t(56)*nn + 4*(EU_MAX_ROCKS + 2); }
};

__device__ __forceinline__ void tables_to_smem(const EuTablesDev& t)
{
    const int nn = t.n_nodes_total;
    double4* coef = reinterpret_cast<double4*>(eu_smem);
    double2* jc = reinterpret_cast<double2*>(eu_smem + size_t(32)*nn);
    double* xb = reinterpret_cast<double*>(eu_smem + size_t(48)*nn);
    int* off = reinterpret_cast<int*>(eu_smem + size_t(56)*nn);
    unsigned char* bucket = eu_smem + size_t(56)*nn + 4*(EU_MAX_ROCKS + 2);
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
    int k = __double2int_rd(sat*double(L.nb));
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

// x/d for a positive, normal-range denominator (sum of two mobilities): reciprocal seed + two Newton
// steps + one residual correction; no special-case branch.  Relative error <= 2^-52.
__device__ __forceinline__ double div_pos(double x, double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    double qv = x*r;
    const double rem = fma(-d, qv, x);
    return fma(rem, r, qv);
}

// One face of the gather, from the point of view of cell "self" (mobilities lw0/lo0) against the cell or
// boundary value on the other side (lw1/lo1).  own: self is the lower-index ("lo") cell, whose frame q, G
// and T are expressed in.  Returns the contribution to residual[self].
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
    const double gfn = triv_w ? -(lam_t*G) : lam_t*G;
    const bool u2_self = ((q + gfn) >= 0.0) == own;
    const double lam_n = u2_self ? n0 : n1;
    const double lw = triv_w ? lam_t : lam_n;
    const double lo = triv_w ? lam_n : lam_t;
    double num = method_viscous ? qq : 0.0;
    num = (method_gravity && interior) ? fma(lo, G, num) : num;
    double dS = div_pos(lw*num, lam_t + lam_n);
    if (CAP) dS = interior ? fma(cap_coef, Tdpc, dS) : dS;
    return own ? -dS : dS;
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int W, int B>
__device__ __forceinline__ double gather_cell(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t,
                                              const EuFastDev& f, const EuStepArgs& a, const int2* __restrict__ recp,
                                              const int2* __restrict__ dscp, int width, int c, double S0, int rock0, double pc0)
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
    double lw0, lo0;
    Mob<ROCKS, MULTIROCK>::both(L, t, rock0, S0, lw0, lo0);
    double acc = 0.0;
#pragma unroll
    for (int j0 = 0; j0 < W; j0 += B) {
        // phase B: every gather of the batch in flight at once
        double S1[B], q[B], G[B], T[CAP ? B : 1], pc1[CAP ? B : 1], nn[NN ? B : 1];
        int rk[MULTIROCK ? B : 1];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int j = j0 + k;
            S1[k] = 0.0; q[k] = 0.0; G[k] = 0.0;
            if (CAP) { T[k] = 0.0; pc1[k] = 0.0; }
            if (NN) nn[k] = 1.0;
            if (MULTIROCK) rk[k] = rock0;
            if (j < W && r[j].x != EU_REC_PAD) {
                q[k] = __ldg(f.q + r[j].y);
                G[k] = __ldg(f.G + r[j].y);
                if (NN) nn[k] = __ldg(f.nn + r[j].y);
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
            const int rk1 = MULTIROCK ? rk[k] : 0;
            double lw1, lo1;
            Mob<ROCKS, MULTIROCK>::both(L, t, rk1, S1[k], lw1, lo1);
            double cap_coef = 0.0, Tdpc = 0.0;
            if (CAP) {
                const double Sa = 0.5*(S0 + S1[k]);
                double lwa, loa;
                Mob<ROCKS, MULTIROCK>::both(L, t, rock0, Sa, lwa, loa);
                if (MULTIROCK && rk1 != rock0) {
                    double lwb, lob;
                    Mob<ROCKS, MULTIROCK>::both(L, t, rk1, Sa, lwb, lob);
                    lwa = 0.5*(lwa + lwb);
                    loa = 0.5*(loa + lob);
                }
                cap_coef = div_pos(lwa*loa, lwa + loa);
                Tdpc = T[k]*(own ? (pc1[k] - pc0) : (pc0 - pc1[k]));
            }
            const double contrib = face_contribution<CAP>(own, interior, q[k], NN ? q[k]*nn[k] : q[k], G[k], lw0, lo0, lw1, lo1,
                                                          a.method_viscous, a.method_gravity, cap_coef, Tdpc);
            acc += valid ? contrib : 0.0;
        }
    }
    return acc;
}

// generic width (cells with more than 8 faces): same arithmetic, one face at a time
template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN>
__device__ double gather_cell_loop(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t,
                                   const EuFastDev& f, const EuStepArgs& a, const int2* __restrict__ recp,
                                   int width, int c, double S0, int rock0, double pc0)
{
    double lw0, lo0;
    Mob<ROCKS, MULTIROCK>::both(L, t, rock0, S0, lw0, lo0);
    double acc = 0.0;
    for (int j = 0; j < width; ++j) {
        const int2 r = recp[j*EU_SLICE];
        if (r.x == EU_REC_PAD) continue;
        const bool interior = r.x >= 0;
        const bool own = !interior || c < r.x;
        const double q = f.q[r.y], G = f.G[r.y];
        const double nn = NN ? f.nn[r.y] : 1.0;
        const double S1 = interior ? a.S_in[r.x] : g.bnd_sat[-2 - r.x];
        const int rk = (MULTIROCK && interior) ? f.rock8[r.x] : rock0;
        double lw1, lo1;
        Mob<ROCKS, MULTIROCK>::both(L, t, rk, S1, lw1, lo1);
        double cap_coef = 0.0, Tdpc = 0.0;
        if (CAP && interior) {
            const double Sa = 0.5*(S0 + S1);
            double lwa, loa;
            Mob<ROCKS, MULTIROCK>::both(L, t, rock0, Sa, lwa, loa);
            if (MULTIROCK && rk != rock0) {
                double lwb, lob;
                Mob<ROCKS, MULTIROCK>::both(L, t, rk, Sa, lwb, lob);
                lwa = 0.5*(lwa + lwb);
                loa = 0.5*(loa + lob);
            }
            cap_coef = div_pos(lwa*loa, lwa + loa);
            const double pc1 = a.pc_in[r.x];
            Tdpc = f.T[r.y]*(own ? (pc1 - pc0) : (pc0 - pc1));
        }
        acc += face_contribution<CAP>(own, interior, q, q*nn, G, lw0, lo0, lw1, lo1, a.method_viscous, a.method_gravity,
                                      cap_coef, Tdpc);
    }
    return acc;
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int B6, int B8, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_fast_step(EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a,
                                                            EuHaloDev halo, int slice_lo, int slice_hi)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.shift = 0;
    if (ROCKS) {
        tables_to_smem(t);
        __syncthreads();
    }
    if (!halo.enabled) {   // (with a halo exchange every rank keeps stepping so that the flags stay in lockstep)
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x*kWarpsPerBlock + (threadIdx.x >> 5);
    const int n_warps = gridDim.x*kWarpsPerBlock;
    // processing order: boundary range A, boundary range B, then the interior
    const int nA = halo.enabled ? halo.a_hi - slice_lo : 0;
    const int nB = halo.enabled ? slice_hi - halo.b_lo : 0;
    bool waited = false;

    for (int v = warp_global; v < slice_hi - slice_lo; v += n_warps) {
        int s, range = -1;
        if (v < nA)           { s = slice_lo + v; range = 0; }
        else if (v < nA + nB) { s = halo.b_lo + (v - nA); range = 1; }
        else                  { s = slice_lo + nA + (v - nA - nB); }
        if (range >= 0 && !waited) {
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
            waited = true;
        }
        const int c = s*EU_SLICE + lane;
        const bool active = (c >= g.own_lo) && (c < g.own_hi);
        const int base = f.slice_base[s];
        const int width = (f.slice_base[s + 1] - base) >> 5;
        if (!active && range < 0) continue;
        if (active) {
        const double S0 = a.S_in[c];
        const int rock0 = MULTIROCK ? f.rock8[c] : 0;
        const double pc0 = CAP ? a.pc_in[c] : 0.0;
        const double inv_pv = f.inv_porevol[c];
        const int2* __restrict__ recp = f.rec + base + lane;
        const int2* __restrict__ dscp = f.desc + (base >> 5);
        double acc;
        if (width <= 6)      acc = gather_cell<ROCKS, MULTIROCK, CAP, NN, 6, B6>(L, g, t, f, a, recp, dscp, width, c, S0, rock0, pc0);
        else if (width <= 8) acc = gather_cell<ROCKS, MULTIROCK, CAP, NN, 8, B8>(L, g, t, f, a, recp, dscp, width, c, S0, rock0, pc0);
        else                 acc = gather_cell_loop<ROCKS, MULTIROCK, CAP, NN>(L, g, t, f, a, recp, width, c, S0, rock0, pc0);

        double rate = 0.0;
        if (a.n_src > 0) {
            int lo = 0, hi = a.n_src;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.src_cell[mid] < c) lo = mid + 1; else hi = mid; }
            if (lo < a.n_src && a.src_cell[lo] == c) rate = a.src_rate[lo];
            if (rate < 0.0) {
                double lw, lo_;
                Mob<ROCKS, MULTIROCK>::both(L, t, rock0, S0, lw, lo_);
                rate *= lw/(lw + lo_);
            }
        }
        acc += rate;
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
        double pcn = 0.0;
        if (CAP) { pcn = Mob<ROCKS, MULTIROCK>::pc(L, rock0, sat, ROCKS ? f.pcscale[c] : 1.0); a.pc_out[c] = pcn; }
        if (range >= 0) {
            // this cell is a ghost of the neighbour rank: store it into the neighbour's HBM as well
            const int first = (range == 0 ? slice_lo : halo.b_lo)*EU_SLICE;
            const int d = halo.dst[range][c - first];
            if (d >= 0) {
                halo.peer_S[range][d] = sat;
                if (CAP && halo.peer_pc[range]) halo.peer_pc[range][d] = pcn;
            }
            __threadfence_system();
        }
        }   // active
        if (range >= 0) {
            __syncwarp();
            if (lane == 0) {
                const unsigned done = atomicAdd(halo.counter[range], 1u);
                if (done == halo.total[range] - 1u) {
                    *halo.counter[range] = 0u;
                    __threadfence_system();
                    *(volatile unsigned*)halo.peer_flag[range] = halo.epoch;
                    __threadfence_system();
                }
            }
        }
    }
}

template <bool ROCKS, bool MULTIROCK>
__global__ void __launch_bounds__(kBlock) k_fast_pc(EuGridDev g, EuTablesDev t, EuFastDev f,
                                                    const double* __restrict__ S, double* __restrict__ pc, int lo, int hi)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.shift = 0;
    if (ROCKS) {
        tables_to_smem(t);
        __syncthreads();
    }
    for (int c = lo + blockIdx.x*blockDim.x + threadIdx.x; c < hi; c += gridDim.x*blockDim.x) {
        pc[c] = Mob<ROCKS, MULTIROCK>::pc(L, MULTIROCK ? f.rock8[c] : 0, S[c], ROCKS ? f.pcscale[c] : 1.0);
    }
}

} // namespace

size_t eu_fast_smem_bytes(const EuTablesDev& t)
{
    if (t.n_rocks == 0) return 0;
    size_t b = size_t(56)*t.n_nodes_total + 4*(EU_MAX_ROCKS + 2) + size_t(t.n_rocks)*t.n_buckets;
    return (b + 15) & ~size_t(15);
}

void eu_launch_fast_pc(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const double* S, double* pc,
                       int lo, int hi, cudaStream_t st)
{
    if (hi <= lo) return;
    const size_t smem = eu_fast_smem_bytes(t);
    int blocks = (hi - lo + kBlock - 1)/kBlock;
    if (blocks > 148*8) blocks = 148*8;
    if (t.n_rocks > 1)       k_fast_pc<true, true><<<blocks, kBlock, smem, st>>>(g, t, f, S, pc, lo, hi);
    else if (t.n_rocks == 1) k_fast_pc<true, false><<<blocks, kBlock, smem, st>>>(g, t, f, S, pc, lo, hi);
    else                     k_fast_pc<false, false><<<blocks, kBlock, 0, st>>>(g, t, f, S, pc, lo, hi);
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN, int B6, int B8, int MINB>
static void launch_variant(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                           const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, size_t smem, cudaStream_t st)
{
    static int blocks_per_sm = 0;
    auto kern = k_fast_step<ROCKS, MULTIROCK, CAP, NN, B6, B8, MINB>;
    if (blocks_per_sm == 0) {
        if (smem > 48*1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBlock, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    // persistent grid: every SM full, capped by the work available
    const int n = slice_hi - slice_lo;
    int blocks = n_sms*blocks_per_sm;
    const int need = (n + kWarpsPerBlock - 1)/kWarpsPerBlock;
    if (blocks > need) blocks = need;
    kern<<<blocks, kBlock, smem, st>>>(g, t, f, a, halo, slice_lo, slice_hi);
}

// EU_FAST_VARIANT (tuning knob, read once): faces per load batch / resident blocks per SM
static int fast_variant()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("EU_FAST_VARIANT"); v = e ? atoi(e) : 0; }
    return v;
}

template <bool ROCKS, bool MULTIROCK, bool CAP, bool NN>
static void launch_fast(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                        const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, size_t smem, cudaStream_t st)
{
    switch (fast_variant()) {
    case 1:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 3, 4, 5>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    case 2:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 2, 2, 5>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    case 3:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 2, 2, 6>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    case 4:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 1, 1, 6>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    case 5:  launch_variant<ROCKS, MULTIROCK, CAP, NN, 1, 1, 8>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    default: launch_variant<ROCKS, MULTIROCK, CAP, NN, 3, 4, 4>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st); break;
    }
}

template <bool ROCKS, bool MULTIROCK>
static void launch_fast2(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                         const EuHaloDev& halo, int slice_lo, int slice_hi, int n_sms, size_t smem, cudaStream_t st)
{
    const bool cap = a.method_capillary != 0;
    const bool nn = f.nn != nullptr;
    if (cap && nn)       launch_fast<ROCKS, MULTIROCK, true, true>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else if (cap)        launch_fast<ROCKS, MULTIROCK, true, false>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else if (nn)         launch_fast<ROCKS, MULTIROCK, false, true>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
    else                 launch_fast<ROCKS, MULTIROCK, false, false>(g, t, f, a, halo, slice_lo, slice_hi, n_sms, smem, st);
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
