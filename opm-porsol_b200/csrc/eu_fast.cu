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
#include "eu_internal.h"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlock = kWarpsPerBlock*32;

struct SmemTables {
    const int* offset;        // n_rocks+1
    const double* x;          // nodes
    const double* lam[2];
    const double* lams[2];
    const double* J;
    const double* Js;
    const unsigned char* bucket;
};

__device__ __forceinline__ size_t align8(size_t v) { return (v + 7) & ~size_t(7); }

__device__ __forceinline__ SmemTables smem_tables_load(const EuTablesDev& t, unsigned char* smem)
{
    SmemTables s;
    const int nn = t.n_nodes_total;
    double* d = reinterpret_cast<double*>(smem);
    double* x = d;           d += nn;
    double* l0 = d;          d += nn;
    double* l1 = d;          d += nn;
    double* s0 = d;          d += nn;
    double* s1 = d;          d += nn;
    double* J = d;           d += nn;
    double* Js = d;          d += nn;
    int* off = reinterpret_cast<int*>(d);
    unsigned char* bucket = reinterpret_cast<unsigned char*>(off + (t.n_rocks + 1 + 1)/2*2);
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        x[i] = t.s[i];
        l0[i] = t.lam[0][i];  l1[i] = t.lam[1][i];
        s0[i] = t.lam_slope[0][i];  s1[i] = t.lam_slope[1][i];
        J[i] = t.J[i];  Js[i] = t.J_slope[i];
    }
    for (int i = threadIdx.x; i <= t.n_rocks; i += blockDim.x) off[i] = t.offset[i];
    for (int i = threadIdx.x; i < t.n_rocks*EU_BUCKETS; i += blockDim.x) bucket[i] = (unsigned char)t.bucket[i];
    s.offset = off; s.x = x; s.lam[0] = l0; s.lam[1] = l1; s.lams[0] = s0; s.lams[1] = s1; s.J = J; s.Js = Js;
    s.bucket = bucket;
    return s;
}

// interval of the rock table containing sat: same answer as the reference's binary search
// (first/last interval outside the table), found from a 64-bucket index plus a short scan.
__device__ __forceinline__ int interval(const SmemTables& tb, int rock, double sat)
{
    const int b = tb.offset[rock];
    const int last = tb.offset[rock + 1] - b - 2;        // index of the last interval
    int k = __double2int_rd(sat*EU_BUCKETS);
    k = min(max(k, 0), EU_BUCKETS - 1);
    int j = tb.bucket[rock*EU_BUCKETS + k];
    while (j < last && sat >= tb.x[b + j + 1]) ++j;
    return b + j;
}

template <bool ROCKS>
struct Mob {
    // mobilities of both phases at (rock, sat)
    static __device__ __forceinline__ void both(const SmemTables& tb, const EuTablesDev& t, int rock, double sat,
                                                double& lw, double& lo)
    {
        if (ROCKS) {
            const int j = interval(tb, rock, sat);
            const double ds = sat - tb.x[j];
            lw = fma(tb.lams[0][j], ds, tb.lam[0][j]);
            lo = fma(tb.lams[1][j], ds, tb.lam[1][j]);
        } else {
            lw = sat*sat/t.visc[0];
            lo = (1.0 - sat)*(1.0 - sat)/t.visc[1];
        }
    }
    static __device__ __forceinline__ double one(const SmemTables& tb, const EuTablesDev& t, int phase, int rock, double sat)
    {
        if (ROCKS) {
            const int j = interval(tb, rock, sat);
            return fma(tb.lams[phase][j], sat - tb.x[j], tb.lam[phase][j]);
        } else {
            return phase == 0 ? sat*sat/t.visc[0] : (1.0 - sat)*(1.0 - sat)/t.visc[1];
        }
    }
    static __device__ __forceinline__ double pc(const SmemTables& tb, int rock, double sat, double scale)
    {
        if (ROCKS) {
            const int j = interval(tb, rock, sat);
            return fma(tb.Js[j], sat - tb.x[j], tb.J[j])*scale;
        } else {
            return 1e5*(1.0 - sat);
        }
    }
};

template <bool ROCKS, bool CAP>
__global__ void __launch_bounds__(kBlock) k_fast_step(EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a,
                                                      int slice_lo, int slice_hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTables tb = {};
    if (ROCKS) {
        tb = smem_tables_load(t, smem_raw);
        __syncthreads();
    }
    {
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x*kWarpsPerBlock + (threadIdx.x >> 5);
    const int n_warps = gridDim.x*kWarpsPerBlock;

    for (int s = slice_lo + warp_global; s < slice_hi; s += n_warps) {
        const int c = s*EU_SLICE + lane;
        const bool active = (c >= g.own_lo) && (c < g.own_hi);
        const int base = f.slice_base[s];
        const int width = (f.slice_base[s + 1] - base) >> 5;
        double S0 = 0.0, pc0 = 0.0;
        int rock0 = 0;
        if (active) {
            S0 = a.S_in[c];
            if (ROCKS) rock0 = f.rock8[c];
            if (CAP) pc0 = a.pc_in[c];
        }
        double acc = 0.0;
        const int2* __restrict__ recp = f.rec + base + lane;
        for (int j = 0; j < width; ++j) {
            const int2 r = recp[j*EU_SLICE];
            if (!active || r.x == EU_REC_PAD) continue;
            const double q = f.q[r.y];
            const double G = f.G[r.y];
            double S1, pc1 = 0.0;
            int rock1 = rock0;
            bool own = true, interior = true;
            if (r.x >= 0) {
                S1 = a.S_in[r.x];
                if (ROCKS) rock1 = f.rock8[r.x];
                if (CAP) pc1 = a.pc_in[r.x];
                own = c < r.x;
            } else {
                S1 = g.bnd_sat[-2 - r.x];
                interior = false;
            }
            // (lo, hi) ordering
            const double S_lo = own ? S0 : S1, S_hi = own ? S1 : S0;
            const int r_lo = own ? rock0 : rock1, r_hi = own ? rock1 : rock0;
            const bool triv_w = G >= 0.0;
            const bool u_lo = q >= 0.0;
            const double lam_t = Mob<ROCKS>::one(tb, t, triv_w ? 0 : 1, u_lo ? r_lo : r_hi, u_lo ? S_lo : S_hi);
            const double gfn = (triv_w ? -1.0 : 1.0)*(lam_t*G);
            const bool u2_lo = (q + gfn) >= 0.0;
            const double lam_n = Mob<ROCKS>::one(tb, t, triv_w ? 1 : 0, u2_lo ? r_lo : r_hi, u2_lo ? S_lo : S_hi);
            const double lw = triv_w ? lam_t : lam_n;
            const double lo = triv_w ? lam_n : lam_t;
            const double inv = 1.0/(lw + lo);
            double dS = 0.0;
            if (a.method_viscous) {
                const double qq = f.nn ? q*f.nn[r.y] : q;
                dS += lw*(inv*qq);
            }
            if (a.method_gravity && interior) dS += lw*(inv*(lo*G));
            if (CAP && interior) {
                const double Sa = 0.5*(S_lo + S_hi);
                double lwa, loa;
                Mob<ROCKS>::both(tb, t, r_lo, Sa, lwa, loa);
                if (ROCKS && r_hi != r_lo) {
                    double lwb, lob;
                    Mob<ROCKS>::both(tb, t, r_hi, Sa, lwb, lob);
                    lwa = 0.5*(lwa + lwb);
                    loa = 0.5*(loa + lob);
                }
                const double inva = 1.0/(lwa + loa);
                const double pc_lo = own ? pc0 : pc1, pc_hi = own ? pc1 : pc0;
                dS += lwa*(inva*(loa*(f.T[r.y]*(pc_hi - pc_lo))));
            }
            acc += own ? -dS : dS;
        }
        if (active) {
            double rate = 0.0;
            if (a.n_src > 0) {
                int lo = 0, hi = a.n_src;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.src_cell[mid] < c) lo = mid + 1; else hi = mid; }
                if (lo < a.n_src && a.src_cell[lo] == c) rate = a.src_rate[lo];
                if (rate < 0.0) {
                    double lw, lo_;
                    Mob<ROCKS>::both(tb, t, rock0, S0, lw, lo_);
                    rate *= lw/(lw + lo_);
                }
            }
            acc += rate;
            if (a.residual_out) a.residual_out[c] = acc;
            double sat = S0 + a.dt*acc/f.porevol[c];
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
            if (CAP) a.pc_out[c] = Mob<ROCKS>::pc(tb, rock0, sat, ROCKS ? f.pcscale[c] : 1.0);
        }
    }
}

template <bool ROCKS>
__global__ void __launch_bounds__(kBlock) k_fast_pc(EuGridDev g, EuTablesDev t, EuFastDev f,
                                                    const double* __restrict__ S, double* __restrict__ pc, int lo, int hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTables tb = {};
    if (ROCKS) {
        tb = smem_tables_load(t, smem_raw);
        __syncthreads();
    }
    for (int c = lo + blockIdx.x*blockDim.x + threadIdx.x; c < hi; c += gridDim.x*blockDim.x) {
        pc[c] = Mob<ROCKS>::pc(tb, ROCKS ? f.rock8[c] : 0, S[c], ROCKS ? f.pcscale[c] : 1.0);
    }
}

} // namespace

size_t eu_fast_smem_bytes(const EuTablesDev& t)
{
    if (t.n_rocks == 0) return 0;
    size_t b = size_t(7)*t.n_nodes_total*sizeof(double);
    b += size_t((t.n_rocks + 1 + 1)/2*2)*sizeof(int);
    b += size_t(t.n_rocks)*EU_BUCKETS;
    return (b + 15) & ~size_t(15);
}

void eu_launch_fast_pc(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const double* S, double* pc,
                       int lo, int hi, cudaStream_t st)
{
    if (hi <= lo) return;
    const size_t smem = eu_fast_smem_bytes(t);
    int blocks = (hi - lo + kBlock - 1)/kBlock;
    if (blocks > 148*8) blocks = 148*8;
    if (t.n_rocks > 0) k_fast_pc<true><<<blocks, kBlock, smem, st>>>(g, t, f, S, pc, lo, hi);
    else               k_fast_pc<false><<<blocks, kBlock, 0, st>>>(g, t, f, S, pc, lo, hi);
}

void eu_launch_fast_step(const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f, const EuStepArgs& a,
                         int slice_lo, int slice_hi, int n_sms, cudaStream_t st)
{
    const int n = slice_hi - slice_lo;
    if (n <= 0) return;
    const size_t smem = eu_fast_smem_bytes(t);
    static bool attr_set = false;
    if (!attr_set && smem > 48*1024) {
        cudaFuncSetAttribute(k_fast_step<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_fast_step<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    // persistent grid: a multiple of the SM count, capped by the work available
    int blocks = n_sms*4;
    const int need = (n + kWarpsPerBlock - 1)/kWarpsPerBlock;
    if (blocks > need) blocks = need;
    const bool rocks = t.n_rocks > 0;
    const bool cap = a.method_capillary != 0;
    if (rocks && cap)        k_fast_step<true, true><<<blocks, kBlock, smem, st>>>(g, t, f, a, slice_lo, slice_hi);
    else if (rocks)          k_fast_step<true, false><<<blocks, kBlock, smem, st>>>(g, t, f, a, slice_lo, slice_hi);
    else if (cap)            k_fast_step<false, true><<<blocks, kBlock, 0, st>>>(g, t, f, a, slice_lo, slice_hi);
    else                     k_fast_step<false, false><<<blocks, kBlock, 0, st>>>(g, t, f, a, slice_lo, slice_hi);
}
