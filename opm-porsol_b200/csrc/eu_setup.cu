// Structure building, CFL reductions and the STRICT substep kernels.
// Compile with -fmad=false: everything numeric here follows the reference's operation order
// (see eu_strict_math.cuh) and is bit-identical to it.
#include "eu_internal.h"
#include <algorithm>

#include <climits>
#include "eu_strict_math.cuh"

namespace {

constexpr int kThreads = 256;

inline int div_up(long long a, int b) { return int((a + b - 1)/b); }

// ---------------------------------------------------------------------------------------
// neighbour translation: global cell id -> local cell id over the uploaded ranges
// ---------------------------------------------------------------------------------------
__global__ void k_translate_nbr(int* __restrict__ hf_nbr, long long H, const int* __restrict__ rfirst,
                                const int* __restrict__ rcount, const int* __restrict__ rlocal, int n_ranges)
{
    long long h = blockIdx.x*(long long)blockDim.x + threadIdx.x;
    if (h >= H) return;
    int v = hf_nbr[h];
    if (v < 0) return;
    int out = -1;                      // neighbour not held by this rank (allowed for ghost cells only)
    for (int r = 0; r < n_ranges; ++r) {
        int d = v - rfirst[r];
        if (d >= 0 && d < rcount[r]) { out = rlocal[r] + d; break; }
    }
    hf_nbr[h] = out;
}

// ---------------------------------------------------------------------------------------
// owner half-face of every half-face: the half-face of the lower-index cell of the pair
// (euler/EulerUpstreamResidual_impl.hpp:125-128,138-141: a face is evaluated once, from
// the lower-index cell, with that cell's area/normal/flux, :146-148)
// ---------------------------------------------------------------------------------------
__global__ void k_owner(EuGridDev g, int* __restrict__ owner_hf, int* __restrict__ err_flag)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
    const bool own_cell = (c >= g.own_lo && c < g.own_hi);
    for (int h = b; h < e; ++h) {
        const int n = g.hf_nbr[h];
        int owner = -1;
        if (n >= 0) {
            if (c < n) {
                owner = h;
            } else {
                int m = 0;                                  // which of my faces towards n is this?
                for (int k = b; k < h; ++k) m += (g.hf_nbr[k] == n);
                const int nb = g.hf_offset[n], ne = g.hf_offset[n + 1];
                for (int k = nb; k < ne; ++k) {
                    if (g.hf_nbr[k] == c) {
                        if (m == 0) { owner = k; break; }
                        --m;
                    }
                }
                if (owner < 0 && own_cell) atomicExch(err_flag, 1);   // asymmetric connectivity
            }
        } else if (n <= -2) {
            const int bi = -2 - n;
            if (g.bnd_kind[bi] == EU_HF_PERIODIC) {
                const int pc = g.bnd_partner_cell[bi];
                if (pc < 0) {
                    if (own_cell) atomicExch(err_flag, 2);            // partner cell not uploaded
                } else {
                    owner = (c < pc) ? h : g.bnd_partner_hf[bi];
                }
            } else {
                owner = h;
            }
        } else if (own_cell) {
            atomicExch(err_flag, 3);                                  // neighbour of an own cell missing
        }
        owner_hf[h] = owner;
    }
}

__device__ __forceinline__ int hf_other_cell(const EuGridDev& g, int c, int h)
{
    const int n = g.hf_nbr[h];
    if (n >= 0) return n;
    if (n <= -2) {
        const int bi = -2 - n;
        if (g.bnd_kind[bi] == EU_HF_PERIODIC) return g.bnd_partner_cell[bi];
        return c;                                                    // Dirichlet: cell[1] = cell[0]
    }
    return -1;
}

// ---------------------------------------------------------------------------------------
// STRICT accumulation list.  residual[c] receives, in the serial reference, first "+= dS"
// from every face owned by a lower-index cell (in order of that cell, then its local face
// order == ascending owner half-face index), then "-= dS" for its own faces in local order
// (euler/EulerUpstreamResidual_impl.hpp:284-290 executed in the cell loop :496-503).
// ---------------------------------------------------------------------------------------
__global__ void k_strict_list(EuGridDev g, const int* __restrict__ owner_hf, int2* __restrict__ list)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
    int n_in = 0;
    for (int h = b; h < e; ++h) {
        const int o = owner_hf[h];
        if (o != h) {
            // insertion sort by owner half-face index
            int pos = b + n_in;
            const int lo = (o >= 0) ? hf_other_cell(g, c, h) : -1;
            while (pos > b && list[pos - 1].x > o) { list[pos] = list[pos - 1]; --pos; }
            list[pos] = make_int2(o, lo);
            ++n_in;
        }
    }
    int pos = b + n_in;
    for (int h = b; h < e; ++h) {
        if (owner_hf[h] == h) list[pos++] = make_int2(h, c);
    }
}

__global__ void k_porevol(EuGridDev g, double* __restrict__ porevol)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    porevol[c] = g.cell_volume[c]*g.poro[c];          // euler/EulerUpstream_impl.hpp:125
}

// ---------------------------------------------------------------------------------------
// SELL-32 structure for the FAST kernel
// ---------------------------------------------------------------------------------------
// Canonical face slots of the FAST layout.  The SELL slot of a half-face (and with it the plane of its face id) need
// not be the cell's local face index: FAST results are gated by a tolerance, not by the reference's summation order
// (STRICT keeps that order through its own list).  Per slice of 32 cells the neighbour offsets d = nbr - cell that a
// majority of the cells share become the slice's standard slots, in ascending order of d; a cell's remaining faces
// (boundary faces, fault faces, split faces) take its empty standard slots first and then extra slots.  A Cartesian
// slice is unchanged up to a permutation; a slice next to a fault plane keeps five regular slots instead of losing
// every slot behind the first cell with an extra face.  One warp per slice; cells with more than kMaxCanon faces
// keep their local order.
constexpr int kMaxCanon = 16;
constexpr int kNoOffset = INT_MIN;

__global__ void k_canonical_slots(EuGridDev g, unsigned char* __restrict__ slot_of_hf, int* __restrict__ slice_width)
{
    const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_slices = (g.n_local + EU_SLICE - 1)/EU_SLICE;
    if (warp >= n_slices) return;
    const int c = warp*EU_SLICE + lane;
    int b = 0, cnt = 0;
    if (c < g.n_local) { b = g.hf_offset[c]; cnt = g.hf_offset[c + 1] - b; }
    int maxcnt = cnt;
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
    int width = 0;
    if (maxcnt > kMaxCanon) {                       // warp-uniform: identity
        for (int j = 0; j < cnt; ++j) slot_of_hf[b + j] = (unsigned char)min(j, 255);
        width = maxcnt;
    } else {
        int d[kMaxCanon];
#pragma unroll
        for (int j = 0; j < kMaxCanon; ++j) {
            d[j] = kNoOffset;
            if (j < cnt) {
                const int n = g.hf_nbr[b + j];
                int other = n;
                if (n < 0) {
                    const int bi = -2 - n;
                    other = (g.bnd_kind[bi] == EU_HF_PERIODIC) ? g.bnd_partner_cell[bi] : -1;
                }
                if (other >= 0 && other != c) d[j] = other - c;
            }
        }
        // standard offsets by majority vote; every face of every lane is a candidate once
        const unsigned active = __ballot_sync(0xffffffffu, cnt > 0);
        const int n_active = __popc(active);
        int stdo[8];
        int n_std = 0;
        for (int l = 0; l < 32; ++l) {
#pragma unroll
            for (int j = 0; j < kMaxCanon; ++j) {
                if (j >= maxcnt) break;
                const int cand = __shfl_sync(0xffffffffu, d[j], l);
                if (cand == kNoOffset) continue;
                bool known = false;
                for (int q = 0; q < n_std; ++q) known = known || stdo[q] == cand;
                if (known || n_std >= 8) continue;
                bool mine = false;
#pragma unroll
                for (int i = 0; i < kMaxCanon; ++i) mine = mine || d[i] == cand;
                const int votes = __popc(__ballot_sync(0xffffffffu, mine));
                if (2*votes > n_active) stdo[n_std++] = cand;
            }
        }
        for (int i = 1; i < n_std; ++i) {           // ascending (warp-uniform insertion sort)
            const int v = stdo[i];
            int q = i;
            while (q > 0 && stdo[q - 1] > v) { stdo[q] = stdo[q - 1]; --q; }
            stdo[q] = v;
        }
        // standard faces to their slots (the first face of a cell with that offset), the others to the free slots
        unsigned used = 0u;
        int slot[kMaxCanon];
#pragma unroll
        for (int j = 0; j < kMaxCanon; ++j) {
            slot[j] = -1;
            if (j < cnt && d[j] != kNoOffset) {
                for (int q = 0; q < n_std; ++q) {
                    if (stdo[q] == d[j] && !((used >> q) & 1u)) { slot[j] = q; used |= 1u << q; break; }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kMaxCanon; ++j) {
            if (j < cnt) {
                if (slot[j] < 0) {
                    const int q = __ffs(~used) - 1;
                    slot[j] = q;
                    used |= 1u << q;
                }
                slot_of_hf[b + j] = (unsigned char)slot[j];
                width = max(width, slot[j] + 1);
            }
        }
        for (int o = 16; o > 0; o >>= 1) width = max(width, __shfl_xor_sync(0xffffffffu, width, o));
    }
    if (lane == 0) slice_width[warp] = width;
}

// Axis planes of the face arrays.  The (up to) three most common positive neighbour offsets of the grid -- 1, nx and
// nx*ny for a logically Cartesian numbering, whatever the grid interface calls its faces -- get the face planes 0, 1, 2:
// the face between cell c and cell c + ax[a] lives at a*n_local + c wherever it exists, and entries of cells without
// such a face stay zero (zero flux, G and T).  Every other face (boundary, fault, periodic wrap, split faces) lives in
// plane 3 + its canonical slot.  With that the regular faces of the whole grid are three dense arrays indexed by the
// lower cell, uniform across slices and planes: slice classes merge, and the box kernel (eu_tile.cuh) reads them as
// tiles.  `box` = {nx, nx*ny} restricts the axis planes to offsets that stay inside a row / plane of the box numbering.
__global__ void k_offset_votes(EuGridDev g, const int* __restrict__ cand, int n_cand, unsigned long long* __restrict__ votes)
{
    __shared__ unsigned int sv[16];
    if (threadIdx.x < 16) sv[threadIdx.x] = 0u;
    __syncthreads();
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c < g.n_local) {
        const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
        for (int h = b; h < e; ++h) {
            const int n = g.hf_nbr[h];
            if (n <= c) continue;
            const int d = n - c;
            for (int q = 0; q < n_cand; ++q) if (cand[q] == d) atomicAdd(&sv[q], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < n_cand && sv[threadIdx.x]) atomicAdd(votes + threadIdx.x, (unsigned long long)sv[threadIdx.x]);
}

// face id of an own half-face: axis plane (see above) or (3 + canonical slot)*n_local + cell
__global__ void k_assign_fid(EuGridDev g, const int* __restrict__ owner_hf, const unsigned char* __restrict__ slot_of_hf,
                             int ax0, int ax1, int ax2, int box, int* __restrict__ fid_of_hf)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
    unsigned taken = 0u;
    for (int h = b; h < e; ++h) {
        int fid = -1;
        if (owner_hf[h] == h) {
            fid = (3 + int(slot_of_hf[h]))*g.n_local + c;
            const int n = g.hf_nbr[h];
            if (n > c) {
                const int d = n - c;
                int a = -1;
                // the offset must stay inside the row (a = 0) / plane (a = 1) of the box numbering c = x + ax1*(y + ..)
                // (only when the local numbering is a box, `box`: otherwise c % ax1 says nothing about the row)
                if (ax0 > 0 && d == ax0 && (!box || (c % ax1) + ax0 < ax1)) a = 0;
                else if (ax1 > 0 && d == ax1 && (!box || (c % ax2) + ax1 < ax2)) a = 1;
                else if (ax2 > 0 && d == ax2) a = 2;
                if (a >= 0 && !((taken >> a) & 1u)) { taken |= 1u << a; fid = a*g.n_local + c; }
            }
        }
        fid_of_hf[h] = fid;
    }
}

// per cell: which record slots (SELL slot j < 16) hold a face that is NOT in an axis plane -- the faces the box kernel
// adds one by one from the records after its regular faces
// The own cells with a non-zero mask are appended to `list` (their number to *count): the work list of k_box_irregular.
__global__ void k_cell_mask(EuGridDev g, const int* __restrict__ slice_base, const int2* __restrict__ rec,
                            unsigned short* __restrict__ cmask, int* __restrict__ list, int* __restrict__ count)
{
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    const int s = c >> 5, lane = c & 31;
    const int base = slice_base[s];
    const int width = (slice_base[s + 1] - base) >> 5;
    unsigned m = 0u;
    for (int j = 0; j < width && j < 16; ++j) {
        const int2 r = rec[(long long)base + (long long)j*EU_SLICE + lane];
        if (r.x != EU_REC_PAD && r.y >= 3*g.n_local) m |= 1u << j;
    }
    cmask[c] = (unsigned short)m;
    // warp-aggregated append (ascending within a warp; the order across warps does not matter)
    const bool mine = m != 0u && c >= g.own_lo && c < g.own_hi;
    const unsigned act = __activemask();
    const unsigned vote = __ballot_sync(act, mine);
    if (vote) {
        const int leader = __ffs(vote) - 1;
        int start = 0;
        if ((threadIdx.x & 31) == leader) start = atomicAdd(count, __popc(vote));
        start = __shfl_sync(act, start, leader);
        if (mine) list[start + __popc(vote & ((1u << (threadIdx.x & 31)) - 1u))] = c;
    }
}

__global__ void k_build_records(EuGridDev g, const int* __restrict__ owner_hf, const int* __restrict__ fid_of_hf,
                                const unsigned char* __restrict__ slot_of_hf, const int* __restrict__ slice_base,
                                int2* __restrict__ rec, int2* __restrict__ desc, int* __restrict__ n_regular_slots)
{
    const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_slices = (g.n_local + EU_SLICE - 1)/EU_SLICE;
    if (warp >= n_slices) return;
    const int c = warp*EU_SLICE + lane;
    int b = 0, cnt = 0;
    if (c < g.n_local) { b = g.hf_offset[c]; cnt = g.hf_offset[c + 1] - b; }
    const int base = slice_base[warp];
    const int width = (slice_base[warp + 1] - base)/EU_SLICE;
    // every lane owns column `lane` of the slice: pad it, then drop its faces into their canonical slots
    for (int j = 0; j < width; ++j) rec[(long long)base + (long long)j*EU_SLICE + lane] = make_int2(EU_REC_PAD, -1);
    for (int k = 0; k < cnt; ++k) {
        const int h = b + k;
        const int o = owner_hf[h];
        const int n = g.hf_nbr[h];
        int2 r = make_int2(EU_REC_PAD, -1);
        if (o >= 0) {
            r.y = fid_of_hf[o];
            if (n >= 0) {
                r.x = n;
            } else {
                const int bi = -2 - n;
                r.x = (g.bnd_kind[bi] == EU_HF_PERIODIC) ? g.bnd_partner_cell[bi] : n;
            }
        }
        const int j = slot_of_hf[h];
        if (j < width) rec[(long long)base + (long long)j*EU_SLICE + lane] = r;
    }
    for (int j = 0; j < width; ++j) {
        const int2 r = rec[(long long)base + (long long)j*EU_SLICE + lane];      // written by this very thread
        // regular slot: every lane has an interior (or periodic) neighbour at the same offset d whose face
        // lives in the same plane k, so that nbr = c + d and fid = k*n_local + (d > 0 ? c : c + d)
        const int d = r.x - c;
        const int k = (r.x >= 0 && r.y >= 0) ? r.y/g.n_local : -1;
        const int d0 = __shfl_sync(0xffffffffu, d, 0), k0 = __shfl_sync(0xffffffffu, k, 0);
        const bool lane_ok = c < g.n_local && r.x >= 0 && k >= 0 && d == d0 && k == k0 && d != 0 &&
                             r.y == k*g.n_local + (d > 0 ? c : r.x);
        const bool regular = __all_sync(0xffffffffu, lane_ok);
        const bool any = __any_sync(0xffffffffu, r.x != EU_REC_PAD);
        if (lane == 0) {
            desc[base/EU_SLICE + j] = regular ? make_int2(d0, k0) : make_int2(0, any ? -1 : -2);
            if (regular) atomicAdd(n_regular_slots, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Contraction of the static per-face geometry/permeability to scalars (FAST mode inputs),
// in the reference's operation order:
//   G = area * inner(n, (Kavg g) drho)            euler/EulerUpstreamResidual_impl.hpp:195-205
//   T = area * inner(n, Kavg dhat) / (d0 + d1)    :271-273 with the direction of :532-545
// ---------------------------------------------------------------------------------------
__global__ void k_contract(EuGridDev g, EuTablesDev t, const int* __restrict__ owner_hf,
                           const int* __restrict__ fid_of_hf, double gx, double gy, double gz, int method_gravity,
                           double* __restrict__ G, double* __restrict__ T, double* __restrict__ nn,
                           unsigned long long* __restrict__ nn_maxdev_bits, double* __restrict__ Ga, int* __restrict__ gmask)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
    const double gravity[3] = { gx, gy, gz };
    double maxdev = 0.0;
    for (int h = b; h < e; ++h) {
        if (owner_hf[h] != h) continue;
        const int fid = fid_of_hf[h];
        const int n = g.hf_nbr[h];
        int c1 = c, nbhf = h;
        bool interior_like = true;
        if (n >= 0) {
            c1 = n;
        } else {
            const int bi = -2 - n;
            if (g.bnd_kind[bi] == EU_HF_PERIODIC) { c1 = g.bnd_partner_cell[bi]; nbhf = g.bnd_partner_hf[bi]; }
            else interior_like = false;
        }
        double aver[9];
        sm_aver9(g.perm + 9LL*c, g.perm + 9LL*c1, aver);
        double gi[3];
        sm_prod3(aver, gravity, gi);
        for (int i = 0; i < 3; ++i) gi[i] *= t.delta_rho;
        const double area = g.hf_area[h];
        const double nrm[3] = { g.hf_normal[3LL*h], g.hf_normal[3LL*h + 1], g.hf_normal[3LL*h + 2] };
        const double Gv = method_gravity ? area*sm_inner3(nrm, gi) : 0.0;
        G[2LL*fid] = Gv;                                                   // interleaved {q, G} pairs
        if (Ga && fid < 3*g.n_local) {                                     // axis planes: also as a separate array
            Ga[fid] = Gv;
            if (Gv != 0.0 && gmask) atomicOr(gmask, 1 << (fid/g.n_local));
        }
        double Tv = 0.0;
        if (interior_like) {
            double dirhat[3], d0d1, ci[3];
            sm_cap_direction(g.cell_centroid + 3LL*c, g.cell_centroid + 3LL*c1, g.hf_centroid + 3LL*h,
                             g.hf_centroid + 3LL*nbhf, dirhat, &d0d1);
            sm_prod3(aver, dirhat, ci);
            Tv = area*sm_inner3(nrm, ci)/d0d1;
        }
        T[fid] = Tv;
        const double nnv = sm_inner3(nrm, nrm);
        if (nn) nn[fid] = nnv;
        maxdev = fmax(maxdev, fabs(nnv - 1.0));
    }
    if (maxdev > 0.0) atomicMax(nn_maxdev_bits, (unsigned long long)__double_as_longlong(maxdev));
}

// FAST mode with diagonal tensor mobility needs axis-aligned face normals (eu_fast.cu): exactly one component
// of magnitude > 1e-14.
__device__ __forceinline__ int normal_axis(const double* __restrict__ n, bool* aligned)
{
    const double a0 = fabs(n[0]), a1 = fabs(n[1]), a2 = fabs(n[2]);
    const int ax = (a0 >= a1 && a0 >= a2) ? 0 : (a1 >= a2 ? 1 : 2);
    const double minor = (ax == 0) ? fmax(a1, a2) : (ax == 1 ? fmax(a0, a2) : fmax(a0, a1));
    *aligned = minor <= 1e-14;
    return ax;
}
__global__ void k_axis_check(EuGridDev g, int* __restrict__ flag)
{
    const long long h = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (h >= g.H) return;
    bool aligned;
    normal_axis(g.hf_normal + 3*h, &aligned);
    if (!aligned) atomicOr(flag, 1);
}
__global__ void k_face_axis(EuGridDev g, const int* __restrict__ owner_hf, const int* __restrict__ fid_of_hf,
                            unsigned char* __restrict__ axis8)
{
    const long long h = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (h >= g.H || owner_hf[h] != h) return;
    bool aligned;
    axis8[fid_of_hf[h]] = (unsigned char)normal_axis(g.hf_normal + 3*h, &aligned);
}

// Static per-face vectors of the FAST path for diagonal tensor mobility on general (oblique) normals: with
// M = diag(m_x, m_y, m_z) the terms of the face flux (Residual_impl.hpp:205-262) are sums over the axes,
//   n.(M x) = sum_k n_k m_k x_k,
// so the mobility-independent factors are kept per axis: Gv_k = area n_k ((K g) drho)_k, nn_k = n_k^2,
// Tv_k = area n_k (K dhat)_k / (d0 + d1).  Layout: fv[k*stride + face], k = 0..2 Gv, 3..5 nn, 6..8 Tv.
__global__ void k_contract_t3(EuGridDev g, EuTablesDev t, const int* __restrict__ owner_hf, const int* __restrict__ fid_of_hf,
                              double gx, double gy, double gz, int method_gravity, double* __restrict__ fv, long long stride)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
    const double gravity[3] = { gx, gy, gz };
    for (int h = b; h < e; ++h) {
        if (owner_hf[h] != h) continue;
        const int fid = fid_of_hf[h];
        const int n = g.hf_nbr[h];
        int c1 = c, nbhf = h;
        bool interior_like = true;
        if (n >= 0) {
            c1 = n;
        } else {
            const int bi = -2 - n;
            if (g.bnd_kind[bi] == EU_HF_PERIODIC) { c1 = g.bnd_partner_cell[bi]; nbhf = g.bnd_partner_hf[bi]; }
            else interior_like = false;
        }
        double aver[9];
        sm_aver9(g.perm + 9LL*c, g.perm + 9LL*c1, aver);
        double gi[3];
        sm_prod3(aver, gravity, gi);
        for (int i = 0; i < 3; ++i) gi[i] *= t.delta_rho;
        const double area = g.hf_area[h];
        const double nrm[3] = { g.hf_normal[3LL*h], g.hf_normal[3LL*h + 1], g.hf_normal[3LL*h + 2] };
        double ci[3] = { 0.0, 0.0, 0.0 }, d0d1 = 1.0;
        if (interior_like) {
            double dirhat[3];
            sm_cap_direction(g.cell_centroid + 3LL*c, g.cell_centroid + 3LL*c1, g.hf_centroid + 3LL*h,
                             g.hf_centroid + 3LL*nbhf, dirhat, &d0d1);
            sm_prod3(aver, dirhat, ci);
        }
        for (int k = 0; k < 3; ++k) {
            fv[(long long)k*stride + fid] = method_gravity ? area*(nrm[k]*gi[k]) : 0.0;
            fv[(long long)(3 + k)*stride + fid] = nrm[k]*nrm[k];
            fv[(long long)(6 + k)*stride + fid] = interior_like ? area*(nrm[k]*ci[k])/d0d1 : 0.0;
        }
    }
}

__global__ void k_pcscale(EuGridDev g, EuTablesDev t, double* __restrict__ pcscale, unsigned char* __restrict__ rock8,
                          double* __restrict__ inv_porevol)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    double sc = 1.0;
    if (t.kind == EU_MOB_SCALAR && t.n_rocks > 0 && t.use_j) {
        const double* K = g.perm + 9LL*c;
        double tr = 0;
        tr += K[0]; tr += K[4]; tr += K[8];
        sc = t.sigma_cos_theta/sqrt(tr/(3*g.poro[c]));
    }
    pcscale[c] = sc;
    rock8[c] = (unsigned char)g.rock[c];
    inv_porevol[c] = 1.0/(g.cell_volume[c]*g.poro[c]);
}

// ---------------------------------------------------------------------------------------
// min reductions (order independent; NaN candidates never win, like "if (x < dt) dt = x")
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double min_keep(double a, double b) { return (b < a) ? b : a; }

__device__ __forceinline__ void block_min_store(double v, double* __restrict__ block_min)
{
    __shared__ double sm[kThreads/32];
    for (int o = 16; o > 0; o >>= 1) v = min_keep(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < kThreads/32 ? sm[threadIdx.x] : 1e100;
        for (int o = 16; o > 0; o >>= 1) w = min_keep(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) block_min[blockIdx.x] = w;
    }
}

// min over n block minima: a grid of up to kFinalBlocks blocks writes one partial minimum each behind the inputs
// (block_min has room for them, eu_cfl_blocks), a last block reduces those.  (One block over all the minima of a 67 M-cell
// grid was a chain of a thousand dependent loads per thread: 60 us per CFL term.)  The minimum is exact in any order.
constexpr int kFinalBlocks = 256;
__global__ void k_final_min(const double* __restrict__ block_min, int n, double* __restrict__ out)
{
    double v = 1e100;
    for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) v = min_keep(v, block_min[i]);
    out += blockIdx.x;
    __shared__ double sm[kThreads/32];
    for (int o = 16; o > 0; o >>= 1) v = min_keep(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < kThreads/32 ? sm[threadIdx.x] : 1e100;
        for (int o = 16; o > 0; o >>= 1) w = min_keep(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) out[0] = w;
    }
}

// euler/CflCalculator.hpp:54-83 fused with the per-call compaction of the half-face fluxes to
// one value per unique face (the owner's outflux, euler/EulerUpstreamResidual_impl.hpp:147)
__global__ void k_cfl_velocity_compact(EuGridDev g, double cfl_factor, const double* __restrict__ hf_flux,
                                       const int* __restrict__ fid_of_hf, double* __restrict__ q, double* __restrict__ qa,
                                       double* __restrict__ block_min, int* __restrict__ zero_flag)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    double dt = 1e100;
    if (c < g.n_local) {
        const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
        double flux_p = 0.0, flux_n = 0.0;
        for (int h = b; h < e; ++h) {
            const double f = hf_flux[h];
            if (f > 0) flux_p += f; else flux_n -= f;
            if (q) {
                const int fid = fid_of_hf[h];
                if (fid >= 0) {
                    q[2LL*fid] = f;                                        // interleaved {q, G} pairs
                    if (qa && fid < 3*g.n_local) qa[fid] = f;              // axis planes: also as a separate array (box kernel)
                }
            }
        }
        if (c >= g.own_lo && c < g.own_hi) {
            const double flux = flux_n > flux_p ? flux_n : flux_p;         // std::max(flux_n, flux_p)
            const double loc_dt = (cfl_factor*g.cell_volume[c]*g.poro[c])/flux;
            if (loc_dt == 0.0) atomicExch(zero_flag, 1);
            if (loc_dt < dt) dt = loc_dt;
        }
    }
    block_min_store(dt, block_min);
}

// euler/CflCalculator.hpp:90-134
__global__ void k_cfl_gravity(EuGridDev g, EuTablesDev t, double cfl_factor, double gx, double gy, double gz,
                              double* __restrict__ block_min)
{
    int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    double dt = 1e100;
    if (c < g.own_hi) {
        const double gravity[3] = { gx, gy, gz };
        const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
        double flux = 0.0;
        for (int h = b; h < e; ++h) {
            double aver[9];
            const double* K = g.perm + 9LL*c;
            const int n = g.hf_nbr[h];
            if (n >= 0) { sm_aver9(g.perm + 9LL*c, g.perm + 9LL*n, aver); K = aver; }
            double lgf = 0.0;
            for (int k = 0; k < 3; ++k) {
                for (int qd = 0; qd < 3; ++qd) {
                    lgf += g.hf_normal[3LL*h + qd]*(K[3*qd + k]*gravity[k]*t.delta_rho);
                }
            }
            lgf *= g.hf_area[h];
            if (lgf > 0) flux += lgf;
        }
        const double loc_dt = (cfl_factor*g.cell_volume[c]*g.poro[c])/flux;
        if (loc_dt < dt) dt = loc_dt;
    }
    block_min_store(dt, block_min);
}

// common/MatrixInverse.hpp:85-123
__device__ __forceinline__ void sm_inverse3x3(const double* m, double* mi)
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

// euler/CflCalculator.hpp:142-176
__global__ void k_cfl_capillary(EuGridDev g, double cfl_factor, double* __restrict__ block_min)
{
    int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    double dt = 1e100;
    if (c < g.own_hi) {
        const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
        for (int h = b; h < e; ++h) {
            double aver[9], inv[9];
            const double* K = g.perm + 9LL*c;
            const int n = g.hf_nbr[h];
            if (n >= 0) { sm_aver9(g.perm + 9LL*c, g.perm + 9LL*n, aver); K = aver; }
            sm_inverse3x3(K, inv);
            double d[3], v[3];
            for (int i = 0; i < 3; ++i) d[i] = g.hf_centroid[3LL*h + i] - g.cell_centroid[3LL*c + i];
            sm_prod3(inv, d, v);
            double spatial = 0.0;
            spatial += d[0]*v[0];
            spatial += d[1]*v[1];
            spatial += d[2]*v[2];
            const double loc_dt = spatial/cfl_factor;
            dt = min_keep(dt, loc_dt);
        }
    }
    block_min_store(dt, block_min);
}

// ---------------------------------------------------------------------------------------
// STRICT substep.  euler/EulerUpstreamResidual_impl.hpp:459-467 (cap pressures), :100-300
// (face flux), euler/EulerUpstream_impl.hpp:355-385 + :336-349 (update, range check).
// ---------------------------------------------------------------------------------------
__global__ void k_strict_pc(EuGridDev g, EuTablesDev t, const double* __restrict__ S, double* __restrict__ pc)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.n_local) return;
    pc[c] = sm_cap_pressure(t, g.rock[c], g.perm + 9LL*c, g.poro[c], S[c]);
}

template <int KIND>
__device__ double strict_face_flux(const EuGridDev& g, const EuTablesDev& t, const EuStepArgs& a,
                                   const double* __restrict__ hf_flux, int h, int c0)
{
    constexpr int M = KIND == 0 ? 1 : 9;
    int cell[2];
    double cell_sat[2];
    cell[0] = c0;
    cell_sat[0] = a.S_in[c0];
    int nbhf = h;
    bool interior_like = true;
    const int n = g.hf_nbr[h];
    if (n >= 0) {
        cell[1] = n;
        cell_sat[1] = a.S_in[n];
    } else {
        const int bi = -2 - n;
        if (g.bnd_kind[bi] == EU_HF_PERIODIC) {
            nbhf = g.bnd_partner_hf[bi];
            cell[1] = g.bnd_partner_cell[bi];
            cell_sat[1] = a.S_in[cell[1]];
        } else {
            cell[1] = c0;
            cell_sat[1] = g.bnd_sat[bi];
            interior_like = false;
        }
    }
    const int rock[2] = { g.rock[cell[0]], g.rock[cell[1]] };
    const double loc_area = g.hf_area[h];
    const double loc_flux = hf_flux[h];
    const double loc_normal[3] = { g.hf_normal[3LL*h], g.hf_normal[3LL*h + 1], g.hf_normal[3LL*h + 2] };

    double aver_perm[9];
    sm_aver9(g.perm + 9LL*cell[0], g.perm + 9LL*cell[1], aver_perm);
    double grav_influence[3];
    sm_prod3(aver_perm, a.gravity, grav_influence);
    for (int i = 0; i < 3; ++i) grav_influence[i] *= t.delta_rho;
    const double G = a.method_gravity ? loc_area*sm_inner3(loc_normal, grav_influence) : 0.0;
    const int triv_phase = G >= 0.0 ? 0 : 1;
    const int ups_cell = loc_flux >= 0.0 ? 0 : 1;
    double m_ups[2][M];
    sm_mobility<KIND>(t, triv_phase, rock[ups_cell], cell_sat[ups_cell], m_ups[triv_phase]);
    const double sign_G = triv_phase == 0 ? -1.0 : 1.0;
    double tmp[3], tmp2[3], tmp3[3];
    sm_mob_multiply<KIND>(m_ups[triv_phase], grav_influence, tmp);
    const double grav_flux_nontriv = sign_G*loc_area*sm_inner3(loc_normal, tmp);
    const int ups_cell_nontriv = (loc_flux + grav_flux_nontriv >= 0.0) ? 0 : 1;
    const int nontriv_phase = (triv_phase + 1) % 2;
    sm_mobility<KIND>(t, nontriv_phase, rock[ups_cell_nontriv], cell_sat[ups_cell_nontriv], m_ups[nontriv_phase]);
    double m_tot[M], m_totinv[M];
    sm_mob_sum<KIND>(m_ups[0], m_ups[1], m_tot);
    sm_mob_inverse<KIND>(m_tot, m_totinv);

    double dS = 0.0;
    if (a.method_viscous) {
        double v[3] = { loc_normal[0], loc_normal[1], loc_normal[2] };
        for (int i = 0; i < 3; ++i) v[i] *= loc_flux;
        sm_mob_multiply<KIND>(m_totinv, v, tmp);
        sm_mob_multiply<KIND>(m_ups[0], tmp, tmp2);
        dS += sm_inner3(loc_normal, tmp2);
    }
    if (a.method_gravity) {
        if (cell[0] != cell[1]) {
            sm_mob_multiply<KIND>(m_ups[1], grav_influence, tmp);
            sm_mob_multiply<KIND>(m_totinv, tmp, tmp2);
            sm_mob_multiply<KIND>(m_ups[0], tmp2, tmp3);
            dS += loc_area*sm_inner3(loc_normal, tmp3);
        }
    }
    if (a.method_capillary) {
        // the average-saturation mobilities only feed this term (:228-242, :268-281)
        double aver_sat = cell_sat[0];
        aver_sat += cell_sat[1];
        aver_sat *= 0.5;
        double m1c0[M], m1c1[M], m2c0[M], m2c1[M];
        sm_mobility<KIND>(t, 0, rock[0], aver_sat, m1c0);
        sm_mobility<KIND>(t, 0, rock[1], aver_sat, m1c1);
        sm_mobility<KIND>(t, 1, rock[0], aver_sat, m2c0);
        sm_mobility<KIND>(t, 1, rock[1], aver_sat, m2c1);
        double m_aver[2][M], m_aver_tot[M], m_aver_totinv[M];
        sm_mob_average<KIND>(m1c0, m1c1, m_aver[0]);
        sm_mob_average<KIND>(m2c0, m2c1, m_aver[1]);
        sm_mob_sum<KIND>(m_aver[0], m_aver[1], m_aver_tot);
        sm_mob_inverse<KIND>(m_aver_tot, m_aver_totinv);
        double grad[3] = { 0.0, 0.0, 0.0 };
        if (interior_like) {
            double d0d1;
            sm_cap_direction(g.cell_centroid + 3LL*cell[0], g.cell_centroid + 3LL*cell[1], g.hf_centroid + 3LL*h,
                             g.hf_centroid + 3LL*nbhf, grad, &d0d1);
            const double val = (a.pc_in[cell[1]] - a.pc_in[cell[0]])/d0d1;
            for (int i = 0; i < 3; ++i) grad[i] *= val;
        }
        double cap_influence[3];
        sm_prod3(aver_perm, grad, cap_influence);
        sm_mob_multiply<KIND>(m_aver[1], cap_influence, tmp);
        sm_mob_multiply<KIND>(m_aver_totinv, tmp, tmp2);
        sm_mob_multiply<KIND>(m_aver[0], tmp2, tmp3);
        dS += loc_area*sm_inner3(loc_normal, tmp3);
    }
    return dS;
}

template <int KIND>
__global__ void __launch_bounds__(128) k_strict_step(EuGridDev g, EuTablesDev t, EuStrictDev s,
                                                      const double* __restrict__ hf_flux, EuStepArgs a)
{
    {   // an earlier substep of this attempt already failed: the attempt is void, stop working
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.own_hi) return;
    const int b = g.hf_offset[c], e = g.hf_offset[c + 1];
    double residual = 0.0;
    for (int k = b; k < e; ++k) {
        const int2 ent = s.list[k];
        const double dS = strict_face_flux<KIND>(g, t, a, hf_flux, ent.x, ent.y);
        if (ent.y == c) residual -= dS; else residual += dS;
    }
    // source term (:292-299)
    double rate = 0.0;
    if (a.n_src > 0) {
        int lo = 0, hi = a.n_src;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (a.src_cell[mid] < c) lo = mid + 1; else hi = mid; }
        if (lo < a.n_src && a.src_cell[lo] == c) rate = a.src_rate[lo];
    }
    if (rate < 0.0) {
        const double s0 = a.S_in[c];
        constexpr int M = KIND == 0 ? 1 : 9;
        double m1[M], m2[M];
        sm_mobility<KIND>(t, 0, g.rock[c], s0, m1);
        sm_mobility<KIND>(t, 1, g.rock[c], s0, m2);
        double ff;
        if (KIND == 0) {
            ff = m1[0]/(m1[0] + m2[0]);
        } else {
            ff = 0.0;
            for (int d = 0; d < 3; ++d) { double l1 = m1[(KIND ? 4 : 0)*d], l2 = m2[(KIND ? 4 : 0)*d]; ff += l1/(l1 + l2); }
            ff /= 3.0;
        }
        rate *= ff;
    }
    residual += rate;
    if (a.residual_out) a.residual_out[c] = residual;
    double sat = a.S_in[c];
    const double sat_change = a.dt*residual/s.porevol[c];
    sat += sat_change;
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
}

// ---------------------------------------------------------------------------------------
// Halo exchange.  The cells a neighbour rank holds as ghosts are written into ITS saturation buffer
// with plain stores through a peer mapping (NVLink P2P); the last block to finish publishes the epoch in
// the neighbour's flag word.  The neighbour spins on its own flag word before its next substep.
// ---------------------------------------------------------------------------------------
__global__ void k_halo_push(const int* __restrict__ send_src, const int* __restrict__ send_dst, int n,
                            const double* __restrict__ S_local, double* __restrict__ S_peer,
                            const double* __restrict__ pc_local, double* __restrict__ pc_peer,
                            unsigned* __restrict__ block_counter, unsigned* peer_flag, unsigned epoch)
{
    for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
        const int s = send_src[i], d = send_dst[i];
        S_peer[d] = S_local[s];
        if (pc_peer) pc_peer[d] = pc_local[s];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(block_counter, 1u);
        if (done == gridDim.x - 1) {
            *block_counter = 0;
            __threadfence_system();
            *(volatile unsigned*)peer_flag = epoch;
            __threadfence_system();
        }
    }
}

__global__ void k_halo_wait(const unsigned* my_flags, const int* __restrict__ wait_ranks, int n_wait, unsigned epoch,
                            long long timeout_cycles, int* err_flag)
{
    if (threadIdx.x >= n_wait) return;
    const volatile unsigned* f = my_flags + wait_ranks[threadIdx.x];
    const long long t0 = clock64();
    while ((int)(*f - epoch) < 0) {
        __nanosleep(200);
        if (clock64() - t0 > timeout_cycles) { atomicExch(err_flag, 1); break; }   // never hang the GPU
    }
    __threadfence_system();
}

// own cells that read a ghost: out[0] = max index of an own cell with a neighbour below the own range,
// out[1] = min index of an own cell with a neighbour above it (boundary ranges of the fused halo exchange)
__global__ void k_ghost_adjacent(EuGridDev g, int* __restrict__ out)
{
    int c = g.own_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= g.own_hi) return;
    bool down = false, up = false;
    for (int h = g.hf_offset[c]; h < g.hf_offset[c + 1]; ++h) {
        const int n = hf_other_cell(g, c, h);
        if (n >= 0 && n < g.own_lo) down = true;
        if (n >= g.own_hi) up = true;
    }
    if (down) atomicMax(out + 0, c);
    if (up) atomicMin(out + 1, c);
}

} // namespace

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
void eu_launch_translate_nbr(int* hf_nbr, long long H, const int* rf, const int* rc, const int* rl, int nr, cudaStream_t st)
{
    if (H > 0) k_translate_nbr<<<div_up(H, kThreads), kThreads, 0, st>>>(hf_nbr, H, rf, rc, rl, nr);
}
void eu_launch_owner(const EuGridDev& g, int* owner_hf, int* err_flag, cudaStream_t st)
{
    k_owner<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, owner_hf, err_flag);
}
void eu_launch_strict_list(const EuGridDev& g, const int* owner_hf, int2* list, cudaStream_t st)
{
    k_strict_list<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, owner_hf, list);
}
void eu_launch_porevol(const EuGridDev& g, double* porevol, cudaStream_t st)
{
    k_porevol<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, porevol);
}
void eu_launch_canonical_slots(const EuGridDev& g, unsigned char* slot_of_hf, int* slice_width, cudaStream_t st)
{
    const int n_slices = (g.n_local + EU_SLICE - 1)/EU_SLICE;
    k_canonical_slots<<<div_up((long long)n_slices*32, kThreads), kThreads, 0, st>>>(g, slot_of_hf, slice_width);
}
void eu_launch_assign_fid(const EuGridDev& g, const int* owner_hf, const unsigned char* slot_of_hf, const int ax[3], int box,
                          int* fid_of_hf, cudaStream_t st)
{
    k_assign_fid<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, owner_hf, slot_of_hf, ax[0], ax[1], ax[2], box, fid_of_hf);
}
void eu_launch_offset_votes(const EuGridDev& g, const int* cand, int n_cand, unsigned long long* votes, cudaStream_t st)
{
    k_offset_votes<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, cand, n_cand, votes);
}
void eu_launch_cell_mask(const EuGridDev& g, const int* slice_base, const int2* rec, unsigned short* cmask, int* list, int* count,
                         cudaStream_t st)
{
    k_cell_mask<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, slice_base, rec, cmask, list, count);
}
void eu_launch_build_records(const EuGridDev& g, const int* owner_hf, const int* fid_of_hf, const unsigned char* slot_of_hf,
                             const int* slice_base, int2* rec, int2* desc, int* n_regular_slots, cudaStream_t st)
{
    const int n_slices = (g.n_local + EU_SLICE - 1)/EU_SLICE;
    k_build_records<<<div_up((long long)n_slices*32, kThreads), kThreads, 0, st>>>(g, owner_hf, fid_of_hf, slot_of_hf, slice_base,
                                                                                  rec, desc, n_regular_slots);
}
void eu_launch_contract(const EuGridDev& g, const EuTablesDev& t, const int* owner_hf, const int* fid_of_hf,
                        const double gravity[3], int method_gravity, double* G, double* T, double* nn,
                        double* nn_maxdev, double* Ga, int* gmask_out, cudaStream_t st)
{
    cudaMemsetAsync(nn_maxdev, 0, sizeof(double), st);
    if (gmask_out) cudaMemsetAsync(gmask_out, 0, sizeof(int), st);
    k_contract<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, t, owner_hf, fid_of_hf, gravity[0], gravity[1], gravity[2],
                                                                 method_gravity, G, T, nn, (unsigned long long*)nn_maxdev, Ga, gmask_out);
}
void eu_launch_contract_t3(const EuGridDev& g, const EuTablesDev& t, const int* owner_hf, const int* fid_of_hf,
                           const double gravity[3], int method_gravity, double* fv, long long stride, cudaStream_t st)
{
    k_contract_t3<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, t, owner_hf, fid_of_hf, gravity[0], gravity[1], gravity[2],
                                                                    method_gravity, fv, stride);
}
void eu_launch_axis_check(const EuGridDev& g, int* flag, cudaStream_t st)
{
    if (g.H > 0) k_axis_check<<<div_up(g.H, kThreads), kThreads, 0, st>>>(g, flag);
}
void eu_launch_face_axis(const EuGridDev& g, const int* owner_hf, const int* fid_of_hf, unsigned char* axis8, cudaStream_t st)
{
    if (g.H > 0) k_face_axis<<<div_up(g.H, kThreads), kThreads, 0, st>>>(g, owner_hf, fid_of_hf, axis8);
}
void eu_launch_pcscale(const EuGridDev& g, const EuTablesDev& t, double* pcscale, unsigned char* rock8, double* inv_porevol,
                       cudaStream_t st)
{
    k_pcscale<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, t, pcscale, rock8, inv_porevol);
}
// blocks of the CFL kernels = block minima, plus room for the partial minima of the final reduction
int eu_cfl_blocks(int n_cells) { return div_up(n_cells > 0 ? n_cells : 1, kThreads) + kFinalBlocks; }
static int cfl_grid(int n_cells) { return div_up(n_cells > 0 ? n_cells : 1, kThreads); }
static void launch_final_min(double* block_min, int nb, double* out, cudaStream_t st)
{
    const int parts = std::min(kFinalBlocks, div_up(nb, kThreads));
    if (parts <= 1) { k_final_min<<<1, kThreads, 0, st>>>(block_min, nb, out); return; }
    k_final_min<<<parts, kThreads, 0, st>>>(block_min, nb, block_min + nb);
    k_final_min<<<1, kThreads, 0, st>>>(block_min + nb, parts, out);
}

void eu_launch_cfl_velocity_compact(const EuGridDev& g, double cfl_factor, const double* hf_flux, const int* fid_of_hf,
                                    double* q, double* qa, double* block_min, int* zero_flag, double* out, cudaStream_t st)
{
    const int nb = cfl_grid(g.n_local);
    k_cfl_velocity_compact<<<nb, kThreads, 0, st>>>(g, cfl_factor, hf_flux, fid_of_hf, q, qa, block_min, zero_flag);
    launch_final_min(block_min, nb, out, st);
}
void eu_launch_cfl_gravity(const EuGridDev& g, const EuTablesDev& t, double cfl_factor, const double gravity[3],
                           double* block_min, double* out, cudaStream_t st)
{
    const int nb = cfl_grid(g.own_hi - g.own_lo);
    k_cfl_gravity<<<nb, kThreads, 0, st>>>(g, t, cfl_factor, gravity[0], gravity[1], gravity[2], block_min);
    launch_final_min(block_min, nb, out, st);
}
void eu_launch_cfl_capillary(const EuGridDev& g, double cfl_factor, double* block_min, double* out, cudaStream_t st)
{
    const int nb = cfl_grid(g.own_hi - g.own_lo);
    k_cfl_capillary<<<nb, kThreads, 0, st>>>(g, cfl_factor, block_min);
    launch_final_min(block_min, nb, out, st);
}
void eu_launch_strict_pc(const EuGridDev& g, const EuTablesDev& t, const double* S, double* pc, cudaStream_t st)
{
    k_strict_pc<<<div_up(g.n_local, kThreads), kThreads, 0, st>>>(g, t, S, pc);
}
void eu_launch_strict_step(const EuGridDev& g, const EuTablesDev& t, const EuStrictDev& s, const double* hf_flux,
                           const EuStepArgs& a, cudaStream_t st)
{
    const int n = g.own_hi - g.own_lo;
    if (n <= 0) return;
    if (t.kind == EU_MOB_SCALAR) k_strict_step<0><<<div_up(n, 128), 128, 0, st>>>(g, t, s, hf_flux, a);
    else                         k_strict_step<1><<<div_up(n, 128), 128, 0, st>>>(g, t, s, hf_flux, a);
}

void eu_launch_halo_push(const int* send_src, const int* send_dst, int n, const double* S_local, double* S_peer,
                         const double* pc_local, double* pc_peer, unsigned* block_counter, unsigned* peer_flag,
                         unsigned epoch, cudaStream_t st)
{
    int blocks = div_up(n > 0 ? n : 1, kThreads);
    if (blocks > 64) blocks = 64;
    k_halo_push<<<blocks, kThreads, 0, st>>>(send_src, send_dst, n, S_local, S_peer, pc_local, pc_peer, block_counter,
                                             peer_flag, epoch);
}
void eu_launch_halo_wait(const unsigned* my_flags, const int* wait_ranks, int n_wait, unsigned epoch,
                         long long timeout_cycles, int* err_flag, cudaStream_t st)
{
    if (n_wait <= 0) return;
    k_halo_wait<<<1, 32, 0, st>>>(my_flags, wait_ranks, n_wait, epoch, timeout_cycles, err_flag);
}

void eu_launch_ghost_adjacent(const EuGridDev& g, int* out4, cudaStream_t st)
{
    const int n = g.own_hi - g.own_lo;
    if (n > 0) k_ghost_adjacent<<<div_up(n, kThreads), kThreads, 0, st>>>(g, out4);
}
