// Host-only helpers of the C ABI: no device needed.
//
// eu_compute_cfl_factors restates ReservoirPropertyCapillary<3>::computeCflFactors
// (common/ReservoirPropertyCapillary_impl.hpp:190-281, RockJfunc.hpp:114-127) for callers that do
// not hold the reference's property class (bench.py, the Python binding).  The drop-in C++ header
// does not use it: it asks the caller's ReservoirProperties object for cflFactor*() directly.
#include "../../include/euler_b200.h"
#include "eu_box_units.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iterator>
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

// ---------------------------------------------------------------------------------------------------------
// Periodic partner matching of boundary faces: findPeriodicPartners (common/BoundaryPeriodicity.hpp:86-177) with
// match() (common/BoundaryPeriodicity.cpp:25-49), the host-side step that produces the partner table the transport
// path consumes (eu_grid_chunk::bnd_partner_*).  Same definitions as the reference: bounding box of the face
// centroids; canonical side = the first coordinate direction whose low (then high) bound the centroid touches within
// the spatial tolerance; partners sit on opposite sides, have areas within 1e-6 and centroids -- with the normal
// coordinate zeroed -- within 1e-6 of each other.  The reference sorts the faces by a scalar key
// (c_a + pi c_b of the two transverse coordinates), looks 10 places either way and falls back to a scan over all
// faces, O(n^2) when many faces have no partner; here every candidate within the tolerance is found in the sorted key
// band [key - (1 + pi) tol, key + (1 + pi) tol], O(n log n) on the 1 M boundary faces of the 67 M-cell grid.
// Where exactly one face qualifies -- every well-posed periodic grid -- the result is the reference's.
extern "C" int eu_match_periodic_faces(int n, const double* centroid, const double* area, const int is_periodic[6],
                                       double spatial_tolerance, int* canon_pos, int* partner, double side_areas[6])
{
    if (n < 0 || (n > 0 && (!centroid || !area)) || !is_periodic || !canon_pos || !partner || !side_areas) return EU_ERR_ARG;
    const double pi = 3.14159265358979323846264338327950288;
    const double area_tol = 1e-6, centroid_tol = 1e-6;
    double low[3] = { 1e100, 1e100, 1e100 }, hi[3] = { -1e100, -1e100, -1e100 };
    for (int i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) {
            low[d] = std::min(low[d], centroid[3*size_t(i) + d]);
            hi[d] = std::max(hi[d], centroid[3*size_t(i) + d]);
        }
    }
    for (int k = 0; k < 6; ++k) side_areas[k] = 0.0;
    const size_t nn = size_t(n);
    std::vector<double> cent(3*nn), key(nn);
    for (int i = 0; i < n; ++i) {
        int cp = -1;
        for (int d = 0; d < 3; ++d) {
            const double coord = centroid[3*size_t(i) + d];
            if (std::fabs(coord - low[d]) <= spatial_tolerance) { cp = 2*d; break; }
            if (std::fabs(coord - hi[d]) <= spatial_tolerance) { cp = 2*d + 1; break; }
        }
        if (cp < 0) return EU_ERR_ARG;       // "Boundary face centroid not on bounding box" (:143-148)
        canon_pos[i] = cp;
        partner[i] = -1;
        side_areas[cp] += area[i];
        for (int d = 0; d < 3; ++d) cent[3*size_t(i) + d] = centroid[3*size_t(i) + d];
        cent[3*size_t(i) + cp/2] = 0.0;
        key[size_t(i)] = cent[3*size_t(i) + (cp/2 + 1)%3] + pi*cent[3*size_t(i) + (cp/2 + 2)%3];
    }
    std::vector<int> order(nn);
    for (int i = 0; i < n; ++i) order[size_t(i)] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[size_t(a)] < key[size_t(b)]; });
    std::vector<double> sorted_key(nn);
    for (int i = 0; i < n; ++i) sorted_key[size_t(i)] = key[size_t(order[size_t(i)])];
    const double band = (1.0 + pi)*centroid_tol*1.0000001;
    for (int pos = 0; pos < n; ++pos) {
        const int i = order[size_t(pos)];
        if (partner[i] != -1 || !is_periodic[canon_pos[i]]) continue;
        const int target = canon_pos[i] ^ 1;
        const size_t lo = size_t(std::lower_bound(sorted_key.begin(), sorted_key.end(), key[size_t(i)] - band) - sorted_key.begin());
        for (size_t q = lo; q < size_t(n) && sorted_key[q] <= key[size_t(i)] + band; ++q) {
            const int j = order[q];
            if (canon_pos[j] != target || partner[j] != -1) continue;
            if (std::fabs(area[i] - area[j]) > area_tol) continue;
            double d2 = 0.0;
            for (int d = 0; d < 3; ++d) { const double v = cent[3*size_t(j) + d] - cent[3*size_t(i) + d]; d2 += v*v; }
            if (std::sqrt(d2) <= centroid_tol) { partner[i] = j; partner[j] = i; break; }
        }
    }
    return EU_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Callers and data formats on either side of the path, host only.
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

// GridInterfaceEuler::buildFaceIndices (common/GridInterfaceEuler.hpp:498-611) over the flat CSR adjacency: unique face
// numbers per half-face.  Interior faces are numbered in the order the cell walk discovers them (a face is discovered by
// the cell visited first, :531-541), boundary faces after all interior ones in walk order (:573-575); a half-face of
// the later cell finds its number by the FIRST entry of the earlier cell's discovered faces that names it (:588-593,
// std::find -- two faces between the same pair of cells share the first one's number, as in the reference).  The
// reference's search walks a growing table (its comment: "potentially *VERY* expensive"); here each cell's discovered
// faces are a short contiguous range, O(half-faces x faces per cell).  Cell index == iteration order (as for CpGrid).
int eu_build_face_indices(int n_cells, const int* hf_offset, const int* hf_neighbour, int* face_index, int* num_faces,
                          int* max_faces_per_cell)
{
    if (n_cells < 0 || !hf_offset || !hf_neighbour || !face_index) return EU_ERR_ARG;
    std::vector<int> faces;                       // neighbour cell of every discovered interior face
    std::vector<int> fpos(size_t(n_cells) + 1, 0);
    int max_ncf = 0;
    for (int c = 0; c < n_cells; ++c) {           // first pass (:524-546)
        for (int h = hf_offset[c]; h < hf_offset[c + 1]; ++h) {
            const int c1 = hf_neighbour[h];
            if (c1 >= 0) {
                if (c1 >= n_cells || c1 == c) return EU_ERR_ARG;
                if (c1 > c) faces.push_back(c1);          // neighbour not visited yet: a new interior face
            }
        }
        fpos[size_t(c) + 1] = int(faces.size());
        max_ncf = std::max(max_ncf, hf_offset[c + 1] - hf_offset[c]);
    }
    int total = int(faces.size());
    for (int c = 0; c < n_cells; ++c) {           // second pass (:566-599)
        for (int h = hf_offset[c]; h < hf_offset[c + 1]; ++h) {
            const int c1 = hf_neighbour[h];
            if (c1 < 0) { face_index[h] = total++; continue; }
            const int t = std::min(c, c1), seek = std::max(c, c1);
            const int* b = faces.data() + fpos[size_t(t)];
            const int* e = faces.data() + fpos[size_t(t) + 1];
            const int* p = std::find(b, e, seek);
            if (p == e) return EU_ERR_ARG;                 // adjacency not symmetric
            face_index[h] = int(p - faces.data());
        }
    }
    if (num_faces) *num_faces = total;
    if (max_faces_per_cell) *max_faces_per_cell = max_ncf;
    return EU_OK;
}

// IncompFlowSolverHybrid::postProcessFluxes (mimetic/IncompFlowSolverHybrid.hpp:707-807): the out-fluxes of the two
// half-faces of a face are made exactly antisymmetric, f -> +-0.5 (f_first - f_second) in walk order; boundary faces
// are left alone unless they have a periodic partner (partner_face[face] >= 0; NULL = no periodic faces, :770-772),
// in which case the pair is treated like the two sides of one face.  Same accumulation order as FaceFluxes::put / get.
// Returns the largest modification in *max_modification.  This is the step that produces the hf_flux array handed to
// eu_transport_solve.
int eu_post_process_fluxes(int n_cells, const int* hf_offset, const int* hf_neighbour, const int* face_index, int num_faces,
                           const int* partner_face, double* hf_flux, double* max_modification)
{
    if (n_cells < 0 || !hf_offset || !hf_neighbour || !face_index || !hf_flux || num_faces < 0) return EU_ERR_ARG;
    std::vector<double> fluxes(size_t(num_faces), 0.0);
    std::vector<unsigned char> visited(size_t(num_faces), 0);
    double max_mod = 0.0;
    auto put = [&](double flux, int f) {
        fluxes[size_t(f)] += (visited[size_t(f)] ? -1.0 : 1.0)*flux;
        ++visited[size_t(f)];
    };
    auto get = [&](double& flux, int f) {
        const double nf = 0.5*(visited[size_t(f)] ? -1.0 : 1.0)*fluxes[size_t(f)];
        max_mod = std::max(max_mod, std::fabs(flux - nf));
        flux = nf;
        ++visited[size_t(f)];
    };
    const long long H = hf_offset[n_cells];
    for (long long h = 0; h < H; ++h) {
        const int f = face_index[h];
        if (f < 0 || f >= num_faces) return EU_ERR_ARG;
        if (hf_neighbour[h] < 0) {
            if (!partner_face) continue;
            const int pf = partner_face[f];
            if (pf != -1) { put(hf_flux[h], f); put(hf_flux[h], pf); }
        } else {
            put(hf_flux[h], f);
        }
    }
    std::fill(visited.begin(), visited.end(), 0);
    for (long long h = 0; h < H; ++h) {
        const int f = face_index[h];
        if (hf_neighbour[h] < 0) {
            if (!partner_face) continue;
            const int pf = partner_face[f];
            if (pf != -1) {
                get(hf_flux[h], f);
                double dummy = hf_flux[h];
                get(dummy, pf);
            }
        } else {
            get(hf_flux[h], f);
        }
    }
    if (max_modification) *max_modification = max_mod;
    return EU_OK;
}

// writeField (common/SimulatorUtilities.hpp:288-298): the saturation file the drivers write after every step
// (SimulatorBase / SimulatorTester: "<prefix>-<step>.sat"): the size, then one value per line in the stream's default
// formatting.  Returns EU_ERR_ARG when the file cannot be opened (the reference throws).
int eu_write_field(const double* field, long long n, const char* filename)
{
    if (!field || n < 0 || !filename) return EU_ERR_ARG;
    std::ofstream os(filename);
    if (!os) return EU_ERR_ARG;
    os << (unsigned long)n << '\n';
    std::copy(field, field + n, std::ostream_iterator<double>(os, "\n"));
    return os ? EU_OK : EU_ERR_ARG;
}

} // extern "C"


// ---- work units of the box kernel (host logic of eu_tile.cuh's sweep; no device needed) -------------------------------
//
// Work units: tiles x z-ranges of the own planes [z_lo, z_hi), listed block after block (block i sweeps
// units[start[i] .. start[i+1]), in that order).  With a neighbour rank below / above, the unit of a tile that contains
// the bnd_lo / bnd_hi planes whose cells the neighbour keeps as ghosts is flagged (bit 0 / bit 1) and comes first in its
// block's list: the kernel pushes those cells' results to the neighbour as it sweeps them and counts the finished flagged
// units -- ONE per tile and boundary, which is what the exchange's counters are set to (eu_box_plan_units).
//
//   chunks (spans == 0)  every tile's planes cut into the same number of z-chunks, handed out round-robin.  The chunk
//           length balances (units per block) x (planes + 1 prologue step per unit); lz > 0 fixes the length.  A block
//           may do one unit more than another, but the blocks of one SM share its issue slots, so what counts is the sum
//           per SM, and that differs by one unit in ~40
//   spans   the (tile, plane) pairs in tile-major order cut into one span of equal length per block; a span is split into
//           units at tile boundaries, and a cut closer than min_piece planes to a tile boundary moves onto it (so the
//           planes next to a slab boundary stay in ONE unit per tile).  Single-rank runs only.
//           Measured (profiles/README.md, r04a): 512x512x256 2.6 % slower than chunks (neighbouring tiles are no longer
//           swept at the same time: less halo reuse in L2), 32-plane slab 1.5 % faster, C3 0.7 % faster
namespace {

int box_chunks(int tiles, int planes, int grid_blocks, int lz, int min_len)
{
    int best_chunks = 1;
    double best = 1e300;
    for (int chunks = 1; chunks <= planes; ++chunks) {
        const double len = double(planes)/chunks;
        if (len < std::max(min_len, 1) && chunks > 1) break;
        if (lz > 0) { if (len <= lz || chunks == planes) { best_chunks = chunks; break; } continue; }
        if (len > 64.0) continue;
        const double rounds = double(((long long)chunks*tiles + grid_blocks - 1)/grid_blocks);
        const double cost = rounds*(len + 1.3);
        if (cost < best - 1e-9) { best = cost; best_chunks = chunks; }
    }
    return best_chunks;
}

} // namespace

int eu_box_make_units(int nx, int ny, int tx, int ty, int z_lo, int z_hi, int bnd_lo, int bnd_hi, int grid_blocks, int spans,
                      int lz, std::vector<EuBoxUnit>& units, std::vector<int>& start)
{
    units.clear();
    start.clear();
    if (nx < 1 || ny < 1 || tx < 1 || ty < 1 || grid_blocks < 1 || bnd_lo < 0 || bnd_hi < 0) return -1;
    const int tiles_x = (nx + tx - 1)/tx, tiles_y = (ny + ty - 1)/ty, tiles = tiles_x*tiles_y;
    const int planes = z_hi - z_lo;
    if (planes <= 0) return 0;
    if (std::max(bnd_lo, bnd_hi) > planes) return -1;
    auto unit_of = [&](int tile, int s, int e2) {
        const int txi = tile % tiles_x, tyi = tile/tiles_x;
        EuBoxUnit un;
        un.xy = (txi*tx) | ((tyi*ty) << 16);
        un.z0 = z_lo + s;
        un.z1 = z_lo + e2;
        un.flags = ((bnd_lo > 0 && s == 0) ? 1 : 0) | ((bnd_hi > 0 && e2 == planes) ? 2 : 0);
        return un;
    };
    if (!spans || lz > 0 || bnd_lo > 0 || bnd_hi > 0) {
        const int chunks = box_chunks(tiles, planes, grid_blocks, lz, std::max(bnd_lo, bnd_hi));
        std::vector<EuBoxUnit> flat;
        auto add = [&](int q) {
            const int s = int((long long)planes*q/chunks), e2 = int((long long)planes*(q + 1)/chunks);
            if (e2 <= s) return;
            for (int tile = 0; tile < tiles; ++tile) flat.push_back(unit_of(tile, s, e2));
        };
        // flagged chunks first
        if (bnd_lo > 0) add(0);
        if (bnd_hi > 0 && (chunks > 1 || bnd_lo == 0)) add(chunks - 1);
        for (int q = 0; q < chunks; ++q) {
            if ((bnd_lo > 0 && q == 0) || (bnd_hi > 0 && q == chunks - 1)) continue;
            add(q);
        }
        const int blocks = std::max(1, std::min(grid_blocks, int(flat.size())));
        for (int i = 0; i < blocks; ++i) {
            start.push_back(int(units.size()));
            for (size_t u = size_t(i); u < flat.size(); u += size_t(blocks)) units.push_back(flat[u]);
        }
        start.push_back(int(units.size()));
    } else {
        const long long W = (long long)tiles*planes;
        const int min_piece = 3;
        const int blocks = int(std::max(1LL, std::min((long long)grid_blocks, W/4)));
        std::vector<long long> cut(size_t(blocks) + 1);
        for (int i = 0; i <= blocks; ++i) {
            const long long pos = W*i/blocks;
            long long tile = pos/planes;
            int s = int(pos % planes);
            if (s < min_piece) s = 0;
            else if (s > planes - min_piece) { s = 0; ++tile; }
            cut[size_t(i)] = tile*planes + s;
        }
        cut[0] = 0; cut[size_t(blocks)] = W;
        for (int i = 0; i < blocks; ++i) {
            start.push_back(int(units.size()));
            for (long long pos = cut[size_t(i)]; pos < cut[size_t(i) + 1]; ) {
                const int tile = int(pos/planes), s = int(pos % planes);
                const int e2 = int(std::min((long long)planes, s + (cut[size_t(i) + 1] - pos)));
                units.push_back(unit_of(tile, s, e2));
                pos += e2 - s;
            }
        }
        start.push_back(int(units.size()));
    }
    // the exchange counts ONE finished unit per tile and boundary
    int n0 = 0, n1 = 0;
    for (const EuBoxUnit& un : units) { if (un.flags & 1) ++n0; if (un.flags & 2) ++n1; }
    if ((bnd_lo > 0 && n0 != tiles) || (bnd_hi > 0 && n1 != tiles)) return -1;
    return int(units.size());
}

extern "C" int eu_debug_box_units(int nx, int ny, int tx, int ty, int z_lo, int z_hi, int bnd_lo, int bnd_hi, int grid_blocks,
                                  int spans, int lz, int* units4, int max_units, int* start, int max_start, int* n_blocks)
{
    std::vector<EuBoxUnit> u;
    std::vector<int> st;
    const int n = eu_box_make_units(nx, ny, tx, ty, z_lo, z_hi, bnd_lo, bnd_hi, grid_blocks, spans, lz, u, st);
    if (n < 0) return -1;
    if (n_blocks) *n_blocks = st.empty() ? 0 : int(st.size()) - 1;
    if (units4) {
        if (n > max_units) return -1;
        for (int i = 0; i < n; ++i) { units4[4*i] = u[size_t(i)].xy; units4[4*i + 1] = u[size_t(i)].z0; units4[4*i + 2] = u[size_t(i)].z1; units4[4*i + 3] = u[size_t(i)].flags; }
    }
    if (start) {
        if (int(st.size()) > max_start) return -1;
        std::copy(st.begin(), st.end(), start);
    }
    return n;
}
