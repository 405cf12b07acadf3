// Box kernel of FAST mode: plane sweep over tiles with TMA-staged operands.  Included by eu_fast.cu (shares its rock-curve
// tables, face arithmetic and finish_cell).
//
// Applies when the local numbering is a box, c = x + nx*(y + ny*z) over whole planes (eu_api.cu decides; Cartesian and
// corner-point grids in natural order, z-slab ranks with ghost planes).  The regular faces of such a grid are the three
// axis planes of the face arrays (eu_setup.cu, k_assign_fid): face (c, c+1) at c, face (c, c+nx) at n + c, face
// (c, c+nx*ny) at 2n + c, zero where a cell has no such neighbour; everything else (boundary, fault, periodic wrap) is a
// per-cell bit mask into the SELL records and is added face by face after the regular ones.
//
// A thread block owns a tile of tx*ty cells (one per thread) and sweeps it along z over a chunk of planes (a work unit).
// Per plane k one bundle of operands arrives in shared memory by TMA (cp.async.bulk.tensor, mbarrier completion),
// two or three planes ahead of its use:
//     S(k+1)  [and pc(k+1)]   the tile plus a halo (2 cells in x, 1 in y)         3-D box  (tx+4) x (ty+2) x 1
//     q(k) of the x-, y- and z-faces, G(k) of the axes that have gravity [, T(k)]  4-D boxes (tx+2) x (ty+1) x 1 x 1
// (a box of 8-byte elements must start at an even innermost coordinate -- 16 bytes; an odd one raises an illegal-
// instruction fault, tools/probe/tma_probe.cu -- hence the second halo column in x, which is never read)
// Out-of-range coordinates (halo outside the grid, plane -1 or nz) are zero-filled by the TMA unit: a zero face carries no
// flux and the saturation on its far side is never looked at with a non-zero weight.
// Step k of the sweep:
//   A  every thread evaluates the rock curves ONCE for its own cell of plane k+1 and, for the first 2(tx+ty) threads,
//      for one halo cell, and stores {lambda_w, lambda_o} [S, pc] in a ring of three tile buffers in shared memory;
//   -- __syncthreads --                       (the only block-wide synchronisation of a plane)
//   B  the four lateral faces of the cell of plane k from its neighbours' ring entries, the z+ face from the registers
//      (plane k+1 of the own column); the flux of that face is carried over as the z- face of the next plane, like in
//      the warp march of k_fast_step; irregular faces from the records; explicit update, range check, pc of the new
//      state (finish_cell).
// Against the warp march this evaluates the curves 1.3 instead of 5 times per cell, issues no address arithmetic or
// prefetches for its operands (one thread programs the TMA unit) and reads every face pair from shared memory.
#include <cuda.h>

struct EuBoxDev {
    int nx, ny, nz;               // local box
    int tx, ty;                   // tile
    int n_units;
    const int4* units;            // {x0 | y0 << 16, z0, z1 (exclusive), flags: bit 0 / 1 = holds cells of halo range A / B (pushed to a peer)}
    int stages;                   // bundles in flight
    int off_bar, off_lam, off_rk, off_stage;      // byte offsets in dynamic shared memory (from the 128-aligned base)
    int lam_bytes, rk_bytes, stage_bytes;
    int off_S, off_pc, off_q, off_G, off_T;       // inside a stage
    int off_A, off_V, A_bytes;                    // inside a stage: tx x ty boxes of acc_irr and 1 / pore volume of plane k
    int g_mask;                                   // bit a: the faces of axis plane a have a gravity component somewhere (G box loaded)
    int off_fx, off_fy, fx_bytes, fy_bytes;       // SHARE: two buffers each of x+ fluxes [ty][tx+1] and y+ fluxes [ty+1][tx]
    int T_bytes;                                  // one axis' box of q, G or T
    int n_flagged;                // units with a push flag
    const int* unit_start;        // [blocks + 1] block i sweeps units[unit_start[i] .. unit_start[i+1]), flagged ones first
};

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, int x, int y, int z, int w, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(w), "r"(bar) : "memory");
}

// Programmatic dependent launch (EU_PDL=1, off by default: measured slower, see launch_box): the next kernel of the
// stream may start -- take its SM slots, copy the rock tables to shared memory, set up its mbarriers -- while this one
// drains; it touches nothing a substep writes before pdl_wait(), which returns when the whole previous grid has finished
// and its stores are visible.  Without the launch attribute both instructions do nothing.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// G == 0 (no gravity component through the face: every lateral face of a grid with horizontal layers): both phases are
// upwinded by the sign of q alone.  Bit-identical to face_regular for G == 0.
template <bool OWN, bool CAP>
__device__ __forceinline__ double face_no_gravity(double q, double lw0, double lo0, double lw1, double lo1, double cap_coef, double Tdpc)
{
    const bool from_self = (q >= 0.0) == OWN;
    const double cw = from_self ? lw0 : lw1, co = from_self ? lo0 : lo1;
    double dS = div_pos(cw*q, cw + co);
    if (CAP) dS = fma(cap_coef, Tdpc, dS);
    return dS;
}

// one regular face of the box kernel (method_viscous is on: the host sends other runs to the slice-class kernel).
// has_G: the face's axis plane has a gravity component somewhere (uniform over the launch); without it -- every lateral
// face of a horizontally layered grid -- no G is loaded at all and both phases are upwinded by the sign of q.
template <bool ROCKS, bool MULTIROCK, bool CAP, bool OWN>
__device__ __forceinline__ double box_face(const TabLayout& L, const EuTablesDev& t, const MarchCarry& m, double lw1, double lo1,
                                           double S1, double pc1, int rk1, double q, double G, bool has_G, double T)
{
    double cap_coef = 0.0, Tdpc = 0.0;
    if (CAP) {
        cap_coef = cap_coefficient<ROCKS, MULTIROCK>(L, t, m.rock0, rk1, m.S0, S1);
        Tdpc = T*(OWN ? (pc1 - m.pc0) : (m.pc0 - pc1));
    }
    if (!has_G) return face_no_gravity<OWN, CAP>(q, m.lw0, m.lo0, lw1, lo1, cap_coef, Tdpc);
    return face_regular<OWN, CAP>(q, q, G, m.lw0, m.lo0, lw1, lo1, 1, cap_coef, Tdpc);
}

// the faces of a cell outside the axis planes (boundary, fault, periodic wrap), from the SELL records: bit j of mask
template <bool ROCKS, bool MULTIROCK, bool CAP>
__device__ __forceinline__ double box_record_faces(const TabLayout& L, const EuGridDev& g, const EuTablesDev& t, const EuFastDev& f,
                                                   const EuStepArgs& a, unsigned mask, int c, const MarchCarry& m)
{
    const int2* __restrict__ recp = f.rec + f.slice_base[c >> 5] + (c & 31);
    double acc = 0.0;
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1u;
        const int2 r = recp[j*EU_SLICE];
        if (r.x == EU_REC_PAD) continue;
        const bool interior = r.x >= 0;
        const bool own = !interior || c < r.x;
        const double2 qg = f.qg[r.y];
        const double S1 = interior ? a.S_in[r.x] : g.bnd_sat[-2 - r.x];
        const int rk = (MULTIROCK && interior) ? f.rock8[r.x] : m.rock0;
        double lw1, lo1;
        Mob<ROCKS, MULTIROCK>::both(L, t, rk, S1, lw1, lo1);
        double cap_coef = 0.0, Tdpc = 0.0;
        if (CAP && interior) {
            cap_coef = cap_coefficient<ROCKS, MULTIROCK>(L, t, m.rock0, rk, m.S0, S1);
            const double pc1 = a.pc_in[r.x];
            Tdpc = f.T[r.y]*(own ? (pc1 - m.pc0) : (m.pc0 - pc1));
        }
        acc += face_contribution<CAP>(own, interior, qg.x, qg.x, qg.y, m.lw0, m.lo0, lw1, lo1, 1, a.method_gravity, cap_coef, Tdpc);
    }
    return acc;
}

// Pre-pass of a substep in box mode: the faces outside the axis planes (boundary, fault, periodic wrap), one thread per
// cell that has any (compact list built at upload).  Writes the sum of their contributions to the cell's residual;
// k_box_step adds it to the regular faces.  A few per cent of the cells: taking these faces out of the sweep keeps
// the sweep's warps in step (a single lane with a boundary face used to hold back its whole block at the barrier).
template <bool ROCKS, bool MULTIROCK, bool CAP>
__global__ void __launch_bounds__(256) k_box_irregular(EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a, EuHaloDev halo,
                                                       const int* __restrict__ cells, int n, const unsigned short* __restrict__ cmask,
                                                       double* __restrict__ acc_irr)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.nbd = double(t.n_buckets);
    pdl_launch_dependents();
    if (ROCKS) tables_to_smem(t);
    pdl_wait();
    if (halo.enabled && threadIdx.x < halo.n_wait) {
        // some of these faces may look at ghost cells: the previous substep's pushes must have landed
        const volatile unsigned* fl = halo.my_flags + halo.wait_rank[threadIdx.x];
        const long long t0 = clock64();
        while ((int)(*fl - (halo.epoch - 1u)) < 0) {
            __nanosleep(100);
            if (clock64() - t0 > halo.timeout_cycles) { atomicExch(halo.err_flag, 1); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
        const int c = __ldg(cells + i);
        MarchCarry m;
        m.S0 = a.S_in[c];
        m.rock0 = MULTIROCK ? f.rock8[c] : 0;
        m.pc0 = CAP ? a.pc_in[c] : 0.0;
        m.dS4 = 0.0;
        Mob<ROCKS, MULTIROCK>::both(L, t, m.rock0, m.S0, m.lw0, m.lo0);
        acc_irr[c] = box_record_faces<ROCKS, MULTIROCK, CAP>(L, g, t, f, a, unsigned(__ldg(cmask + c)), c, m);
    }
}

// SHARE (used with the capillary term, whose faces are expensive): every lateral face is evaluated ONCE per tile -- a
// thread evaluates the x+, y+ and z+ faces of its cell and leaves the x+ / y+ fluxes in shared memory for its right / upper
// neighbour; the faces on the tile's low x / y edges are evaluated by tx + ty threads on the side.  The cell of plane k is
// then finished one step later (after the next barrier, when its neighbours' fluxes are visible): still one barrier per
// plane, 3 + (tx + ty)/(tx ty) face evaluations per cell instead of 5.
template <bool ROCKS, bool MULTIROCK, bool CAP, int NS, int MINB, bool SHARE>
__global__ void __launch_bounds__(256, MINB)
k_box_step(const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapPc,
           const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapG,
           const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapA,
           const __grid_constant__ CUtensorMap mapV,
           EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a, EuHaloDev halo, EuBoxDev b, int slice_lo, int slice_hi, int tab_bytes)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.nbd = double(t.n_buckets);
    pdl_launch_dependents();
    if (ROCKS) tables_to_smem(t);
    // 128-aligned base of the staging area behind the tables
    const unsigned base_u32 = (smem_u32(eu_smem) + unsigned(tab_bytes) + 127u) & ~127u;
    unsigned char* const base = eu_smem + (base_u32 - smem_u32(eu_smem));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(base_u32 + b.off_bar + 8*s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                                 // from here on the previous kernel's results are read
    if (!halo.enabled) {
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    constexpr int E = 16;                                       // bytes per ring entry: {lw, lo}; with the capillary term a second
    const int oB = b.lam_bytes/2;                               // array {S, pc} follows at oB (two 16-byte arrays: conflict-free LDS.128)
    const int tx = b.tx, ty = b.ty, txp = tx + 2, txf = tx + 1, txs = tx + 4;      // row lengths: ring / T boxes, face boxes, S boxes
    const int lx = tid % tx, ly = tid/tx;
    const bool in_tile = tid < tx*ty;
    // halo duty: 2(tx+ty) threads each own one halo cell of the ring buffers.  With SHARE the tx+ty threads of the edge
    // faces sit in other warps (the block waits for its slowest warp at every barrier: no warp gets both extras)
    const int n_halo = 2*(tx + ty), n_edge = tx + ty, n_thr = int(blockDim.x);
    const int e0 = SHARE ? 0 : 0;                                // edge faces: threads e0 .. e0 + tx + ty - 1
    const int h0 = (SHARE && ((n_edge + 31) & ~31) + n_halo <= n_thr) ? ((n_edge + 31) & ~31) : 0;
    const int ht = tid - h0;                                    // halo duty: threads h0 .. h0 + 2(tx+ty) - 1
    int hx = 0, hy = 0;
    bool has_halo = ht >= 0;
    if (ht < tx)                 { hx = ht;            hy = -1; }
    else if (ht < 2*tx)          { hx = ht - tx;       hy = ty; }
    else if (ht < 2*tx + ty)     { hx = -1;            hy = ht - 2*tx; }
    else if (ht < 2*tx + 2*ty)   { hx = tx;            hy = ht - 2*tx - ty; }
    else has_halo = false;
    // byte offsets of this thread's entries (ring buffers, S / pc boxes, face boxes, T boxes)
    const int o_ring = ((ly + 1)*txp + lx + 1)*E, o_ring_h = ((hy + 1)*txp + hx + 1)*E;
    const int o_rk = (ly + 1)*txp + lx + 1, o_rk_h = (hy + 1)*txp + hx + 1;
    const int o_S = ((ly + 1)*txs + lx + 2)*8, o_S_h = ((hy + 1)*txs + hx + 2)*8;
    const int o_T = (ly*txp + lx)*8;
    const int o_A = (ly*tx + lx)*8;
    const int D = b.nx*b.ny;
    const int n_G = __popc(unsigned(b.g_mask) & 7u);
    const unsigned bundle_bytes = unsigned(txs*(ty + 2)*8*(CAP ? 2 : 1) + (3 + n_G + (CAP ? 3 : 0))*txp*(ty + 1)*8 + 2*tx*ty*8);
    const bool hgx = b.g_mask & 1, hgy = b.g_mask & 2, hgz = b.g_mask & 4;        // axis has gravity: its G box is loaded
    const int oGx = b.off_G, oGy = b.off_G + (hgx ? b.T_bytes : 0), oGz = oGy + (hgy ? b.T_bytes : 0);     // G boxes present are packed
    const unsigned char* const stage0 = base + b.off_stage;
    unsigned char* const ring0 = base + b.off_lam;
    unsigned char* const rk0 = base + b.off_rk;
    int slot = 0;                                               // stage of the current bundle; par = phase parity of its mbarrier
    unsigned par = 0;
    const int u_begin = __ldg(b.unit_start + blockIdx.x), u_end = __ldg(b.unit_start + blockIdx.x + 1);
    if (halo.enabled && u_begin < u_end && (__ldg(&b.units[u_begin].w) & 3) && tid < halo.n_wait) {
        // this block sweeps planes next to a slab boundary: the ghosts of the previous substep must have landed before
        // any TMA load reads them (flagged units are the first of a block's list, so every block that has one gets here)
        const volatile unsigned* fl = halo.my_flags + halo.wait_rank[tid];
        const long long t0 = clock64();
        while ((int)(*fl - (halo.epoch - 1u)) < 0) {
            __nanosleep(100);
            if (clock64() - t0 > halo.timeout_cycles) { atomicExch(halo.err_flag, 1); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    // ---- producer (thread 0): one bundle per plane, NS - 1 planes ahead of the sweep, ACROSS work units -- the first
    // bundles of the next unit are already in flight while the last planes of the current one are swept
    // (its state lives in shared memory, behind the mbarriers: no registers of the sweep are spent on it)
    struct Producer { int4 unit, next; int pu, pk, slot, end; };
    Producer* const P = reinterpret_cast<Producer*>(base + b.off_bar + 64);
    auto produce = [&]() {
        int pu = P->pu;
        const int pend = P->end;
        if (pu >= pend) return;
        const int4 un = P->unit;
        const int x0 = un.x & 0xffff, y0 = un.x >> 16, k = P->pk, ps = P->slot;
        const unsigned bar = base_u32 + b.off_bar + 8*ps;
        const unsigned st = base_u32 + b.off_stage + unsigned(ps)*unsigned(b.stage_bytes);
        mbar_expect_tx(bar, bundle_bytes);
        tma_load_3d(st + b.off_S, &mapS, x0 - 2, y0 - 1, k + 1, bar);
        if (CAP) tma_load_3d(st + b.off_pc, &mapPc, x0 - 2, y0 - 1, k + 1, bar);
        tma_load_4d(st + b.off_q, &mapQ, x0 - 2, y0, k, 0, bar);
        tma_load_4d(st + b.off_q + b.T_bytes, &mapQ, x0, y0 - 1, k, 1, bar);
        tma_load_4d(st + b.off_q + 2*b.T_bytes, &mapQ, x0, y0, k, 2, bar);
        if (hgx) tma_load_4d(st + oGx, &mapG, x0 - 2, y0, k, 0, bar);
        if (hgy) tma_load_4d(st + oGy, &mapG, x0, y0 - 1, k, 1, bar);
        if (hgz) tma_load_4d(st + oGz, &mapG, x0, y0, k, 2, bar);
        if (CAP) {
            tma_load_4d(st + b.off_T, &mapT, x0 - 2, y0, k, 0, bar);
            tma_load_4d(st + b.off_T + b.T_bytes, &mapT, x0, y0 - 1, k, 1, bar);
            tma_load_4d(st + b.off_T + 2*b.T_bytes, &mapT, x0, y0, k, 2, bar);
        }
        tma_load_3d(st + b.off_A, &mapA, x0, y0, k, bar);
        tma_load_3d(st + b.off_V, &mapV, x0, y0, k, bar);
        P->slot = (ps + 1 == NS) ? 0 : ps + 1;
        if (k + 1 >= un.z) {                                    // on to the block's next unit
            ++pu;
            P->pu = pu;
            const int4 nx = P->next;
            P->unit = nx;
            P->pk = nx.y - 1;
            if (pu + 1 < pend) P->next = __ldg(b.units + pu + 1);
        } else {
            P->pk = k + 1;
        }
    };
    if (tid == 0) {
        const int pu = u_begin;
        P->pu = pu; P->slot = 0; P->end = u_end;
        int4 un = make_int4(0, 0, 0, 0);
        if (pu < u_end) un = __ldg(b.units + pu);
        P->unit = un;
        P->pk = un.y - 1;
        if (pu + 1 < u_end) P->next = __ldg(b.units + pu + 1);
        for (int j = 0; j < NS; ++j) produce();
    }
    bool first_step = true;                                     // nothing to refill at the block's very first step
    unsigned done0 = 0u, done1 = 0u;                            // finished units of this block with pushes to range A / B

    for (int u = u_begin; u < u_end; ++u) {
        const int4 unit = __ldg(b.units + u);
        const int x0 = unit.x & 0xffff, y0 = unit.x >> 16, z0 = unit.y, z1 = unit.z;
        const int push = unit.w & 3;
        __syncthreads();                                        // the ring buffers of the previous unit are free
        // ---- own column: the cell of plane z0-1 (operands of the first carried face)
        const int gx = x0 + lx, gy = y0 + ly;
        const bool active = in_tile && gx < b.nx && gy < b.ny;
        int c = gx + b.nx*gy + (z0 - 1)*D;                      // cell of plane k of this column
        MarchCarry m;
        m.S0 = 0.0; m.pc0 = 0.0; m.rock0 = 0; m.dS4 = 0.0;
        if (active && z0 > 0) {
            m.S0 = __ldg(a.S_in + c);
            if (MULTIROCK) m.rock0 = __ldg(f.rock8 + c);
            if (CAP) m.pc0 = __ldg(a.pc_in + c);
        }
        Mob<ROCKS, MULTIROCK>::both(L, t, m.rock0, m.S0, m.lw0, m.lo0);
        // halo cell: its cell index in plane k (or -1 outside the grid), for the rock id
        int hc = 0;                                             // (negative in plane -1: validity is a flag of its own)
        bool h_ok = false;
        if (MULTIROCK && has_halo) {
            const int qx = x0 + hx, qy = y0 + hy;
            h_ok = qx >= 0 && qx < b.nx && qy >= 0 && qy < b.ny;
            hc = qx + b.nx*qy + (z0 - 1)*D;
        }
        int rock_n = 0, rock_hn = 0;                            // rock ids of plane k+1: own cell, halo cell
        if (MULTIROCK && z0 < b.nz) {
            if (active) rock_n = __ldg(f.rock8 + c + D);
            if (h_ok) rock_hn = __ldg(f.rock8 + hc + D);
        }
        // ring buffers: next (plane k+1, written in phase A), cur (plane k, read in phase B), and the one in between
        int r_next = ((z0 % 3) + 3) % 3, r_cur = (r_next + 2) % 3;
        // SHARE: the cell of the previous plane, waiting for its neighbours' fluxes
        double p_acc = 0.0, p_S0 = 0.0, p_ipv = 0.0, p_pcs = 1.0, p_lw = 0.0, p_lo = 0.0;
        int p_rock = 0, fpar = 0;
        bool p_update = false;
        auto finish_prev = [&]() {
            // fluxes of the x- and y- faces: left by the left / lower neighbour (or the edge threads) a step ago
            const unsigned char* fxr = base + b.off_fx + (fpar ^ 1)*b.fx_bytes;
            const unsigned char* fyr = base + b.off_fy + (fpar ^ 1)*b.fy_bytes;
            const double acc = p_acc + *reinterpret_cast<const double*>(fxr + (ly*txf + lx)*8)
                                     + *reinterpret_cast<const double*>(fyr + (ly*tx + lx)*8);
            OwnMob<false> own0;
            own0.lw[0] = p_lw; own0.lo[0] = p_lo;
            double pcn;
            const int cp = c - D;
            const double sat = finish_cell<ROCKS, MULTIROCK, CAP, false>(L, t, f, a, cp, p_S0, p_rock, own0, p_ipv, acc, pcn, true, p_pcs);
            if (push) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (!((push >> r) & 1)) continue;
                    const int first = (r == 0 ? slice_lo : halo.b_lo)*EU_SLICE;
                    const int last = (r == 0 ? halo.a_hi : slice_hi)*EU_SLICE;
                    const int d = (cp >= first && cp < last) ? halo.dst[r][cp - first] : -1;
                    if (d >= 0) {
                        halo.peer_S[r][d] = sat;
                        if (CAP && halo.peer_pc[r]) halo.peer_pc[r][d] = pcn;
                    }
                }
            }
        };
        for (int k = z0 - 1; k < z1; ++k) {
            // the one operand of plane k that is not staged: requested now, used after the barrier
            const bool update = active && k >= z0;
            double pcs = 1.0;
            if (CAP && ROCKS && update) pcs = ldg_f64(f.pcscale + c);
            // rock ids of plane k+1 were requested a step ago; those of plane k+2 are requested now
            const int rock1 = rock_n, rock_h = rock_hn;
            if (MULTIROCK) {
                const bool up2 = k + 2 < b.nz;
                rock_n = (active && up2) ? int(__ldg(f.rock8 + c + 2*D)) : 0;
                rock_hn = (h_ok && up2) ? int(__ldg(f.rock8 + hc + 2*D)) : 0;
            }
            const unsigned char* st = stage0 + slot*b.stage_bytes;
            mbar_wait(base_u32 + b.off_bar + 8*slot, par);
            // (per-cell operands of plane k staged with the bundle, read in phase B: 1 / pore volume, and the sum of the
            // faces outside the axis planes -- k_box_irregular wrote it before this launch; zero for every other cell.
            // They used to be per-thread loads behind a per-cell mask fetched a plane ahead: the compiler turned the fresh
            // mask into the next step's predicate right behind the barrier and waited there for the load, 14 % of the
            // stall samples)
            // ---- phase A: rock curves of plane k+1, once per cell
            unsigned char* ring_next = ring0 + r_next*b.lam_bytes;
            double S1 = 0.0, pc1 = 0.0, lw1 = 0.0, lo1 = 0.0;
            if (in_tile) {
                S1 = *reinterpret_cast<const double*>(st + b.off_S + o_S);
                if (CAP) pc1 = *reinterpret_cast<const double*>(st + b.off_pc + o_S);
                Mob<ROCKS, MULTIROCK>::both(L, t, rock1, S1, lw1, lo1);
                *reinterpret_cast<double2*>(ring_next + o_ring) = make_double2(lw1, lo1);
                if (CAP) {
                    *reinterpret_cast<double2*>(ring_next + oB + o_ring) = make_double2(S1, pc1);
                    if (MULTIROCK) rk0[r_next*b.rk_bytes + o_rk] = (unsigned char)rock1;
                }
            }
            if (has_halo) {
                const double Sh = *reinterpret_cast<const double*>(st + b.off_S + o_S_h);
                double lwh, loh;
                Mob<ROCKS, MULTIROCK>::both(L, t, rock_h, Sh, lwh, loh);
                *reinterpret_cast<double2*>(ring_next + o_ring_h) = make_double2(lwh, loh);
                if (CAP) {
                    const double ph = *reinterpret_cast<const double*>(st + b.off_pc + o_S_h);
                    *reinterpret_cast<double2*>(ring_next + oB + o_ring_h) = make_double2(Sh, ph);
                    if (MULTIROCK) rk0[r_next*b.rk_bytes + o_rk_h] = (unsigned char)rock_h;
                }
            }
            __syncthreads();
            // the stage of the previous step is free now: refill it with the next bundle of the stream
            if (tid == 0 && !first_step) produce();
            first_step = false;
            // ---- phase B: faces of plane k
            if (SHARE) {
                unsigned char* fxw = base + b.off_fx + fpar*b.fx_bytes;
                unsigned char* fyw = base + b.off_fy + fpar*b.fy_bytes;
                const unsigned char* ringc = ring0 + r_cur*b.lam_bytes;
                const unsigned char* rkc = rk0 + r_cur*b.rk_bytes;
                if (in_tile) {
                    if (p_update) finish_prev();
                    const unsigned char* QQ = st + b.off_q + o_T;          // q, G and T boxes share their indexing
                    const unsigned char* TT = st + b.off_T + o_T;
                    const double q5 = *reinterpret_cast<const double*>(QQ + 2*b.T_bytes);
                    const double G5 = hgz ? *reinterpret_cast<const double*>(st + oGz + o_T) : 0.0;
                    const double T5 = CAP ? *reinterpret_cast<const double*>(TT + 2*b.T_bytes) : 0.0;
                    const double dS5 = box_face<ROCKS, MULTIROCK, CAP, true>(L, t, m, lw1, lo1, S1, pc1, rock1, q5, G5, hgz, T5);
                    if (update) {
                        // x+ and y+ faces: this cell is their lo cell
                        const double2 ex = *reinterpret_cast<const double2*>(ringc + o_ring + E);
                        const double2 ey = *reinterpret_cast<const double2*>(ringc + o_ring + txp*E);
                        double2 sx = make_double2(0.0, 0.0), sy = make_double2(0.0, 0.0);
                        double Txp = 0.0, Typ = 0.0;
                        int rxp = 0, ryp = 0;
                        if (CAP) {
                            sx = *reinterpret_cast<const double2*>(ringc + oB + o_ring + E);
                            sy = *reinterpret_cast<const double2*>(ringc + oB + o_ring + txp*E);
                            Txp = *reinterpret_cast<const double*>(TT + 16);
                            Typ = *reinterpret_cast<const double*>(TT + b.T_bytes + txp*8);
                            if (MULTIROCK) { rxp = int(rkc[o_rk + 1]); ryp = int(rkc[o_rk + txp]); }
                        }
                        const double qxp = *reinterpret_cast<const double*>(QQ + 16);
                        const double qyp = *reinterpret_cast<const double*>(QQ + b.T_bytes + txp*8);
                        const double Gxp = hgx ? *reinterpret_cast<const double*>(st + oGx + o_T + 16) : 0.0;
                        const double Gyp = hgy ? *reinterpret_cast<const double*>(st + oGy + o_T + txp*8) : 0.0;
                        const double dSx = box_face<ROCKS, MULTIROCK, CAP, true>(L, t, m, ex.x, ex.y, sx.x, sx.y, rxp, qxp, Gxp, hgx, Txp);
                        const double dSy = box_face<ROCKS, MULTIROCK, CAP, true>(L, t, m, ey.x, ey.y, sy.x, sy.y, ryp, qyp, Gyp, hgy, Typ);
                        *reinterpret_cast<double*>(fxw + (ly*txf + lx + 1)*8) = dSx;
                        *reinterpret_cast<double*>(fyw + ((ly + 1)*tx + lx)*8) = dSy;
                        p_acc = ((m.dS4 - dS5) - dSx) - dSy + *reinterpret_cast<const double*>(st + b.off_A + o_A);
                        p_S0 = m.S0; p_ipv = *reinterpret_cast<const double*>(st + b.off_V + o_A); p_pcs = pcs; p_lw = m.lw0; p_lo = m.lo0; p_rock = m.rock0;
                    }
                    p_update = update;
                    m.dS4 = dS5;
                }
                // the faces on the tile's low edges: thread i < tx the y- face of cell (i, 0), thread tx + j the x- face of (0, j)
                if (tid >= e0 && tid < e0 + n_edge && k >= z0) {
                    const int et = tid - e0;
                    const bool yedge = et < tx;
                    const int i = yedge ? et : 0, j = yedge ? 0 : et - tx;
                    const int hi_idx = (j + 1)*txp + i + 1;                       // ring entry of the tile cell (the face's hi cell)
                    const int lo_idx = yedge ? hi_idx - txp : hi_idx - 1;       // halo cell below / to the left (the lo cell)
                    const double2 el = *reinterpret_cast<const double2*>(ringc + lo_idx*E);
                    const double2 eh = *reinterpret_cast<const double2*>(ringc + hi_idx*E);
                    MarchCarry ml;
                    ml.lw0 = el.x; ml.lo0 = el.y; ml.S0 = 0.0; ml.pc0 = 0.0; ml.rock0 = 0; ml.dS4 = 0.0;
                    double2 sh = make_double2(0.0, 0.0);
                    double Tf = 0.0;
                    int rh = 0;
                    if (CAP) {
                        const double2 sl = *reinterpret_cast<const double2*>(ringc + oB + lo_idx*E);
                        sh = *reinterpret_cast<const double2*>(ringc + oB + hi_idx*E);
                        ml.S0 = sl.x; ml.pc0 = sl.y;
                        if (MULTIROCK) { ml.rock0 = int(rkc[lo_idx]); rh = int(rkc[hi_idx]); }
                        Tf = yedge ? *reinterpret_cast<const double*>(st + b.off_T + b.T_bytes + i*8)
                                   : *reinterpret_cast<const double*>(st + b.off_T + (j*txp + 1)*8);
                    }
                    const int eo = yedge ? i*8 : (j*txp + 1)*8;           // entry of the edge face in its q / G / T box
                    const double qe = *reinterpret_cast<const double*>(st + b.off_q + (yedge ? b.T_bytes : 0) + eo);
                    const bool ge = yedge ? hgy : hgx;
                    const double Ge = ge ? *reinterpret_cast<const double*>(st + (yedge ? oGy : oGx) + eo) : 0.0;
                    const double dSe = box_face<ROCKS, MULTIROCK, CAP, true>(L, t, ml, eh.x, eh.y, sh.x, sh.y, rh, qe, Ge, ge, Tf);
                    if (yedge) *reinterpret_cast<double*>(fyw + i*8) = dSe;
                    else       *reinterpret_cast<double*>(fxw + j*txf*8) = dSe;
                }
                fpar ^= 1;
            } else
            if (in_tile) {
                const unsigned char* QQ = st + b.off_q + o_T;              // q, G and T boxes share their indexing
                const unsigned char* TT = st + b.off_T + o_T;
                const double q5 = *reinterpret_cast<const double*>(QQ + 2*b.T_bytes);
                const double G5 = hgz ? *reinterpret_cast<const double*>(st + oGz + o_T) : 0.0;
                const double T5 = CAP ? *reinterpret_cast<const double*>(TT + 2*b.T_bytes) : 0.0;
                const double dS5 = box_face<ROCKS, MULTIROCK, CAP, true>(L, t, m, lw1, lo1, S1, pc1, rock1, q5, G5, hgz, T5);
                if (update) {
                    const unsigned char* ring = ring0 + r_cur*b.lam_bytes + o_ring;
                    const unsigned char* rk = rk0 + r_cur*b.rk_bytes + o_rk;
                    double acc = m.dS4 - dS5;
                    // lateral faces x-, x+, y-, y+: neighbour entry, face pair, T -- one face at a time (short live ranges:
                    // the operands come from shared memory, their latency is covered by the other warps)
                    const int dr[4] = { -E, E, -txp*E, txp*E };
                    const int dk[4] = { -1, 1, -txp, txp };
                    const int dt[4] = { 8, 16, b.T_bytes, b.T_bytes + txp*8 };
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 e = *reinterpret_cast<const double2*>(ring + dr[q]);
                        double2 e2 = make_double2(0.0, 0.0);
                        double Tq = 0.0;
                        int rq = 0;
                        if (CAP) {
                            e2 = *reinterpret_cast<const double2*>(ring + oB + dr[q]);
                            rq = MULTIROCK ? int(rk[dk[q]]) : 0;
                            Tq = *reinterpret_cast<const double*>(TT + dt[q]);
                        }
                        const double qq = *reinterpret_cast<const double*>(QQ + dt[q]);
                        const bool gq = q < 2 ? hgx : hgy;
                        const double Gq = gq ? *reinterpret_cast<const double*>(st + (q < 2 ? oGx : oGy) + o_T + (dt[q] - (q < 2 ? 0 : b.T_bytes))) : 0.0;
                        if (q & 1) acc -= box_face<ROCKS, MULTIROCK, CAP, true>(L, t, m, e.x, e.y, e2.x, e2.y, rq, qq, Gq, gq, Tq);
                        else       acc += box_face<ROCKS, MULTIROCK, CAP, false>(L, t, m, e.x, e.y, e2.x, e2.y, rq, qq, Gq, gq, Tq);
                    }
                    acc += *reinterpret_cast<const double*>(st + b.off_A + o_A);
                    const double inv_pv = *reinterpret_cast<const double*>(st + b.off_V + o_A);
                    OwnMob<false> own0;
                    own0.lw[0] = m.lw0; own0.lo[0] = m.lo0;
                    double pcn;
                    const double sat = finish_cell<ROCKS, MULTIROCK, CAP, false>(L, t, f, a, c, m.S0, m.rock0, own0, inv_pv, acc, pcn, true, pcs);
                    if (push) {
                        // ghost slot of this cell in a neighbour rank (the tables cover the slices of the two ranges)
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            if (!((push >> r) & 1)) continue;
                            const int first = (r == 0 ? slice_lo : halo.b_lo)*EU_SLICE;
                            const int last = (r == 0 ? halo.a_hi : slice_hi)*EU_SLICE;
                            const int d = (c >= first && c < last) ? halo.dst[r][c - first] : -1;
                            if (d >= 0) {
                                halo.peer_S[r][d] = sat;
                                if (CAP && halo.peer_pc[r]) halo.peer_pc[r][d] = pcn;
                            }
                        }
                    }
                }
                m.dS4 = dS5;
            }
            m.S0 = S1; m.lw0 = lw1; m.lo0 = lo1; m.rock0 = rock1; m.pc0 = pc1;
            c += D;
            if (MULTIROCK) hc += D;
            ++slot;
            if (slot == NS) { slot = 0; par ^= 1u; }
            r_cur = r_next;
            r_next = (r_next == 2) ? 0 : r_next + 1;
        }
        if (SHARE) {
            __syncthreads();                                    // the last plane's fluxes are visible
            if (in_tile && p_update) finish_prev();
        }
        if (push) {
            if (push & 1) ++done0;
            if (push & 2) ++done1;
        }
        if ((done0 | done1) && (u + 1 >= u_end || (__ldg(&b.units[u + 1].w) & 3) == 0)) {
            // this was the block's last unit with pushes (they are the first of its list): all of them are ordered
            // before thread 0 by the block barrier, its ONE system-scope fence is cumulative, then the finished-unit
            // counters of the two ranges (a fence per unit held the whole block back for its NVLink round trip)
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                const unsigned dn[2] = { done0, done1 };
                for (int r = 0; r < 2; ++r) {
                    if (dn[r] == 0u) continue;
                    const unsigned before = atomicAdd(halo.counter[r], dn[r]);
                    if (before + dn[r] == halo.total[r]) {
                        *halo.counter[r] = 0u;
                        __threadfence_system();
                        *(volatile unsigned*)halo.peer_flag[r] = halo.epoch;
                        __threadfence_system();
                    }
                }
            }
            done0 = done1 = 0u;
        }
    }
}

} // namespace
