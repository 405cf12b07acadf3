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
//     {q, G}(k) of the x-, y- and z-faces [and T(k)]                              4-D boxes (tx+1) x (ty+1) x 1 x 1
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
    const int4* units;            // {x0 | y0 << 16, z0, z1 (exclusive), flags: bit 0/1 = halo range A/B, i.e. push results to a peer}
    int stages;                   // bundles in flight
    int off_bar, off_lam, off_rk, off_stage;      // byte offsets in dynamic shared memory (from the 128-aligned base)
    int lam_bytes, rk_bytes, stage_bytes;
    int off_S, off_pc, off_qg, off_T;             // inside a stage
    int qg_bytes, T_bytes;                        // one axis' box
    const unsigned short* cmask;
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

template <bool ROCKS, bool MULTIROCK, bool CAP>
__global__ void __launch_bounds__(256, CAP ? 2 : 3)
k_box_step(const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapPc,
           const __grid_constant__ CUtensorMap mapQG, const __grid_constant__ CUtensorMap mapT,
           EuGridDev g, EuTablesDev t, EuFastDev f, EuStepArgs a, EuHaloDev halo, EuBoxDev b, int slice_lo, int tab_bytes)
{
    TabLayout L;
    L.nn = t.n_nodes_total; L.nb = t.n_buckets; L.shift = 0;
    if (ROCKS) tables_to_smem(t);
    // 128-aligned base of the staging area behind the tables
    const unsigned base_u32 = (smem_u32(eu_smem) + unsigned(tab_bytes) + 127u) & ~127u;
    unsigned char* const base = eu_smem + (base_u32 - smem_u32(eu_smem));
    const int tid = threadIdx.x;
    const int NS = b.stages;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(base_u32 + b.off_bar + 8*s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (!halo.enabled) {
        const unsigned long long key = *a.fail_key;
        if (key != ~0ULL && (unsigned)(key >> 32) < (unsigned)a.substep) return;
    }
    const int tx = b.tx, ty = b.ty, txp = tx + 2, txf = tx + 1, txs = tx + 4;      // row lengths: ring / T boxes, face boxes, S boxes
    const int n_tile = tx*ty;
    const int lx = tid % tx, ly = tid/tx;
    const bool in_tile = tid < n_tile;
    const int li = (ly + 1)*txp + lx + 1;                       // own entry in a ring buffer
    const int ls = (ly + 1)*txs + lx + 2;                       // own entry in an S / pc box
    // halo duty: threads 0 .. 2(tx+ty)-1 each own one halo cell of the ring buffers
    int hx = 0, hy = 0;
    bool has_halo = true;
    if (tid < tx)                 { hx = tid;            hy = -1; }
    else if (tid < 2*tx)          { hx = tid - tx;       hy = ty; }
    else if (tid < 2*tx + ty)     { hx = -1;             hy = tid - 2*tx; }
    else if (tid < 2*tx + 2*ty)   { hx = tx;             hy = tid - 2*tx - ty; }
    else has_halo = false;
    const int hi = (hy + 1)*txp + hx + 1;
    const int hs = (hy + 1)*txs + hx + 2;
    const int D = b.nx*b.ny;
    unsigned gb = 0;                                            // bundles consumed by this block so far
    const unsigned bundle_bytes = unsigned(txs*(ty + 2)*8*(CAP ? 2 : 1) + 3*txf*(ty + 1)*16 + (CAP ? 3*txp*(ty + 1)*8 : 0));

    for (int u = blockIdx.x; u < b.n_units; u += gridDim.x) {
        const int4 unit = __ldg(b.units + u);
        const int x0 = unit.x & 0xffff, y0 = unit.x >> 16, z0 = unit.y, z1 = unit.z;
        const int nb = z1 - z0 + 1;                             // bundles = steps k = z0-1 .. z1-1
        const int range = (unit.w & 1) ? 0 : ((unit.w & 2) ? 1 : -1);
        if (range >= 0 && tid < halo.n_wait) {
            // ghosts of the previous substep must have landed before this unit's TMA loads read them
            const volatile unsigned* fl = halo.my_flags + halo.wait_rank[tid];
            const long long t0 = clock64();
            while ((int)(*fl - (halo.epoch - 1u)) < 0) {
                __nanosleep(100);
                if (clock64() - t0 > halo.timeout_cycles) { atomicExch(halo.err_flag, 1); break; }
            }
            __threadfence_system();
        }
        __syncthreads();                                        // every stage and ring buffer of the previous unit is free
        const unsigned gb0 = gb;                                // number of this unit's first bundle
        auto issue = [&](int j) {                               // bundle j of this unit: plane k = z0 - 1 + j
            const int k = z0 - 1 + j;
            const unsigned slot = (gb0 + unsigned(j)) % unsigned(NS);
            const unsigned bar = base_u32 + b.off_bar + 8*slot;
            const unsigned st = base_u32 + b.off_stage + slot*unsigned(b.stage_bytes);
            mbar_expect_tx(bar, bundle_bytes);
            tma_load_3d(st + b.off_S, &mapS, x0 - 2, y0 - 1, k + 1, bar);
            if (CAP) tma_load_3d(st + b.off_pc, &mapPc, x0 - 2, y0 - 1, k + 1, bar);
            tma_load_4d(st + b.off_qg, &mapQG, 2*(x0 - 1), y0, k, 0, bar);
            tma_load_4d(st + b.off_qg + b.qg_bytes, &mapQG, 2*x0, y0 - 1, k, 1, bar);
            tma_load_4d(st + b.off_qg + 2*b.qg_bytes, &mapQG, 2*x0, y0, k, 2, bar);
            if (CAP) {
                tma_load_4d(st + b.off_T, &mapT, x0 - 2, y0, k, 0, bar);
                tma_load_4d(st + b.off_T + b.T_bytes, &mapT, x0, y0 - 1, k, 1, bar);
                tma_load_4d(st + b.off_T + 2*b.T_bytes, &mapT, x0, y0, k, 2, bar);
            }
        };
        if (tid == 0) {
            for (int j = 0; j < NS && j < nb; ++j) issue(j);
        }
        // ---- own column: the cell of plane z0-1 (operands of the first carried face)
        const int gx = x0 + lx, gy = y0 + ly;
        const bool active = in_tile && gx < b.nx && gy < b.ny;
        const int col = gx + b.nx*gy;                           // cell of plane 0 of this column
        MarchCarry m;
        m.S0 = 0.0; m.pc0 = 0.0; m.rock0 = 0; m.dS4 = 0.0;
        if (active && z0 > 0) {
            const int c = col + (z0 - 1)*D;
            m.S0 = __ldg(a.S_in + c);
            if (MULTIROCK) m.rock0 = __ldg(f.rock8 + c);
            if (CAP) m.pc0 = __ldg(a.pc_in + c);
        }
        Mob<ROCKS, MULTIROCK>::both(L, t, m.rock0, m.S0, m.lw0, m.lo0);
        // halo cell: its column index in the grid (or -1 outside), for the rock id
        int hcol = -1;
        if (has_halo) {
            const int qx = x0 + hx, qy = y0 + hy;
            if (qx >= 0 && qx < b.nx && qy >= 0 && qy < b.ny) hcol = qx + b.nx*qy;
        }
        for (int j = 0; j < nb; ++j, ++gb) {
            const int k = z0 - 1 + j;
            const unsigned slot = gb % unsigned(NS);
            const unsigned char* st = base + b.off_stage + slot*size_t(b.stage_bytes);
            mbar_wait(base_u32 + b.off_bar + 8*slot, (gb/unsigned(NS)) & 1u);
            // ---- phase A: rock curves of plane k+1, once per cell
            const double* Sn = reinterpret_cast<const double*>(st + b.off_S);
            const double* Pn = reinterpret_cast<const double*>(st + b.off_pc);
            double* ring_next = reinterpret_cast<double*>(base + b.off_lam + ((k + 1) % 3)*size_t(b.lam_bytes));
            unsigned char* rk_next = base + b.off_rk + ((k + 1) % 3)*size_t(b.rk_bytes);
            const bool up_ok = k + 1 < b.nz;
            double S1 = 0.0, pc1 = 0.0, lw1, lo1;
            int rock1 = 0;
            if (in_tile) {
                S1 = Sn[ls];
                if (MULTIROCK && active && up_ok) rock1 = __ldg(f.rock8 + col + (k + 1)*D);
                if (CAP) pc1 = Pn[ls];
                Mob<ROCKS, MULTIROCK>::both(L, t, rock1, S1, lw1, lo1);
                if (CAP) {
                    reinterpret_cast<double4*>(ring_next)[li] = make_double4(lw1, lo1, S1, pc1);
                    if (MULTIROCK) rk_next[li] = (unsigned char)rock1;
                } else {
                    reinterpret_cast<double2*>(ring_next)[li] = make_double2(lw1, lo1);
                }
            }
            if (has_halo) {
                const double Sh = Sn[hs];
                int rh = 0;
                if (MULTIROCK && hcol >= 0 && up_ok) rh = __ldg(f.rock8 + hcol + (k + 1)*D);
                double lwh, loh;
                Mob<ROCKS, MULTIROCK>::both(L, t, rh, Sh, lwh, loh);
                if (CAP) {
                    reinterpret_cast<double4*>(ring_next)[hi] = make_double4(lwh, loh, Sh, Pn[hs]);
                    if (MULTIROCK) rk_next[hi] = (unsigned char)rh;
                } else {
                    reinterpret_cast<double2*>(ring_next)[hi] = make_double2(lwh, loh);
                }
            }
            __syncthreads();
            // the stage of the previous step is free now: refill it (bundle j - 1 + NS)
            if (tid == 0 && j >= 1 && j - 1 + NS < nb) issue(j - 1 + NS);
            // ---- phase B: faces of plane k
            const double2* QG = reinterpret_cast<const double2*>(st + b.off_qg);
            const double* TT = reinterpret_cast<const double*>(st + b.off_T);
            const int fz = ly*txf + lx;
            const size_t qstride = size_t(b.qg_bytes)/16, tstride = size_t(b.T_bytes)/8;
            const int tz = ly*txp + lx;                           // T boxes have rows of tx+2
            double2 lam5 = make_double2(lw1, lo1);
            if (in_tile) {
                const double2 qg5 = QG[2*qstride + fz];
                const double T5 = CAP ? TT[2*tstride + tz] : 0.0;
                const double dS5 = regular_slot<ROCKS, MULTIROCK, CAP, true>(L, t, a, m, lam5, S1, rock1, qg5, 1.0, false, T5, pc1);
                if (active && k >= z0) {
                    const int c = col + k*D;
                    const double inv_pv = ldg_f64(f.inv_porevol + c);
                    const unsigned mask = __ldg(b.cmask + c);
                    const double* ring = reinterpret_cast<const double*>(base + b.off_lam + (k % 3)*size_t(b.lam_bytes));
                    const unsigned char* rk = base + b.off_rk + (k % 3)*size_t(b.rk_bytes);
                    double acc = m.dS4 - dS5;
                    // lateral faces: slot order x-, x+, y-, y+
                    const int nbi[4] = { li - 1, li + 1, li - txp, li + txp };
                    const int qi[4] = { fz, fz + 1, int(qstride) + fz, int(qstride) + fz + txf };
                    const int ti[4] = { tz + 1, tz + 2, int(tstride) + tz, int(tstride) + tz + txp };   // (x box starts at x0 - 2)
                    double2 lam[4], qg[4];
                    double Sx[4], Px[4], Tx[4];
                    int rx[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (CAP) {
                            const double4 e = reinterpret_cast<const double4*>(ring)[nbi[q]];
                            lam[q] = make_double2(e.x, e.y); Sx[q] = e.z; Px[q] = e.w;
                            rx[q] = MULTIROCK ? int(rk[nbi[q]]) : 0;
                            Tx[q] = TT[ti[q]];
                        } else {
                            lam[q] = reinterpret_cast<const double2*>(ring)[nbi[q]];
                            Sx[q] = 0.0; Px[q] = 0.0; Tx[q] = 0.0; rx[q] = 0;
                        }
                        qg[q] = QG[qi[q]];
                    }
                    acc += regular_slot<ROCKS, MULTIROCK, CAP, false>(L, t, a, m, lam[0], Sx[0], rx[0], qg[0], 1.0, false, Tx[0], Px[0]);
                    acc -= regular_slot<ROCKS, MULTIROCK, CAP, true>(L, t, a, m, lam[1], Sx[1], rx[1], qg[1], 1.0, false, Tx[1], Px[1]);
                    acc += regular_slot<ROCKS, MULTIROCK, CAP, false>(L, t, a, m, lam[2], Sx[2], rx[2], qg[2], 1.0, false, Tx[2], Px[2]);
                    acc -= regular_slot<ROCKS, MULTIROCK, CAP, true>(L, t, a, m, lam[3], Sx[3], rx[3], qg[3], 1.0, false, Tx[3], Px[3]);
                    OwnMob<false> own0;
                    own0.lw[0] = m.lw0; own0.lo[0] = m.lo0;
                    if (mask) {                                 // boundary / fault / periodic faces of this cell
                        const int base_rec = f.slice_base[c >> 5];
                        acc += gather_cell_loop<ROCKS, MULTIROCK, CAP, false, false>(L, g, t, f, a, f.rec + base_rec + (c & 31),
                                                                                     32 - __clz(mask), mask, c, m.S0, m.rock0, m.pc0, own0);
                    }
                    double pcn;
                    const double sat = finish_cell<ROCKS, MULTIROCK, CAP, false>(L, t, f, a, c, m.S0, m.rock0, own0, inv_pv, acc, pcn);
                    if (range >= 0) {
                        const int first = (range == 0 ? slice_lo : halo.b_lo)*EU_SLICE;
                        const int d = halo.dst[range][c - first];
                        if (d >= 0) {
                            halo.peer_S[range][d] = sat;
                            if (CAP && halo.peer_pc[range]) halo.peer_pc[range][d] = pcn;
                        }
                    }
                }
                m.dS4 = dS5;
            }
            m.S0 = S1; m.lw0 = lw1; m.lo0 = lo1; m.rock0 = rock1; m.pc0 = pc1;
        }
        if (range >= 0) {
            // all pushes of this unit, one system-scope fence, then the finished-unit counter of the range
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                const unsigned before = atomicAdd(halo.counter[range], 1u);
                if (before + 1u == halo.total[range]) {
                    *halo.counter[range] = 0u;
                    __threadfence_system();
                    *(volatile unsigned*)halo.peer_flag[range] = halo.epoch;
                    __threadfence_system();
                }
            }
        }
    }
}

} // namespace
