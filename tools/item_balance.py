#!/usr/bin/env python
"""Static load balance of the persistent substep kernel, computed on the host (no GPU).

The kernel hands work item v to warp v mod #warps (148 SMs x 3 blocks x 8 warps on B200, 2 blocks with the capillary
term).  An item is a march of L z-steps over one slice of 32 cells; the kernel ends when the warp with the most steps
ends.  This tool reproduces the item lists eu_api.cu:build_items builds for a Cartesian slab (one chain per column
slice, single-slice items for the domain-boundary planes, whose class differs) and prints, per piece-length rule, the
busiest warp's load against the mean -- the number behind the choice of the work-item length (DESIGN.md sections 5-7).

    python tools/item_balance.py --nx 512 --ny 512 --planes 256            # one GPU, whole grid
    python tools/item_balance.py --nx 512 --ny 512 --planes 32 --ranks 8   # a middle rank of an 8-GPU run
"""
import argparse


def pieces_equal(n, L):
    """ceil(n/L) pieces of nearly equal length, as build_items cuts a chain."""
    out = []
    left = n
    while left > 0:
        p = (left + L - 1)//L
        ln = (left + p - 1)//p
        out.append(ln)
        left -= ln
    return out


def items_for(columns, chain, singles, L):
    """Item lengths in list order: build_items scans slices in ascending order, so all pieces that start in the same
    plane are adjacent.  singles = planes (0, 1 or 2) whose slices form one-step items (domain boundary planes)."""
    lens = []
    if singles >= 1:
        lens += [1]*columns                      # bottom plane
    for ln in pieces_equal(chain, L):
        lens += [ln]*columns
    if singles >= 2:
        lens += [1]*columns                      # top plane
    return lens


def makespan(lens, n_warps, head):
    load = [0.0]*n_warps
    for v, ln in enumerate(lens):
        load[v % n_warps] += ln + head
    return max(load), sum(load)/n_warps


def fixed_rule(n_slices, n_warps):
    L = 32
    while L > 2 and n_slices//L < 6*n_warps:
        L //= 2
    return L


def auto_rule(columns, chain, singles, n_warps):
    best, bestL = 1e300, 32
    for L in range(4, 65):
        pieces = columns*((chain + L - 1)//L) + columns*singles
        steps = columns*(chain + singles)
        cost = ((pieces + n_warps - 1)//n_warps)*(steps/pieces + 0.35)
        if cost <= best:
            best, bestL = cost, L
    return bestL


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--ny", type=int, default=512)
    ap.add_argument("--planes", type=int, default=256, help="own planes of this rank")
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--position", default="middle", choices=["first", "middle", "last"], help="rank position in the slab stack")
    ap.add_argument("--blocks-per-sm", type=int, default=3)
    ap.add_argument("--head", type=float, default=0.35, help="cost of starting a march, in steps")
    a = ap.parse_args()
    n_warps = 148*a.blocks_per_sm*8
    columns = a.nx*a.ny//32
    if a.ranks == 1:
        boundary_planes, singles = 0, 2
    else:
        boundary_planes = 2 if a.position == "middle" else 1        # planes next to a neighbour: generic path, processed first
        singles = 0 if a.position == "middle" else 1
    interior = a.planes - boundary_planes
    chain = interior - singles
    print(f"{columns} column slices, {interior} interior planes (chain {chain} + {singles} single-step planes), "
          f"{boundary_planes} halo planes, {n_warps} warps")
    def sim_rule(lo, hi, head):
        best, bestL = 1e300, lo
        for L in range(lo, hi + 1):
            mx, _ = makespan(items_for(columns, chain, singles, L), n_warps, head)
            if mx < best:
                best, bestL = mx, L
        return bestL
    rules = {"fixed": fixed_rule(columns*interior, n_warps), "auto(<=64)": auto_rule(columns, chain, singles, n_warps),
             "sim(6..16, head 0.75: decomposed runs, if >= 4 % better than fixed)": sim_rule(6, 16, 0.75)}
    print(f"rules: {rules}")
    print(f"{'L':>4} {'items':>8} {'items/warp':>10} {'max load':>9} {'mean':>8} {'imbalance':>9}")
    print("(eu_api.cu evaluates the auto rule with 32 warps per SM, as measured; fixed and sim as shown here)")
    for L in sorted(set(list(range(4, 65, 2)) + list(rules.values()))):
        lens = items_for(columns, chain, singles, L)
        mx, mean = makespan(lens, n_warps, a.head)
        tag = " ".join(k for k, v in rules.items() if v == L)
        print(f"{L:4d} {len(lens):8d} {len(lens)/n_warps:10.2f} {mx:9.1f} {mean:8.1f} {mx/mean - 1:9.1%}  {tag}")


if __name__ == "__main__":
    main()
