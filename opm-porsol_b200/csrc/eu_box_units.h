// Work units of the box kernel: the host logic (eu_host.cpp), shared with the launcher (eu_fast.cu).
#pragma once
#include <vector>

struct EuBoxUnit { int xy, z0, z1, flags; };     // x0 | y0 << 16, first plane, last plane + 1, bit 0 / 1: pushes to the rank below / above

// fills `units` (block after block; flagged first within a block in chunk mode) and `start` ([blocks + 1]);
// returns the number of units, -1 on bad arguments or when the boundary planes do not fit the slab
int eu_box_make_units(int nx, int ny, int tx, int ty, int z_lo, int z_hi, int bnd_lo, int bnd_hi, int grid_blocks, int spans,
                      int lz, std::vector<EuBoxUnit>& units, std::vector<int>& start);
