#!/bin/bash
# Time the FAST substep kernel under the tuning knobs (one process per setting: the knobs are read once).
#   tools/variant_sweep.sh [nz] [extra bench.py args]
NZ=${1:-64}; shift
for cfg in "EU_NO_CLASSES=1" "EU_FAST_VARIANT=0" "EU_FAST_VARIANT=1" "EU_FAST_VARIANT=2" \
           "EU_FAST_VARIANT=0 EU_ROW_TILES=0" "EU_FAST_VARIANT=0 EU_MARCH_LEN=8" "EU_FAST_VARIANT=0 EU_MARCH_LEN=4"; do
  out=$(env $cfg python bench.py --nz $NZ --steps 3 --warmup 1 --substeps 20 --no-cpu --no-e2e "$@" 2>&1 | tail -1)
  echo "$cfg :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.2f Gcs/s kernel_ms %.4f frac %.3f mhz %s" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))' 2>/dev/null || echo "$out" | tail -c 400)"
done
