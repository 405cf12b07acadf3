#!/bin/bash
TAG=${1:-sweep2}
O=gpurun_out/$TAG
mkdir -p $O
line() { python -c 'import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l)
    print("%.2f Gcs/s kernel_ms %.4f frac %.3f mhz %s" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
' ; }
run() { local name=$1; local cfg=$2; shift 2
  out=$(env $cfg timeout 300 "$@" 2>$O/err_$name.log | tail -1)
  echo "$name [$cfg] :: $(echo "$out" | line 2>/dev/null || echo FAILED)" | tee -a $O/sweep.txt
}
B="python bench.py --nz 64 --steps 3 --warmup 3 --substeps 20 --no-cpu --no-e2e"
run vg_default  "EU_X=0"                        $B
run vg_hint     "EU_L2_HINT=1"                  $B
run vg_pf2      "EU_PREFETCH=2"                 $B
run vg_pf2_hint "EU_PREFETCH=2 EU_L2_HINT=1"    $B
run vg_pf0_hint "EU_PREFETCH=0 EU_L2_HINT=1"    $B
run vg_len64    "EU_MARCH_LEN=64"               $B
run vg_len16    "EU_MARCH_LEN=16"               $B
run vg_default2 "EU_X=0"                        $B
run vgc_default "EU_X=0"                        $B --capillary
run vgc_hint    "EU_L2_HINT=1"                  $B --capillary
run vgc_pf2     "EU_PREFETCH=2"                 $B --capillary
