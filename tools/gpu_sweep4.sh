#!/bin/bash
TAG=${1:-sweep4}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5 ) > $O/pytest.log
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
B="python bench.py --steps 3 --warmup 3 --substeps 20 --no-cpu --no-e2e"
run nz32_auto   "EU_X=0"            $B --nz 32
run nz32_len8   "EU_MARCH_LEN=8"    $B --nz 32
run nz128_auto  "EU_X=0"            $B --nz 128
run nz128_len32 "EU_MARCH_LEN=32"   $B --nz 128
run nz64_auto   "EU_X=0"            $B --nz 64
run nz256_auto  "EU_X=0"            $B --nz 256
cat $O/pytest.log
