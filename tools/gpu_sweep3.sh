#!/bin/bash
TAG=${1:-sweep3}
O=gpurun_out/$TAG
mkdir -p $O
EXP=$PWD/opm-porsol_b200/lib_exp/libeuler_b200.so
( EU_B200_LIB=$EXP timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -8 ) > $O/pytest_exp.log
line() { python -c 'import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l)
    if "roofline" in d: print("%.2f Gcs/s kernel_ms %.4f frac %.3f mhz %s" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    else: print("%.2f Gcs/s kernel_ms %.4f frac %.3f strictdiff %s" % (d["cell_substeps_per_s"]/1e9, d["kernel_ms"], d["frac_of_measured_peak"], d.get("fast_vs_strict_max_abs")))
' ; }
run() { local name=$1; local cfg=$2; shift 2
  out=$(env $cfg timeout 300 "$@" 2>$O/err_$name.log | tail -1)
  echo "$name [$cfg] :: $(echo "$out" | line 2>/dev/null || echo FAILED)" | tee -a $O/sweep.txt
}
B="python bench.py --nz 64 --steps 3 --warmup 3 --substeps 20 --no-cpu --no-e2e"
run vg_stored   "EU_X=0"           $B
run vg_recomp   "EU_B200_LIB=$EXP" $B
run vgc_stored  "EU_X=0"           $B --capillary
run vgc_recomp  "EU_B200_LIB=$EXP" $B --capillary
run c2_stored   "EU_X=0"           python tools/perf_case.py c2 --n 200
run c2_recomp   "EU_B200_LIB=$EXP" python tools/perf_case.py c2 --n 200 --check-strict
run c3_stored   "EU_X=0"           python tools/perf_case.py c3
run c3_recomp   "EU_B200_LIB=$EXP" python tools/perf_case.py c3 --check-strict
run vg_recomp_m4 "EU_B200_LIB=$EXP EU_FAST_VARIANT=2" $B
cat $O/pytest_exp.log
