#!/bin/bash
# One gpurun call: GPU tests, then timing sweeps of the FAST kernel's tuning knobs on the BASELINE workloads.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sweep.sh <tag>'
TAG=${1:-sweep}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/pytest_gpu.log
line() { python -c 'import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l)
    if "roofline" in d: print("%.2f Gcs/s kernel_ms %.4f frac %.3f mhz %s reg %.3f" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["config"]["regular_slot_fraction"]))
    else: print("%.2f Gcs/s kernel_ms %.4f frac %.3f reg %.3f strictdiff %s" % (d["cell_substeps_per_s"]/1e9, d["kernel_ms"], d["frac_of_measured_peak"], d["regular_slot_fraction"], d.get("fast_vs_strict_max_abs")))
' ; }
run() {   # name, env, command...
  local name=$1; local cfg=$2; shift 2
  out=$(env $cfg timeout 400 "$@" 2>$O/err_$name.log | tail -1)
  echo "$name [$cfg] :: $(echo "$out" | line 2>/dev/null || echo FAILED)" | tee -a $O/sweep.txt
}
B="python bench.py --nz 64 --steps 3 --warmup 3 --substeps 20 --no-cpu --no-e2e"
run vg_default        "EU_X=0"                          $B
run vg_noprefetch     "EU_PREFETCH=0"                   $B
run vg_minb2          "EU_FAST_VARIANT=1"               $B
run vg_minb4          "EU_FAST_VARIANT=2"               $B
run vg_minb2_nopf     "EU_FAST_VARIANT=1 EU_PREFETCH=0" $B
run vgc_default       "EU_X=0"                          $B --capillary
run vgc_noprefetch    "EU_PREFETCH=0"                   $B --capillary
run vgc_minb3         "EU_FAST_VARIANT=3"               $B --capillary
run vgc_minb3_nopf    "EU_FAST_VARIANT=3 EU_PREFETCH=0" $B --capillary
run c2_default        "EU_X=0"                          python tools/perf_case.py c2 --n 200 --check-strict
run c2_minb3          "EU_FAST_VARIANT=3"               python tools/perf_case.py c2 --n 200
run c2_nopf           "EU_PREFETCH=0"                   python tools/perf_case.py c2 --n 200
run c3_default        "EU_X=0"                          python tools/perf_case.py c3 --check-strict
run c3_minb3          "EU_FAST_VARIANT=3"               python tools/perf_case.py c3
run c3_nopf           "EU_PREFETCH=0"                   python tools/perf_case.py c3
run c3_noclass        "EU_NO_CLASSES=1"                 python tools/perf_case.py c3
( timeout 600 python bench.py --capillary --no-cpu --steps 3 > $O/bench_n1_cap.json 2> $O/bench_n1_cap.err )
( timeout 600 python bench.py --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err )
tail -3 $O/bench_n1_cap.err
cat $O/pytest_gpu.log $O/sweep.txt; tail -c 600 $O/bench_n1.json; tail -c 600 $O/bench_n1_cap.json
