#!/bin/bash
# One gpurun call, stages chosen on the command line (replaces the round-1 one-off scripts):
#   gpurun --timeout 1500 -- 'bash tools/gpu.sh <tag> tests smoke bench ref perf launches ncu ncu_cap sanitize'
# Everything lands in gpurun_out/<tag>/.  Multi-GPU: tools/gpu_multi.sh.
TAG=${1:-r02}; shift
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA node\(s\)" > $O/cpu.txt 2>&1
for stage in "$@"; do
case $stage in
tests)    ( timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 ) > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log ;;
smoke)    ( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > $O/smoke.log; cat $O/smoke.log ;;
bench)    ( timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ); cat $O/bench_n1.json; tail -3 $O/bench_n1.err ;;
ref)      ( timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err ); cat $O/bench_ref.json ;;
perf)     for c in "c2 --n 100" "c2 --n 200" "c2b --n 100" "c3" "t3 --n 100"; do
            n=$(echo $c | tr -d ' -'); ( timeout 400 python tools/perf_case.py $c --check-strict > $O/perf_$n.json 2> $O/perf_$n.err ); cat $O/perf_$n.json; done ;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_VG_8M.csv \
            python bench.py --steps 1 --warmup 1 --substeps 10 --nz 32 --no-cpu --no-e2e --no-configs > $O/launches_bench.log 2>&1 ;;
ncu)      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_box_step|k_fast_step" -s 5 -c 1 -f -o $O/fast_VG_8M \
            python bench.py --steps 1 --warmup 1 --substeps 10 --nz 32 --no-cpu --no-e2e --no-configs > $O/ncu_full.log 2>&1 ;;
ncu_cap)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_box_step|k_fast_step" -s 5 -c 1 -f -o $O/fast_VGC_8M \
            python bench.py --steps 1 --warmup 1 --substeps 10 --nz 32 --capillary --no-cpu --no-e2e --no-configs > $O/ncu_full_cap.log 2>&1 ;;
sanitize) for tool in memcheck racecheck synccheck; do
            ( timeout 400 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
              -k "c3_faulted or c4_march or tensor_2rocks or c2_aniso or retry" 2>&1 | tail -12 ) > $O/$tool.log; tail -3 $O/$tool.log; done ;;
*)        # anything else: a bench sweep line "name:ENV=val,ENV2=val:bench args"
          IFS=':' read -r name envs args <<< "$stage"
          out=$(env $(echo $envs | tr ',' ' ') timeout 400 python bench.py --steps 3 --warmup 3 --substeps 20 --no-cpu --no-e2e --no-configs $args 2>$O/err_$name.log | tail -1)
          echo "$name [$envs] [$args] :: $(echo "$out" | python -c 'import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print("%.2f Gcs/s kernel_ms %.4f frac %.3f mhz %s" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
' 2>/dev/null || echo FAILED)" | tee -a $O/sweep.txt ;;
esac
done
ls $O
