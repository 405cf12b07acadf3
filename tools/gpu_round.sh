#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines, the BASELINE parity configurations timed, launch list, full ncu
# captures of the substep kernel, memcheck of a test subset.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r01c}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > $O/smoke.log
( timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err )
( timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err )
( timeout 300 python tools/perf_case.py c2 --n 200 --check-strict > $O/perf_c2_200.json 2> $O/perf_c2_200.err )
( timeout 300 python tools/perf_case.py c2 --n 100 > $O/perf_c2_100.json 2> $O/perf_c2_100.err )
( timeout 300 python tools/perf_case.py c2b --n 200 --check-strict > $O/perf_c2b_200.json 2> $O/perf_c2b_200.err )
( timeout 300 python tools/perf_case.py c2b --n 100 > $O/perf_c2b_100.json 2> $O/perf_c2b_100.err )
( timeout 400 python tools/perf_case.py c3 --check-strict > $O/perf_c3.json 2> $O/perf_c3.err )
( timeout 300 python tools/perf_case.py t3 --n 100 --check-strict > $O/perf_t3_100.json 2> $O/perf_t3_100.err )
( timeout 300 python bench.py --nz 64 --steps 3 --warmup 3 --substeps 20 --no-cpu --no-e2e --capillary > $O/bench_cap_nz64.json 2> $O/bench_cap_nz64.err )
# launch list of a short bench command (per-launch times are cold-cache and serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_VG_8M.csv \
    python bench.py --steps 1 --warmup 1 --substeps 10 --nz 32 --no-cpu --no-e2e > $O/launches_bench.log 2>&1
# full capture of the substep kernel (skip the first launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_step -s 5 -c 1 -f -o $O/fast_VG_8M \
    python bench.py --steps 1 --warmup 1 --substeps 10 --nz 32 --no-cpu --no-e2e > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_step -s 5 -c 1 -f -o $O/fast_VGC_8M \
    python bench.py --steps 1 --warmup 1 --substeps 10 --nz 32 --capillary --no-cpu --no-e2e > $O/ncu_full_cap.log 2>&1
( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "c3_faulted or tensor_2rocks or c2_aniso or retry" 2>&1 | tail -12 ) > $O/memcheck.log
ls -la $O
cat $O/pytest_gpu.log $O/smoke.log $O/bench_n1.json; tail -3 $O/bench_n1.err; tail -4 $O/memcheck.log
