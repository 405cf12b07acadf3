#!/bin/bash
TAG=${1:-call3}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > $O/pytest_gpu.log
( timeout 300 python tools/perf_case.py t3 --n 100 --check-strict > $O/perf_t3_100.json 2> $O/perf_t3_100.err )
cat $O/pytest_gpu.log; tail -3 $O/perf_t3_100.err; cat $O/perf_t3_100.json
