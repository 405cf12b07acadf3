#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh <tag> N'
TAG=${1:-multi}; N=${2:-2}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 600 python -m pytest tests/test_multigpu.py tests/test_dropin_cpp.py -m gpu -x -q 2>&1 | tail -30 ) > $O/pytest_multigpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $O/bench_n$N.json 2> $O/bench_n$N.err )
cat $O/pytest_multigpu.log; tail -3 $O/bench_n$N.err; cat $O/bench_n$N.json
