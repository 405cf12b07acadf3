#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh <tag> N [extra bench args / "ENV=.. ENV2=.." as $4]'
# 2-rank parity check (tests/multigpu_check.py through tests/test_multigpu.py), then torchrun bench.py on N ranks.
TAG=${1:-multi}; N=${2:-2}; ARGS=${3:-}; ENVS=${4:-}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -30 ) > $O/pytest_multigpu.log
( env $ENVS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-configs $ARGS > $O/bench_n$N.json 2> $O/bench_n$N.err )
tail -12 $O/pytest_multigpu.log; tail -3 $O/bench_n$N.err; cat $O/bench_n$N.json
