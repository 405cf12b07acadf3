#!/bin/bash
O=gpurun_out/${1:-multiab}; N=${2:-2}
mkdir -p $O
for cfg in "EU_X=0" "EU_MARCH_LEN=32"; do
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 3 --warmup 3 --substeps 40 --no-cpu --no-e2e 2>/dev/null | grep '^{' | python -c '
import sys,json
d=json.loads(sys.stdin.read()); print("'"$cfg"'", round(d["value"]/1e9,2), "kernel_ms", round(d["roofline"]["kernel_ms"],4))' | tee -a $O/ab.txt
done
