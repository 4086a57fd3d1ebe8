#!/bin/bash
# weak scaling sample on one 8-GPU box: N = 8 and N = 4 (B = 32 per GPU), bucketed all-reduce inside the step graph
mkdir -p gpurun_out
TAG=${TAG:-r03c}
for N in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  echo "N=$N rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d.get('cross_rank_param_checksum_match'))"
done
