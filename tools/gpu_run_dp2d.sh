#!/bin/bash
# N=2 at the end of the round: bench (cross-rank parameter checksum), timeline
mkdir -p gpurun_out
TAG=${TAG:-r03n}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d.get('cross_rank_param_checksum_match'))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/timeline.py > gpurun_out/${TAG}_timeline_n2.txt 2>&1; grep " us " gpurun_out/${TAG}_timeline_n2.txt | tail -24
