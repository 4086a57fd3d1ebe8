#!/bin/bash
# N=2: bench with the NCCL communicator limited to 8 CTAs vs NCCL's default, timeline, NCCL correctness check
mkdir -p gpurun_out
TAG=${TAG:-r02r}
for C in 8 0; do
  D2P_NCCL_MAX_CTAS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n2_ctas$C.json 2> gpurun_out/${TAG}_bench_n2_ctas$C.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n2_ctas$C.json').read().strip().splitlines()[-1]); print('max_ctas=$C', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d.get('cross_rank_param_checksum_match'))"
  D2P_NCCL_MAX_CTAS=$C timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/timeline.py > gpurun_out/${TAG}_timeline_n2_ctas$C.txt 2>&1; grep " us " gpurun_out/${TAG}_timeline_n2_ctas$C.txt | tail -12
done
timeout 600 python -m pytest tests/test_dp_nccl.py -q -x 2>&1 | tail -3
