#!/bin/bash
# 2-GPU run: NCCL correctness (tools/dp_check.py via pytest) and the N=2 bench with the bucketed
# all-reduce inside the step graph vs one all-reduce between two graphs
mkdir -p gpurun_out
TAG=${TAG:-r02e}
timeout 200 python -m pytest tests/test_dp_nccl.py -m gpu -q > gpurun_out/${TAG}_pytest_dp_nccl.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_dp_nccl.log; tail -15 gpurun_out/${TAG}_pytest_dp_nccl.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/dp_check.py > gpurun_out/${TAG}_dp_check_n2.json 2> gpurun_out/${TAG}_dp_check_n2.err
echo "dp_check rc=$?"; tail -3 gpurun_out/${TAG}_dp_check_n2.json; tail -5 gpurun_out/${TAG}_dp_check_n2.err
for mode in 1 0; do
D2P_DP_OVERLAP=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2964$mode bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n2_overlap$mode.json 2> gpurun_out/${TAG}_bench_n2_overlap$mode.err
echo "bench overlap=$mode rc=$?"; tail -3 gpurun_out/${TAG}_bench_n2_overlap$mode.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n2_overlap$mode.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','cross_rank_param_checksum_match')}, d['e2e']['ms_per_step'])"
done
timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'])"
