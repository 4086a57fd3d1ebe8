#!/bin/bash
# quick GPU validation: test-suite + bench (TAG from env)
mkdir -p gpurun_out
TAG=${TAG:-r02g}
D2P_PARITY_LOG=gpurun_out timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'])"
