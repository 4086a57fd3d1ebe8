#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/compact_check.py > gpurun_out/r02c_compact_check.txt 2>&1; tail -12 gpurun_out/r02c_compact_check.txt
timeout 200 python tools/probe_compact.py > gpurun_out/r02c_probe_compact.txt 2>&1; tail -45 gpurun_out/r02c_probe_compact.txt
D2P_PARITY_LOG=gpurun_out timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -15 gpurun_out/r02c_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'])"
timeout 300 python tools/timeline.py > gpurun_out/r02c_timeline_c2.txt 2>&1; cat gpurun_out/r02c_timeline_c2.txt | tail -25
