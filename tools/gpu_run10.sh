#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02s}
timeout 120 python tools/probe_persist.py 2>&1 | grep -A12 "T=50 R=32 forward"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -k "persistent_recurrence or c2 or full_size or determin or matches_oracle or greedy or trajectory" 2>&1 | tail -5 | cut -c1-250
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['second_kernel']['kernel_ms'])"
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_c2.txt 2>&1; tail -21 gpurun_out/${TAG}_timeline_c2.txt
