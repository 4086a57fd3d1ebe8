#!/bin/bash
# wide (4 x 80-row) tiling of the encoder / second-path recurrences: parity, then bench + timeline A/B
mkdir -p gpurun_out
TAG=${TAG:-r02p}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -k "persistent_recurrence or c2 or full_size or determin or matches_oracle" 2>&1 | tail -6 | cut -c1-250
for W in 1 0; do
  D2P_WIDE_LSTM=$W timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_n1_wide$W.json 2> gpurun_out/${TAG}_bench_n1_wide$W.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n1_wide$W.json')); print('wide=$W', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['second_kernel']['kernel_ms'])"
  D2P_WIDE_LSTM=$W timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_c2_wide$W.txt 2>&1; tail -22 gpurun_out/${TAG}_timeline_c2_wide$W.txt
done
