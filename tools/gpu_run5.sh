#!/bin/bash
# session baseline: test-suite, full bench (secondary C1/C4/C5 + cpu baseline), timeline, launch list
mkdir -p gpurun_out
TAG=${TAG:-r02g}
D2P_PARITY_LOG=gpurun_out timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench_n1.json
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_c2.txt 2>&1; tail -40 gpurun_out/${TAG}_timeline_c2.txt
timeout 600 python tools/component_bench.py > gpurun_out/${TAG}_components.json 2> gpurun_out/${TAG}_components.err; tail -c 2500 gpurun_out/${TAG}_components.json
