#!/bin/bash
# round-2 GPU run: full GPU test suite, 1-GPU bench line, ncu launch list of smoke()
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/r02a_smoke_ncu.log 2>&1
echo "ncu smoke rc=$?"
tail -3 gpurun_out/r02a_smoke_ncu.log
