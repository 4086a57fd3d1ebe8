python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 10 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -c 1500 gpurun_out/bench9.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches7.csv python tools/ncu_step.py 2 > gpurun_out/ncu8.log 2>&1; tail -2 gpurun_out/ncu8.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 200 -c 4 -o gpurun_out/prof8 -f python tools/ncu_step.py 2 > gpurun_out/ncu8b.log 2>&1; tail -2 gpurun_out/ncu8b.log
ncu -i gpurun_out/prof8.ncu-rep --page raw --csv > gpurun_out/prof8_raw.csv 2>/dev/null; wc -c gpurun_out/prof8_raw.csv
