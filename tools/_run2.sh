python -m pytest tests -m gpu -x -q -k "pipelined" 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
timeout 600 python tools/component_bench.py > gpurun_out/components1.json 2> gpurun_out/components1.err; cat gpurun_out/components1.json | cut -c1-1500; tail -3 gpurun_out/components1.err
