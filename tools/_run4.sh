python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/component_bench.py c4 c5 > gpurun_out/components2.json 2> gpurun_out/components2.err; python - <<'PY'
import json
for l in open('gpurun_out/components2.json'):
    d=json.loads(l)
    if d.get('component')=='c4': print('c4', d['conv_fwd']['us'], d['conv_fwd_bwd']['us'], d['graph_step_ms'])
    else: print(d)
PY
tail -3 gpurun_out/components2.err
