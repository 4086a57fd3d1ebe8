"""bench.py contract, CPU side: the reference arm (the CPU restatement of the TF1 graph timed on the
host cores) prints exactly one JSON line with the keys the driver reads; under a multi-rank launch
only rank 0 prints; our arm fails loudly without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.pop('RANK', None); e.pop('WORLD_SIZE', None); e.pop('LOCAL_RANK', None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'program_tokens_per_sec_train_step'
    assert d['unit'] == 'program-tokens/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['steps'] == 1 and d['n_gpus'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['config']['workload'] == 'karel_full_k10_b32_T20_L50_H512'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
             env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a CUDA device is present')
    r = _run(['--steps', '1', '--warmup', '0'])
    assert r.returncode != 0 and r.stdout.strip() == ''


def test_roofline_traffic_comes_from_the_committed_capture(tmp_path):
    """roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel,
    taken from the committed `ncu --set full` summary through profiles/ncu_traffic.json (tools/ncu_traffic.py)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import bench
    import ncu_traffic
    traffic, source = bench.ncu_traffic('lstm_persist_fwd')
    assert source and os.path.exists(os.path.join(ROOT, source))
    parsed = ncu_traffic.parse(os.path.join(ROOT, source))
    rec = [v for k, v in parsed.items() if 'lstm_persist_fwd' in k][0]
    assert traffic == rec['dram_bytes_per_launch'] == rec['dram_read'] + rec['dram_write']
    assert 1e6 < traffic < 1e9          # tens of MB per launch: weights once + gates / cells / Y of 20 steps
    assert bench.ncu_traffic('no_such_kernel') == (None, None)
