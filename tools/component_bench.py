"""Developer tool (GPU): per-component device time against the roofline SURVEY.md
section 8(d) assigns to each component.  Writes one JSON object per line to stdout.

  c2   Karel full k=10 B=32 train step: per-op time (eager, CUDA events), conv encoder
       GB/s on its algorithmic bytes, Adam GB/s
  c4   ViZDoom full k=10 B=32 train step (graph) + per-op time of the conv encoder
  c5   Karel induction baseline, greedy decode, B=512, eval-mode BN: encode + decode time,
       decode GB/s on keys+values read once per step
"""
import collections
import json
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from demo2program_b200.config import karel_config, vizdoom_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch, make_vizdoom_batch, program_tokens_in_batch

HBM_PEAK = 6548.0      # GB/s, measured copy bandwidth (B200_PROFILING.md fallback)
try:
    HBM_PEAK = float(json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', HBM_PEAK))
except Exception:
    pass


def timed_ops(eng, n=5):
    """per-op device time (us/step) of an eager step, caches warm"""
    records = []
    orig = eng._call

    def timed(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        records.append((name, e0, e1))

    eng._call = timed
    for _ in range(n):
        eng.train_step_device(True)
    torch.cuda.synchronize()
    eng._call = orig
    agg = collections.OrderedDict()
    for name, a, b in records:
        agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) * 1e3 / n
    return agg


def graph_step_ms(eng, steps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.dev)
    for _ in range(5):
        eng.train_step_device(True)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.train_step_device(True)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def c2():
    cfg = karel_config('full', batch_size=32, k=10)
    batch = make_batch(cfg, seed=123)
    eng = Engine(cfg, use_graph=False, concurrent=False)
    eng.stage_batch(batch)
    for _ in range(3):
        eng.train_step_device(True)
    ops = timed_ops(eng)
    frames = cfg.batch_size * cfg.k * cfg.max_demo_len
    fwd_b, train_b = 1216 * frames, 2432 * frames        # u8 frame in + fp32 feature out
    cf, cb = ops['d2p_conv_encoder_fwd'], ops['d2p_conv_encoder_bwd']
    n = eng.pm.total
    adam_b = n * (7 * 4 + 4)                             # p,g,m,v read + p,m,v written + g re-read for the norm
    out = {'component': 'c2', 'workload': 'karel_full_k10_b32', 'eager_ops_us': {k: round(v, 1) for k, v in ops.items()},
           'conv_fwd': {'us': cf, 'algorithmic_bytes': fwd_b, 'GBps': fwd_b / cf / 1e3, 'frac_hbm': fwd_b / cf / 1e3 / HBM_PEAK},
           'conv_fwd_bwd': {'us': cf + cb, 'algorithmic_bytes': train_b, 'GBps': train_b / (cf + cb) / 1e3,
                            'frac_hbm': train_b / (cf + cb) / 1e3 / HBM_PEAK},
           'adam': {'us': ops['d2p_clip_adam_step'], 'algorithmic_bytes': adam_b,
                    'GBps': adam_b / ops['d2p_clip_adam_step'] / 1e3,
                    'frac_hbm': adam_b / ops['d2p_clip_adam_step'] / 1e3 / HBM_PEAK}}
    del eng
    eng = Engine(cfg, use_graph=True)
    eng.stage_batch(batch)
    out['graph_step_ms'] = graph_step_ms(eng)
    print(json.dumps(out), flush=True)


def c4():
    cfg = vizdoom_config('full', batch_size=32, k=10)
    batch = make_vizdoom_batch(cfg, seed=123)
    toks = program_tokens_in_batch(batch)
    eng = Engine(cfg, use_graph=False, concurrent=False)
    eng.stage_batch(batch)
    for _ in range(2):
        eng.train_step_device(True)
    ops = timed_ops(eng, 3)
    frames = cfg.batch_size * cfg.k * cfg.max_demo_len
    cf, cb = ops['d2p_conv_encoder_fwd'], ops['d2p_conv_encoder_bwd']
    out = {'component': 'c4', 'workload': 'vizdoom_full_k10_b32_T20', 'eager_ops_us': {k: round(v, 1) for k, v in ops.items()},
           'conv_fwd': {'us': cf, 'algorithmic_flops': 9.24e6 * frames, 'TFLOPs': 9.24e6 * frames / cf / 1e6},
           'conv_fwd_bwd': {'us': cf + cb, 'algorithmic_flops': 27.7e6 * frames,
                            'TFLOPs': 27.7e6 * frames / (cf + cb) / 1e6}}
    del eng
    torch.cuda.empty_cache()
    eng = Engine(cfg, use_graph=True)
    eng.stage_batch(batch)
    ms = graph_step_ms(eng, 10)
    out['graph_step_ms'] = ms
    out['program_tokens_per_s'] = toks / (ms * 1e-3)
    print(json.dumps(out), flush=True)


def c5():
    from demo2program_b200.induction import InductionEngine
    cfg = karel_config('induction_baseline', batch_size=512, k=10)
    batch = make_batch(cfg, seed=123)
    eng = InductionEngine(cfg, is_train=False)
    eng.stage_batch(batch)

    def t(fn, n=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    enc = t(lambda: eng.encode(exact=False))
    dec = t(lambda: eng.greedy(exact=False))
    dec_exact = t(lambda: eng.greedy(exact=True), 2)
    kv = 2 * eng.B * eng.k * eng.T * eng.H * 4
    steps = eng.T
    out = {'component': 'c5', 'workload': 'karel_induction_greedy_b512_k10_testk%d' % eng.tk,
           'encode_ms': enc, 'greedy_decode_ms': dec, 'greedy_decode_exact_fp32_ms': dec_exact,
           'decode_steps': steps, 'kv_bytes_per_step': kv,
           'decode_GBps_on_kv_read_once_per_step': kv * steps / dec / 1e6,
           'frac_hbm': kv * steps / dec / 1e6 / HBM_PEAK,
           'demos_decoded_per_s': eng.R2 / ((enc + dec) * 1e-3)}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['c2', 'c4', 'c5']
    for w in which:
        try:
            {'c2': c2, 'c4': c4, 'c5': c5}[w]()
        except Exception as e:   # keep going: each component is independent
            print(json.dumps({'component': w, 'error': repr(e)}), flush=True)
        torch.cuda.empty_cache()
