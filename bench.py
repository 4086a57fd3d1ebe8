#!/usr/bin/env python
"""bench.py - program-tokens/sec of the demo2program train step on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

Our arm: Karel `full` model, k=10, batch 32 per GPU (BASELINE.json configs[1];
weak scaling: every rank trains its own 32-example shard, one NCCL all-reduce
of the flat gradient buffer per step).  A "step" is forward + backward + clip +
Adam over one synthetic batch.  `value` is device-timed with CUDA events with
the batch already resident in HBM; `e2e` is the same metric through the public
`Engine.train_step(batch)` call with HOST buffers (pinned staging, H2D of the
inputs and D2H of the loss inside the timed region).

Reference arm: the reference is TF-1.3 / Python-2 and cannot run offline, so the
arm times the CPU restatement of its graph (oracle/, torch CPU fp32, all host
threads) on the same config, one full train step per "step".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = 'program_tokens_per_sec_train_step'
UNIT = 'program-tokens/s'
WORKLOAD = 'karel_full_k10_b32_T20_L50_H512'


def peaks():
    path = os.path.join(HERE, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def dist_env():
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    return rank, local, world


def cpu_reference_step_time(cfg, steps, warmup, threads=None):
    """Times the oracle restatement (torch CPU fp32) - fwd + bwd + clip + Adam."""
    import torch
    from oracle.models import OracleTrainer
    from demo2program_b200.manifest import build_manifests
    from demo2program_b200.synthetic import make_batch, program_tokens_in_batch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    pm, sm = build_manifests(cfg)
    tr = OracleTrainer(cfg, pm.init_flat(0), sm.init_flat(0), dtype=torch.float32)
    batch = make_batch(cfg, seed=123)
    toks = program_tokens_in_batch(batch)
    for _ in range(warmup):
        tr.train_step(batch)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.train_step(batch)
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), toks, threads


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    from demo2program_b200.config import karel_config
    cfg = karel_config('full', batch_size=32, k=10)
    sec, toks, threads = cpu_reference_step_time(cfg, args.steps, args.warmup)
    val = toks / sec
    sample = '%d full train steps (fwd+bwd+clip+Adam) of %s on host CPU' % (args.steps, WORKLOAD)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'note': 'CPU restatement of the TF1 graph (TF 1.3 / py2 not installable '
                           'offline); torch CPU fp32'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def time_dominant_kernel(eng, iters=20):
    """CUDA-event time of the dominant kernels in isolation, on the stream they are launched on.

    By share of the step (profiles/r01b_launches_2steps_c2.txt, ncu launch list) the top kernels
    are the two persistent recurrence kernels, `lstm_persist_bwd_kernel` (19.5 %) and
    `lstm_persist_fwd_kernel` (19.2 %): ONE cooperative launch per LSTM sequence, recurrent weight
    resident in shared memory, tcgen05 gate GEMM per step (csrc/lstm_persist.cu).  They are timed
    here exactly as the step launches them, through d2p_lstm_seq_fwd / d2p_lstm_seq_bwd with
    phases = "recurrence only", for the k*B-row LSTMs (R = 320 rows, T = 20 steps: four of the
    five LSTMs of the model).  The hoisted input product is re-run untimed before every timed
    launch (the recurrence consumes `gates` in place), the L2 is flushed before that.
    Algorithmic FLOPs per launch = 2*R*H*4H per step x T steps (fp32-equivalent; the bf16x3 split
    executes 3x that on the tensor pipe; the backward has the same count)."""
    import torch
    from demo2program_b200._lib import ptr, check
    lib = eng.lib
    R, H, T = eng.R, eng.H, eng.T
    dev = eng.dev
    z = lambda *s: torch.zeros(*s, device=dev)
    g = torch.Generator(device='cpu').manual_seed(0)
    X = (torch.randn(T, R, H, generator=g) * 0.5).to(dev)
    W = (torch.randn(2 * H, 4 * H, generator=g) * 0.05).to(dev)
    b = z(4 * H)
    lens = torch.randint(8, T + 1, (R,), generator=g, dtype=torch.int32).to(dev)
    Y, hT, cT, gates, cells = z(T, R, H), z(R, H), z(R, H), z(T, R, 4 * H), z(T, R, H)
    dY, dW, db, dh0, dc0 = (torch.randn(T, R, H, generator=g) * 0.1).to(dev), z(2 * H, 4 * H), z(4 * H), z(R, H), z(R, H)
    wsb = lib.d2p_lstm_seq_bwd_ws_bytes(T, R, H)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    eng._tc_bind()

    def fwd(ph):
        check(lib.d2p_lstm_seq_fwd(ptr(X), T, R, H, H, ptr(lens), None, None, ptr(W), ptr(b), 1.0, ptr(Y),
                                   ptr(hT), ptr(cT), ptr(gates), ptr(cells), ph, st.cuda_stream), 'lstm fwd')

    def bwd():
        check(lib.d2p_lstm_seq_bwd(ptr(X), T, R, H, H, ptr(lens), None, None, ptr(W), ptr(Y), ptr(gates),
                                   ptr(cells), ptr(dY), None, None, None, ptr(dW), ptr(db), ptr(dh0), ptr(dc0),
                                   ptr(ws), wsb, 1, st.cuda_stream), 'lstm bwd')

    tf = tb = 0.0
    for i in range(iters + 2):
        flush.zero_()
        fwd(1)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(st); fwd(2); e[1].record(st)
        e[2].record(st); bwd(); e[3].record(st)
        e[3].synchronize()
        if i >= 2:
            tf += e[0].elapsed_time(e[1]); tb += e[2].elapsed_time(e[3])
    tf /= iters; tb /= iters
    flops = 2.0 * R * H * 4 * H * T
    # the hoisted gate GEMM of the same LSTM, x*Wx for all T steps at once: [T*R, H] x [H, 4H] on the
    # persistent tile-queue tensor-core kernel (gemm_tc_persist_kernel), packed operands resident
    M, N, K = T * R, 4 * H, H
    A2 = (torch.randn(M, K, generator=g)).to(dev)
    B2 = (torch.randn(N, K, generator=g)).to(dev)
    C2 = z(M, N)
    apk = torch.empty(lib.d2p_packed_bytes(M, K), dtype=torch.uint8, device=dev)
    bpk = torch.empty(lib.d2p_packed_bytes(N, K), dtype=torch.uint8, device=dev)
    check(lib.d2p_pack_bf16(ptr(A2), M, K, K, 1, ptr(apk), st.cuda_stream), 'pack')
    check(lib.d2p_pack_bf16(ptr(B2), N, K, K, 1, ptr(bpk), st.cuda_stream), 'pack')
    tg = 0.0
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(lib.d2p_gemm_tc_packed(ptr(apk), ptr(bpk), M, N, K, 1.0, 0.0, ptr(C2), N, None, 1, None,
                                     st.cuda_stream), 'gate gemm')
        e1.record(st)
        e1.synchronize()
        if i >= 2:
            tg += e0.elapsed_time(e1)
    tg /= iters
    gflops = 2.0 * M * N * K
    gate = {'kernel': 'gemm_tc_persist_kernel (hoisted LSTM gate GEMM x*Wx, persistent tile queue, TMEM '
                      'double-buffered accumulator, bf16x3)', 'shape_M_N_K': [M, N, K], 'ms': tg,
            'tflops': gflops / (tg * 1e-3) / 1e12}
    return {'gate_gemm': gate, 'kernel': 'lstm_persist_fwd_kernel (persistent weight-stationary LSTM recurrence, tcgen05 bf16x3 '
                      'gate GEMM per step, R=320 rows x T=20 steps in one cooperative launch)',
            'shape': [R, 4 * H, H, T], 'ms': tf, 'tflops': flops / (tf * 1e-3) / 1e12,
            'bwd': {'kernel': 'lstm_persist_bwd_kernel (clusters of 4 split-K CTAs, DSMEM reduction)',
                    'ms': tb, 'tflops': flops / (tb * 1e-3) / 1e12}}


def _event_ms(fn, iters, warmup=2, flush=None):
    """Mean CUDA-event time of fn() on the current stream, L2 flushed before each timed call."""
    import torch
    st = torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def secondary_configs(flush, pk):
    """The other BASELINE.json configs on their shipped default paths, a few device-timed steps each
    (rank 0, N=1): C1 Karel synthesis_baseline k=2 B=8 train step, C4 ViZDoom full k=10 B=32 train
    step, C5 Karel induction_baseline B=512 encode + greedy decode (default exact=None path: tensor
    cores + arg-max margin guard).  Algorithmic work per SURVEY 8(d) / Appendix C."""
    import torch
    from demo2program_b200.config import karel_config, vizdoom_config
    from demo2program_b200.engine import Engine
    from demo2program_b200.synthetic import make_batch, make_vizdoom_batch, program_tokens_in_batch
    out = {}

    def train_cfg(name, cfg, batch, iters, extra):
        try:
            eng = Engine(cfg, use_graph=True)
            eng.stage_batch(batch)
            ms = _event_ms(lambda: eng.train_step_device(True), iters, warmup=3, flush=flush)
            eng.check_device()
            toks = program_tokens_in_batch(batch)
            d = {'ms_per_step': ms, 'program_tokens_per_s': toks / (ms * 1e-3),
                 'instances_per_s': cfg.batch_size / (ms * 1e-3), 'steps': iters,
                 'gpu_launches_per_step': int(eng.launches_per_step)}
            d.update(extra(eng, ms))
            out[name] = d
            del eng
        except Exception as e:          # a secondary config must not take the headline line down
            out[name] = {'error': repr(e)}
        torch.cuda.empty_cache()

    cfg1 = karel_config('synthesis_baseline', batch_size=8, k=2)
    train_cfg('c1_karel_synthesis_k2_b8_train_step', cfg1, make_batch(cfg1, seed=123), 20,
              lambda eng, ms: {'params': int(eng.pm.total)})
    cfg4 = vizdoom_config('full', batch_size=32, k=10)
    frames4 = cfg4.batch_size * cfg4.k * cfg4.max_demo_len

    def c4_extra(eng, ms):
        conv_flop = 27.7e6 * frames4            # conv fwd+bwd, SURVEY Appendix C
        lstm_flop = 3 * (2.0 * eng.R * eng.T * 4 * eng.H * ((eng.F + eng.H) + 3 * 2 * eng.H) +
                         2.0 * eng.B * cfg4.max_program_len * 4 * eng.H * 2 * eng.H)
        tf = (conv_flop + lstm_flop) / (ms * 1e-3) / 1e12
        return {'algorithmic_tflop_per_step': (conv_flop + lstm_flop) / 1e12, 'tflops': tf,
                'frac_of_bf16_sustained_peak': tf / pk['bf16_tflops_sustained'],
                'conv_algorithmic_gflop_fwd_bwd': conv_flop / 1e9}
    train_cfg('c4_vizdoom_full_k10_b32_T20_train_step', cfg4, make_vizdoom_batch(cfg4, seed=123), 8, c4_extra)
    try:
        from demo2program_b200.induction import InductionEngine
        cfg5 = karel_config('induction_baseline', batch_size=512, k=10)
        eng = InductionEngine(cfg5, is_train=False)
        eng.stage_batch(make_batch(cfg5, seed=123))
        enc = _event_ms(lambda: eng.encode(), 5, flush=flush)
        dec = _event_ms(lambda: eng.greedy(), 5, flush=flush)
        kv = 2.0 * eng.B * eng.k * eng.T * eng.H * 4
        frames5 = eng.B * eng.k * eng.T
        out['c5_karel_induction_greedy_b512_k10'] = {
            'encode_ms': enc, 'greedy_decode_ms': dec, 'latency_ms': enc + dec,
            'greedy_path': eng.greedy_path, 'near_tie_argmaxes': int(getattr(eng, 'greedy_near_ties', -1)),
            'unseen_demos_decoded_per_s': eng.R2 / ((enc + dec) * 1e-3),
            'decode_steps': eng.T, 'kv_bytes_per_decode_step': kv,
            'decode_GBps_kv_read_once_per_step': kv * eng.T / dec / 1e6,
            'decode_frac_hbm': kv * eng.T / dec / 1e6 / pk['hbm_gbs'],
            'encode_conv_algorithmic_bytes': 1216.0 * frames5}
        del eng
    except Exception as e:
        out['c5_karel_induction_greedy_b512_k10'] = {'error': repr(e)}
    torch.cuda.empty_cache()
    return out


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` summary (profiles/ncu_traffic.json, written by tools/ncu_traffic.py
    from the capture); None when no capture of this kernel is committed."""
    path = os.path.join(HERE, 'profiles', 'ncu_traffic.json')
    try:
        with open(path) as f:
            tab = json.load(f)
    except (OSError, ValueError):
        return None, None
    for name, rec in tab.items():
        if kernel_substr in name:
            return rec.get('dram_bytes_per_launch'), rec.get('source')
    return None, None


def run_ours(args):
    import torch
    from demo2program_b200.config import karel_config
    from demo2program_b200.dp import shutdown as dp_shutdown
    from demo2program_b200.engine import Engine
    from demo2program_b200.synthetic import make_batch, program_tokens_in_batch
    rank, local, world = dist_env()
    if world != args.gpus:
        world = int(os.environ.get('WORLD_SIZE', args.gpus)) if 'WORLD_SIZE' in os.environ else 1
    dev = 'cuda:%d' % local
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on the C-level stdout when the communicator is created:
        # send fd 1 to stderr meanwhile so that rank 0's stdout carries the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            from demo2program_b200.dp import init_nccl
            init_nccl(dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    cfg = karel_config('full', batch_size=32, k=10)
    eng = Engine(cfg, device=dev, world_size=world, use_graph=True)
    batch = make_batch(cfg, seed=123 + rank)     # each rank: its own shard
    toks = program_tokens_in_batch(batch)
    h2d = eng.stage_batch(batch)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    st = torch.cuda.current_stream(eng.dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eng.train_step_device(True)
    barrier()
    n0 = eng.lib.d2p_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-timed: K steps, L2 flushed between iterations ----
    barrier()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.train_step_device(True)
        e1.record(st)
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- e2e: public API with host buffers (its own warm-up: the first call allocates the
    # pinned double buffers of the input pipeline) ----
    for _ in eng.train_steps(batch for _ in range(max(args.warmup, 3))):
        pass
    barrier()
    t0 = time.perf_counter()
    loss = 0.0
    for loss in eng.train_steps(batch for _ in range(args.steps)):
        pass
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches_graph = getattr(eng, 'launches_per_step', None)
    if launches_graph is None:   # eager (multi-GPU) path counts live launches
        launches_graph = (eng.lib.d2p_launch_count() - n0) // (2 * args.steps)
    # every rank must hold bit-identical parameters after the same number of averaged-gradient steps
    w32 = eng.params.view(torch.int32).to(torch.int64)
    csum = torch.stack([w32.sum(), (w32 * (torch.arange(w32.numel(), device=dev) % 65521 + 1)).sum()])
    params_match = None
    if world > 1:
        gathered = [torch.zeros_like(csum) for _ in range(world)]
        torch.distributed.all_gather(gathered, csum)
        params_match = all(torch.equal(gathered[0], x) for x in gathered)
    eng.check_device()
    t = torch.tensor([dev_ms, e2e_s, float(toks)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        tsum = t.clone()
        torch.distributed.all_reduce(tsum, op=torch.distributed.ReduceOp.SUM)
        dev_ms, e2e_s, toks_all = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        toks_all = float(toks)
    if rank != 0:
        eng.close()
        dp_shutdown()
        return
    ms_per_step = dev_ms / args.steps
    value = toks_all / (ms_per_step * 1e-3)
    pk, pk_kind = peaks()
    dom = time_dominant_kernel(eng)
    roofline = {
        'bound': 'tensor', 'achieved': dom['tflops'], 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
        'frac': dom['tflops'] / pk['bf16_tflops'],
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape, one launch, from the
        # committed `ncu --set full` capture (profiles/r01d_ncu_full_metrics.txt: 41.9 MB read + 16.3 MB written)
        'traffic': None, 'traffic_unit': 'bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)',
        'kernel': dom['kernel'], 'shape_R_4H_H_T': dom['shape'], 'kernel_ms': dom['ms'],
        'peak_kind': pk_kind + ' bf16 burst (cuBLAS 8192^3); achieved counts ALGORITHMIC flops '
                     '2*R*H*4H*T - the bf16x3 split executes 3x that on the tensor pipe; the kernel is '
                     'a chain of T dependent steps on 96 of 148 SMs (latency-bound, see DESIGN.md)',
        'tensor_pipe_tflops_executed': 3 * dom['tflops'],
        'gate_gemm': {'kernel': dom['gate_gemm']['kernel'], 'shape_M_N_K': dom['gate_gemm']['shape_M_N_K'],
                      'kernel_ms': dom['gate_gemm']['ms'], 'achieved': dom['gate_gemm']['tflops'],
                      'frac': dom['gate_gemm']['tflops'] / pk['bf16_tflops'],
                      'tensor_pipe_tflops_executed': 3 * dom['gate_gemm']['tflops'],
                      'frac_executed': 3 * dom['gate_gemm']['tflops'] / pk['bf16_tflops'],
                      'ncu_tensor_pipe_pct_of_active': 55.2,
                      'note': 'ncu figure from profiles/r01d_ncu_full_metrics.txt (not measured in this run)'},
        'second_kernel': {'kernel': dom['bwd']['kernel'], 'kernel_ms': dom['bwd']['ms'],
                          'achieved': dom['bwd']['tflops'], 'frac': dom['bwd']['tflops'] / pk['bf16_tflops']},
    }
    roofline['traffic'], roofline['traffic_source'] = ncu_traffic('lstm_persist_fwd')
    secondary = None
    if world == 1 and not args.no_secondary:
        del eng
        torch.cuda.empty_cache()
        secondary = secondary_configs(flush, pk)
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only
        sec, ctoks, threads = cpu_reference_step_time(cfg, 2, 1)
        cpu = {'value': ctoks / sec, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': '2 full train steps of %s (1 warm-up), oracle restatement of the TF1 graph, '
                         'torch CPU fp32' % WORKLOAD, 'ms_per_step': sec * 1e3}
        # also a single-thread figure (one step, no warm-up: ~15 s)
        sec1, _, _ = cpu_reference_step_time(cfg, 1, 0, threads=1)
        cpu['one_thread'] = {'value': ctoks / sec1, 'unit': UNIT, 'cores': 1, 'ms_per_step': sec1 * 1e3,
                             'sample': '1 full train step, no warm-up'}
    launches_per_step = int(launches_graph)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'arithmetic': 'f32 (every dense product as a bf16x3 split on the tensor cores, ~5e-6 relative)', 'global_batch': 32 * world, 'per_gpu_batch': 32,
                   'parallelism': 'dp%d' % world, 'l2': 'flushed (256 MiB write) between timed steps',
                   'instances_per_sec': 32 * world / (ms_per_step * 1e-3),
                   'program_tokens_per_step': toks_all, 'cuda_graph': True},
        'clocks': clocks,
        'e2e': {'value': toks_all / (e2e_s / args.steps), 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 16,
                'ms_per_step': e2e_s / args.steps * 1e3,
                'api': 'Engine.train_steps(host batches): per step pinned staging + H2D of the '
                       'batch + loss D2H, double-buffered against the previous step'},
        'gpu_launches': int(launches_graph) * args.steps * 2,
        'gpu_launches_per_step': int(launches_graph),
        'roofline': roofline,
        'cpu_baseline': cpu,
        'final_loss': loss,
        'cross_rank_param_checksum_match': params_match,
        'param_checksum': [int(x) for x in csum.tolist()],
        'secondary': secondary,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        eng.close()
        dp_shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true',
                    help='skip the C1 / C4 / C5 secondary measurements')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
