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


def time_dominant_kernel(eng, iters=30):
    """CUDA-event time of the dominant kernel in isolation.

    By share of the step (profiles/: ncu launch list) the top kernel is
    `gemm_tc_kernel<64,4>`: the per-time-step recurrent products of the five LSTMs
    (backward: dh_{t-1} = dZ_t * Wh^T, [R, 4H] x [4H, H], split-K over 6 slabs; 130 launches
    per step) plus the dW products.  It is timed here exactly as the recurrence launches
    it: packed (bf16 hi/lo) operands prepared beforehand, split-K partial sums as output.
    Algorithmic FLOPs = 2*M*N*K (fp32-equivalent; the tensor pipe executes 3x that)."""
    import torch
    from demo2program_b200._lib import ptr
    lib = eng.lib
    R, H = eng.R, eng.H
    M, N, K = R, H, 4 * H
    A = torch.randn(M, K, device=eng.dev)
    Bm = torch.randn(N, K, device=eng.dev)
    apk = torch.empty(lib.d2p_packed_bytes(M, K), dtype=torch.uint8, device=eng.dev)
    bpk = torch.empty(lib.d2p_packed_bytes(N, K), dtype=torch.uint8, device=eng.dev)
    ks = 6
    part = torch.empty(ks * M * N, device=eng.dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.dev)
    st = torch.cuda.current_stream(eng.dev)
    lib.d2p_pack_bf16(ptr(A), M, K, K, 1, ptr(apk), st.cuda_stream)
    lib.d2p_pack_bf16(ptr(Bm), N, K, K, 1, ptr(bpk), st.cuda_stream)

    def launch():
        rc = lib.d2p_gemm_tc_packed(ptr(apk), ptr(bpk), M, N, K, 1.0, 0.0, None, N, None, ks,
                                    ptr(part), st.cuda_stream)
        assert rc == 0, lib.d2p_last_error()

    for _ in range(3):
        launch()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        # the step itself finds both operands in L2 (weights packed once per step, dZ_t written
        # by the preceding kernel): touch them again after the flush
        apk.add_(0); bpk.add_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        launch()
        e1.record(st)
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    flops = 2.0 * M * N * K
    return {'kernel': 'gemm_tc_kernel<64,4> (tcgen05 bf16x3, split-K 6: recurrent dh = dZ_t * Wh^T)',
            'shape': [M, N, K], 'ms': ms, 'tflops': flops / (ms * 1e-3) / 1e12}


def run_ours(args):
    import torch
    from demo2program_b200.config import karel_config
    from demo2program_b200.engine import Engine
    from demo2program_b200.synthetic import make_batch, program_tokens_in_batch
    rank, local, world = dist_env()
    if world != args.gpus:
        world = int(os.environ.get('WORLD_SIZE', args.gpus)) if 'WORLD_SIZE' in os.environ else 1
    dev = 'cuda:%d' % local
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(dev))
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
    # ---- e2e: public API with host buffers ----
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
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    ms_per_step = dev_ms / args.steps
    value = toks_all / (ms_per_step * 1e-3)
    pk, pk_kind = peaks()
    dom = time_dominant_kernel(eng)
    roofline = {
        'bound': 'tensor', 'achieved': dom['tflops'], 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
        'frac': dom['tflops'] / pk['bf16_tflops'], 'traffic': None,
        'kernel': dom['kernel'], 'shape_MNK': dom['shape'], 'kernel_ms': dom['ms'],
        'peak_kind': pk_kind + ' bf16 burst (cuBLAS 8192^3); achieved counts ALGORITHMIC flops '
                     '2*M*N*K - the bf16x3 split executes 3x that on the tensor pipe',
        'tensor_pipe_tflops_executed': 3 * dom['tflops'],
    }
    cpu = None
    if not args.no_cpu_baseline:
        sec, ctoks, threads = cpu_reference_step_time(cfg, 2, 1)
        cpu = {'value': ctoks / sec, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': '2 full train steps of %s (1 warm-up), oracle restatement of the TF1 graph, '
                         'torch CPU fp32' % WORKLOAD, 'ms_per_step': sec * 1e3}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': 32 * world, 'per_gpu_batch': 32,
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
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
