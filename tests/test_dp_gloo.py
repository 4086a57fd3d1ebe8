"""world_size-2 gloo test (CPU) of the data-parallel host logic: shards are
independent, ONE all-reduce (sum) of the flat gradient buffer, 1/world scaling
folded into the clip+Adam step.  The per-rank gradients come from
the oracle here (test infrastructure); the GPU path uses the same helper with
NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from demo2program_b200.config import karel_config
from demo2program_b200.manifest import build_manifests
from demo2program_b200.synthetic import make_batch
from demo2program_b200.dp import BucketedAllReduce, allreduce_flat_gradients, gradient_buckets, shard_seed


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle.models import OracleModel
    from oracle import tf_ops as T
    cfg = karel_config('synthesis_baseline', batch_size=2, k=2, num_lstm_cell_units=8,
                       max_program_len=6, max_demo_len=4)
    pm, sm = build_manifests(cfg)
    p0 = pm.init_flat(0).astype(np.float64)
    m = OracleModel(cfg, p0, sm.init_flat(0))
    batch = make_batch(cfg, seed=shard_seed(100, rank), min_demo_len=2, min_prog_len=5)
    loss, grad, _ = m.loss_and_grad(batch)
    local = grad.clone()
    scale = allreduce_flat_gradients(grad, world)       # in-place SUM, returns 1/world
    avg = grad * scale
    # the same sum, bucket by bucket in completion order (what Engine issues under the backward pass)
    piece = local.clone()
    dp = BucketedAllReduce(piece, pm, world)
    for b in range(len(dp.buckets)):
        dp.reduce(b)
    assert dp.join() == scale and torch.equal(piece, grad)
    q.put((rank, float(loss), local.numpy(), avg.numpy()))
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=300) for _ in ps], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    (_, l0, g0, a0), (_, l1, g1, a1) = res
    assert l0 != l1                                   # different shards
    np.testing.assert_allclose(a0, a1, rtol=0, atol=0)  # every rank holds the same average
    np.testing.assert_allclose(a0, (g0 + g1) / 2, rtol=1e-12, atol=1e-15)


def test_shard_seeds_are_distinct():
    assert len({shard_seed(123, r) for r in range(8)}) == 8


def test_gradient_buckets_partition_the_flat_buffer():
    """Every element of the flat gradient buffer is in exactly one bucket range; the demonstration
    and second-path encoders form the last bucket (completion order of the backward pass)."""
    for model in ('full', 'summarizer', 'synthesis_baseline', 'induction_baseline'):
        pm, _ = build_manifests(karel_config(model, batch_size=4, k=3))
        buckets = gradient_buckets(pm)
        cover = np.zeros(pm.total, np.int32)
        for r in buckets:
            for lo, hi in r:
                cover[lo:hi] += 1
        assert (cover == 1).all()
        e = pm['Demo_Encoder/rnn/basic_lstm_cell/kernel']
        assert any(lo <= e.offset < hi for lo, hi in buckets[-1])
        if model in ('full', 'summarizer'):
            e = pm['SecondPathEncoder/rnn/basic_lstm_cell/kernel']
            assert any(lo <= e.offset < hi for lo, hi in buckets[1])
