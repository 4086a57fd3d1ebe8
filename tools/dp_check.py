#!/usr/bin/env python
"""Multi-GPU correctness check of the data-parallel train step (run under torchrun, one rank per
GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29611 tools/dp_check.py

Every rank trains its own shard with `Engine(world_size=N)`; the check asserts
  * the all-reduced, 1/N-scaled gradient equals the average of the per-shard fp64 oracle gradients;
  * after K optimizer steps all ranks hold bit-identical parameters and Adam slots;
  * losses / parameters follow the oracle's averaged-gradient trajectory (N reference towers at
    B per rank with averaged gradients).
Rank 0 prints one JSON line ("ok": true/false).  The oracle is the checker only."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, 'tests'))


def bits_checksum(t):
    """Order-independent-free checksum of the exact bit pattern (int64 sum of the int32 view and
    of index-weighted words)."""
    w = t.contiguous().view(torch.int32).to(torch.int64)
    idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64) % 65521 + 1
    return torch.stack([w.sum(), (w * idx).sum()])


def main():
    rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
    dev = 'cuda:%d' % local
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=torch.device(dev))
    from demo2program_b200.config import karel_config
    from demo2program_b200.dp import shard_seed
    from demo2program_b200.engine import Engine
    from demo2program_b200.manifest import build_manifests
    from demo2program_b200.synthetic import make_batch
    from oracle.models import OracleTrainer
    from oracle import tf_ops as T
    B, k, steps = int(os.environ.get('DPCHECK_B', 4)), int(os.environ.get('DPCHECK_K', 3)), 3
    cfg = karel_config('full', batch_size=B, k=k)
    pm, sm = build_manifests(cfg)
    p0, s0 = pm.init_flat(0), sm.init_flat(0)
    batch = make_batch(cfg, seed=shard_seed(200, rank))
    res = {'world': world, 'B_per_rank': B, 'k': k, 'steps': steps}
    ok = True
    # ---- averaged gradient vs the average of the per-shard oracle gradients ----
    orc = OracleTrainer(cfg, p0, s0, dtype=torch.float64)
    loss_o, grad_o, _ = orc.model.loss_and_grad(batch)
    avg_o = grad_o.to(dev)
    dist.all_reduce(avg_o)
    avg_o = (avg_o / world).cpu().numpy()
    eng = Engine(cfg, device=dev, flat_params=p0, flat_state=s0, world_size=world, use_graph=True)
    eng.stage_batch(batch)
    eng.train_step_device(False)                       # forward + backward of this shard
    torch.cuda.synchronize()
    res['local_loss_err'] = abs(float(eng.loss[0]) - loss_o)
    g = eng.grads.clone()
    dist.all_reduce(g)
    g = (g / world).cpu().numpy()
    gmax = np.abs(avg_o).max()
    worst = 0.0
    for e in pm:
        a, b = g[e.offset:e.offset + e.size], avg_o[e.offset:e.offset + e.size]
        worst = max(worst, float(np.abs(a - b).max() / (np.abs(b).max() + 0.05 * gmax)))
    res['avg_grad_rel_err'] = worst
    ok &= res['local_loss_err'] < 1e-4 and worst < 3e-4
    # ---- K optimizer steps: identical parameters on all ranks, oracle trajectory ----
    eng = Engine(cfg, device=dev, flat_params=p0, flat_state=s0, world_size=world, use_graph=True)
    orc = OracleTrainer(cfg, p0, s0, dtype=torch.float64)
    m, v = torch.zeros_like(orc.model.flat.detach()), torch.zeros_like(orc.model.flat.detach())
    loss_err = 0.0
    for step in range(1, steps + 1):
        batch = make_batch(cfg, seed=shard_seed(300 + step, rank))
        le = eng.train_step(batch)
        lo, grad, _ = orc.model.loss_and_grad(batch)
        gd = grad.to(dev)
        dist.all_reduce(gd)
        (grad,), _ = T.clip_by_global_norm([(gd / world).cpu()], 20.0)
        with torch.no_grad():
            T.adam_step(orc.model.flat, grad, m, v, step)
        orc.model.commit_state()
        loss_err = max(loss_err, abs(le - lo))
        if step == 1:
            res['first_step_loss_err'] = loss_err
    res['max_loss_err'] = loss_err
    sums = torch.cat([bits_checksum(eng.params), bits_checksum(eng.adam_m), bits_checksum(eng.adam_v)])
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    res['params_bit_identical_across_ranks'] = bool(same)
    d = np.abs(eng.params.cpu().numpy() - orc.model.flat.detach().numpy())
    res['param_median_abs_err'], res['param_max_abs_err'] = float(np.median(d)), float(d.max())
    # step 1 is a pure forward-parity statement (1e-4, north_star); later losses are taken at parameters
    # that already differ by up to steps*lr where Adam normalised a ~0 gradient of either sign
    ok &= same and res['first_step_loss_err'] < 1e-4 and loss_err < 5e-4 and res['param_median_abs_err'] < 5e-6 and res['param_max_abs_err'] < 1e-2
    res['step_count'] = eng.step_count()
    ok &= res['step_count'] == steps
    res['ok_rank'] = bool(ok)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res['ok'] = bool(flag.item() == 1.0)
    if rank == 0:
        print(json.dumps(res), flush=True)
    else:
        sys.stderr.write('rank %d: %s\n' % (rank, json.dumps(res)))
    eng.close()
    from demo2program_b200.dp import shutdown
    shutdown()
    sys.exit(0 if res['ok'] else 1)


if __name__ == '__main__':
    main()
