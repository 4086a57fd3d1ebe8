"""Data-parallel plumbing: one process per GPU, the batch sharded across ranks,
ONE all-reduce (sum) of the flat gradient buffer per step and nothing else
(BASELINE.json north_star).  The reference is single-device
(trainer.py:134-138, device_count={'GPU': 1}); semantics here are "N reference
towers at B=32 with averaged gradients": BatchNorm statistics and the loss
normalisers stay rank-local."""
import torch.distributed as dist


def shard_seed(base_seed, rank):
    """Every rank draws its own shard of synthetic examples."""
    return int(base_seed) + 1009 * int(rank)


def allreduce_flat_gradients(flat_grad, world_size=None):
    """In-place SUM all-reduce of the flat gradient buffer (NCCL on GPU, gloo in
    the CPU tests).  Returns the 1/world scale the fused clip+Adam kernel
    applies (d2p_clip_adam_step's grad_scale)."""
    world = world_size or (dist.get_world_size() if dist.is_initialized() else 1)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world
