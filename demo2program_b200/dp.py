"""Data-parallel plumbing: one process per GPU, the batch sharded across ranks, the flat gradient
buffer summed over the ranks with NCCL and nothing else (BASELINE.json north_star).
The reference is single-device (trainer.py:134-138, device_count={'GPU': 1}); semantics here are
"N reference towers at B=32 with averaged gradients": BatchNorm statistics and the loss normalisers
stay rank-local.

The sum is ONE logical all-reduce of the flat buffer, issued as two contiguous pieces in the order
the backward pass finishes them (`gradient_buckets`), on a communication stream that runs under
the rest of the backward pass: the decoders' and summary pools' gradients (71 % of the parameters
of `full`) are final while the second-path / encoder recurrences and the frame encoder still
back-propagate; only the second piece (the two encoders, 29 %) is exposed.  (Measured at N=2,
profiles/r02f_timeline_n2_*: a separate third piece for the second-path encoder does not pay - its
weight-gradient products finish on the low-priority gradient streams only ~50 us before the
demonstration encoder's, and every piece costs ~40 us of launch + synchronisation latency.)"""
import torch
import torch.distributed as dist


def init_nccl(device=None):
    """init_process_group('nccl') for one process per GPU.  D2P_NCCL_MAX_CTAS > 0 limits the
    communicator to that many CTAs (the gradient all-reduce runs UNDER the backward pass and shares the
    SMs with its cooperative recurrence kernels).  Default 0 = NCCL's own choice: measured at N=2
    (profiles/r02r_*), 8 CTAs make the hidden 32 MB bucket no cheaper for the backward pass and the
    exposed last bucket 43 us slower (3.009 vs 2.986 ms/step)."""
    import os
    kw = {}
    if device is not None:
        kw['device_id'] = torch.device(device)
    max_ctas = int(os.environ.get('D2P_NCCL_MAX_CTAS', '0'))
    if max_ctas > 0:
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = min(max_ctas, 4)
            kw['pg_options'] = opts
        except (AttributeError, RuntimeError):      # a torch build without the NCCL config fields
            pass
    dist.init_process_group('nccl', **kw)


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


# order in which the backward pass completes the gradients of the top-level variable scopes
# (reference models/model_full.py: decoders and summary pools first, then the second-path encoder and
# the demonstration encoder with its State_Encoder convolutions)
BUCKET_OF_SCOPE = {'Demo_Encoder': 1, 'SecondPathEncoder': 1}
N_BUCKETS = 2


def gradient_buckets(manifest):
    """[(lo, hi), ...] per bucket: maximal contiguous index ranges of the flat buffer whose
    variables share a bucket (0 = final first).  Every element belongs to exactly one range."""
    buckets = [[] for _ in range(N_BUCKETS)]
    for e in manifest:
        b = BUCKET_OF_SCOPE.get(e.name.split('/')[0], 0)
        r = buckets[b]
        if r and r[-1][1] == e.offset:
            r[-1][1] = e.offset + e.size
        else:
            r.append([e.offset, e.offset + e.size])
    return [[tuple(x) for x in r] for r in buckets]


class BucketedAllReduce:
    """In-place SUM of the flat gradient buffer over the ranks, bucket by bucket.

    reduce(b, producers) is called by the backward pass as soon as everything that writes bucket b
    has been ENQUEUED on the `producers` streams: the communication stream waits for those streams'
    current tails and all-reduces the bucket's ranges there, overlapping whatever the producers
    enqueue next.  join(stream) makes `stream` wait for all reductions (before clip + Adam).  All of
    it is stream-ordered, so it can be captured into the step's CUDA graph (NCCL collectives are
    capturable).  On CPU tensors (gloo tests) the calls are synchronous."""

    def __init__(self, flat_grad, manifest, world_size, device=None):
        self.flat = flat_grad
        self.world = int(world_size)
        self.buckets = gradient_buckets(manifest)
        self.cuda = flat_grad.is_cuda
        self.comm = torch.cuda.Stream(device, priority=-1) if self.cuda and self.world > 1 else None
        self.reduced = set()

    def reduce(self, b, producers=()):
        if self.world <= 1:
            return
        assert b not in self.reduced, 'bucket %d reduced twice in one step' % b
        self.reduced.add(b)
        if self.comm is not None:
            for s in producers:
                self.comm.wait_stream(s)
            with torch.cuda.stream(self.comm):
                for lo, hi in self.buckets[b]:
                    dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM)
        else:
            for lo, hi in self.buckets[b]:
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM)

    def join(self, stream=None):
        """All buckets must have been reduced; returns the 1/world scale of clip+Adam."""
        if self.world > 1:
            missing = [b for b in range(N_BUCKETS) if b not in self.reduced and self.buckets[b]]
            assert not missing, 'gradient buckets %s were never reduced' % missing
            if self.comm is not None:
                (stream or torch.cuda.current_stream()).wait_stream(self.comm)
        self.reduced = set()
        return 1.0 / self.world


def shutdown(timeout_s=20.0):
    """destroy_process_group with a deadline: a communicator that still has captured work
    outstanding can block its teardown forever; results have been reported by then, so the process
    leaves with os._exit(0) rather than hang its launcher."""
    import os
    import sys
    import threading
    if not dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    if t.is_alive():
        sys.stdout.flush()
        sys.stderr.write('dp.shutdown: destroy_process_group did not return in %.0f s; exiting\n' % timeout_s)
        sys.stderr.flush()
        os._exit(0)
