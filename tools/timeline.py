"""Developer tool (GPU): in-graph timeline of one C2 train step.  One-thread stamp kernels
(%globaltimer) are captured between the ops of the step; the replay tells when each branch
of the concurrent graph finishes."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

import os
cfg = karel_config('full', batch_size=32, k=10)
from demo2program_b200 import _lib
_lib.load().d2p_lstm_set_persistent(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
# under torchrun: the data-parallel step of rank 0 (bucketed NCCL all-reduces inside the graph)
world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    from demo2program_b200.dp import init_nccl
    init_nccl('cuda:%d' % local)
    dist.all_reduce(torch.zeros(1, device='cuda'))
eng = Engine(cfg, device='cuda:%d' % local, world_size=world, use_graph=True)
eng.timeline = torch.zeros(64, dtype=torch.int64, device=eng.dev)
eng.timeline_names = []
eng.stage_batch(make_batch(cfg, seed=123 + rank))
for _ in range(5):
    eng.train_step_device(True)
torch.cuda.synchronize()
t = eng.timeline.cpu().tolist()
t0 = t[eng.timeline_names.index('fwd start')]
if rank == 0:
    for name, v in sorted(zip(eng.timeline_names, t), key=lambda kv: kv[1]):
        print('%9.1f us  %s' % ((v - t0) / 1e3, name))
if world > 1:
    eng.close()
    from demo2program_b200.dp import shutdown
    shutdown()
