"""Developer tool (GPU): in-graph timeline of one C2 train step.  One-thread stamp kernels
(%globaltimer) are captured between the ops of the step; the replay tells when each branch
of the concurrent graph finishes."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

cfg = karel_config('full', batch_size=32, k=10)
from demo2program_b200 import _lib
_lib.load().d2p_lstm_set_persistent(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
eng = Engine(cfg, use_graph=True)
eng.timeline = torch.zeros(64, dtype=torch.int64, device=eng.dev)
eng.timeline_names = []
eng.stage_batch(make_batch(cfg, seed=123))
for _ in range(5):
    eng.train_step_device(True)
torch.cuda.synchronize()
t = eng.timeline.cpu().tolist()
t0 = t[eng.timeline_names.index('fwd start')]
for name, v in sorted(zip(eng.timeline_names, t), key=lambda kv: kv[1]):
    print('%9.1f us  %s' % ((v - t0) / 1e3, name))
