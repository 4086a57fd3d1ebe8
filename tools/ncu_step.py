"""Run under ncu: one eager (non-graph) C2 train step after one warm-up step so
every kernel of the step shows up in the launch list."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

import os
from demo2program_b200 import _lib
# ncu cannot launch a kernel that is both cooperative and clustered (LaunchFailed): profile the
# persistent kernels with a plain launch (mode 2; launches are serialised under ncu anyway)
_lib.load().d2p_lstm_set_persistent(int(os.environ.get('D2P_PERSIST', '2')))
cfg = karel_config('full', batch_size=32, k=10)
eng = Engine(cfg, use_graph=False)
batch = make_batch(cfg, seed=123)
eng.stage_batch(batch)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(steps):
    eng.train_step_device(True)
torch.cuda.synchronize()
print('loss', float(eng.loss[0]), 'launches', eng.lib.d2p_launch_count())
