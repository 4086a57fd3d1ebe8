"""Run under ncu: one eager ViZDoom (C4) train step and one induction (C5) greedy decode,
so their kernels show up in a launch list.  argv[1] = c4 | c5."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config, vizdoom_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch, make_vizdoom_batch

which = sys.argv[1] if len(sys.argv) > 1 else 'c4'
import os
from demo2program_b200 import _lib
# ncu cannot launch cooperative + clustered kernels: plain launch for the persistent recurrences
_lib.load().d2p_lstm_set_persistent(int(os.environ.get('D2P_PERSIST', '2')))
if which == 'c4':
    cfg = vizdoom_config('full', batch_size=32, k=10)
    eng = Engine(cfg, use_graph=False, concurrent=False)
    eng.stage_batch(make_vizdoom_batch(cfg, seed=123))
    for _ in range(2):
        eng.train_step_device(True)
else:
    from demo2program_b200.induction import InductionEngine
    cfg = karel_config('induction_baseline', batch_size=512, k=10)
    eng = InductionEngine(cfg, is_train=False)
    eng.stage_batch(make_batch(cfg, seed=123))
    for _ in range(2):
        eng.encode(exact=False)
        eng.greedy(exact=False)
torch.cuda.synchronize()
print('done', eng.lib.d2p_launch_count())
