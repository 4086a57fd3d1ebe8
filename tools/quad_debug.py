"""Developer tool (GPU): quad-form input gradient vs the per-class form, forcing several tiles per CTA."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from demo2program_b200 import _lib
from demo2program_b200.config import vizdoom_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch
lib = _lib.load()
def run(cfg, batch, mode):
    lib.d2p_conv_set_tc(mode); lib.d2p_conv_set_fused(0)
    eng = Engine(cfg, use_graph=False, concurrent=False)
    eng.stage_batch(batch); eng.forward(); eng.backward(); torch.cuda.synchronize()
    g = eng.grads.cpu().numpy().copy()
    lib.d2p_conv_set_tc(7); lib.d2p_conv_set_fused(1)
    return g
which = sys.argv[1]
cfg = vizdoom_config('full', batch_size=2, k=2, max_demo_len=3, test_k=2, max_program_len=8) if which == 'b2' else \
      vizdoom_config('full', batch_size=8, k=3, max_demo_len=8, test_k=2, max_program_len=8)
batch = make_batch(cfg, seed=3)
ref = run(cfg, batch, 2 | 16 | 32)
out = run(cfg, batch, 2 | 16)
print(which, 'grid', os.environ.get('D2P_CONV_TC_GRID'), 'mask', os.environ.get('D2P_CONV_QUAD_MASK'),
      'rel err %.3e' % (np.abs(out - ref).max() / np.abs(ref).max()))
