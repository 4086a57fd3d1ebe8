"""Developer tool (GPU): tensor-core conv variants against the CUDA-core kernels on the ViZDoom B=8 geometry
(several tiles per CTA), under the developer switches D2P_CONV_TC_STAGES / D2P_CONV_TC_GRID / D2P_CONV_QUAD_MASK."""
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
    out = (eng.conv_saved.cpu().numpy().copy(), eng.grads.cpu().numpy().copy())
    lib.d2p_conv_set_tc(7); lib.d2p_conv_set_fused(1)
    return out
which = sys.argv[1]
cfg = vizdoom_config('full', batch_size=2, k=2, max_demo_len=3, test_k=2, max_program_len=8) if which == 'b2' else \
      vizdoom_config('full', batch_size=8, k=3, max_demo_len=8, test_k=2, max_program_len=8)
batch = make_batch(cfg, seed=3)
for name, mode in (('fwd', 1 | 16), ('dx per class', 2 | 16 | 32), ('dx quad', 2 | 16)):
    ref = run(cfg, batch, 16)
    out = run(cfg, batch, mode)
    print(which, name, {k: os.environ.get(k) for k in ('D2P_CONV_TC_STAGES', 'D2P_CONV_TC_GRID', 'D2P_CONV_QUAD_MASK')},
          'saved %.2e grads %.2e' % (np.abs(out[0] - ref[0]).max() / np.abs(ref[0]).max(),
                                     np.abs(out[1] - ref[1]).max() / np.abs(ref[1]).max()))
