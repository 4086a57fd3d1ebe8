"""Developer tool (GPU, under compute-sanitizer): one forward + backward call of the fused Karel conv
kernels at a small size (4 x 3 demonstrations, 20 frames each; u8 and fp32 frames)."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

lib = _lib.load()
for dtype in (np.uint8, np.float32):
    cfg = karel_config('full', batch_size=4, k=3)
    eng = Engine(cfg, use_graph=False, concurrent=False, frames_dtype=dtype)
    batch = make_batch(cfg, seed=1)
    if dtype == np.float32:
        batch = dict(batch)
        batch['s_h'] = np.asarray(batch['s_h'], np.float32)
    eng.stage_batch(batch)
    st = torch.cuda.current_stream().cuda_stream
    eng.dfeat.normal_()
    check(lib.d2p_conv_encoder_fwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.feat), ptr(eng.conv_saved), 1,
                                   ptr(eng.ws), eng.ws_bytes, st), 'fwd')
    check(lib.d2p_conv_encoder_bwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.dfeat), ptr(eng.conv_saved), 1,
                                   ptr(eng.ws), eng.ws_bytes, st), 'bwd')
    torch.cuda.synchronize()
    print(dtype.__name__, 'feat', float(eng.feat.abs().sum()), 'grads', float(eng.grads.abs().sum()))
