"""Developer tool (GPU): SM-clock timeline of CTA 0 of the fused Karel conv backward kernel."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

lib = _lib.load()
cfg = karel_config('full', batch_size=32, k=10)
eng = Engine(cfg, use_graph=False, concurrent=False)
eng.stage_batch(make_batch(cfg, seed=1))
st = torch.cuda.current_stream().cuda_stream
probe = torch.zeros(128, dtype=torch.int64, device=eng.dev)
fwd = lambda: check(lib.d2p_conv_encoder_fwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.feat), ptr(eng.conv_saved), 1, ptr(eng.ws), eng.ws_bytes, st), 'f')
bwd = lambda: check(lib.d2p_conv_encoder_bwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.dfeat), ptr(eng.conv_saved), 1, ptr(eng.ws), eng.ws_bytes, st), 'b')
for _ in range(3):
    fwd(); bwd()
torch.cuda.synchronize()
lib.d2p_debug_set_probe(ptr(probe))
bwd()
torch.cuda.synchronize()
lib.d2p_debug_set_probe(None)
p = probe.cpu().tolist()[96:]
names = ['start', 'loads issued', 'layer-3 BN backward done', 'dW3 done', 'dy2 done', 'layer-2 BN backward done', 'dW2 done',
         'dy1 done', 'layer-1 BN backward done', 'dW1 done', 'grid barrier passed', 'partials reduced']
for i, n in enumerate(names):
    print('%-28s +%d cycles' % (n, p[i] - p[0]))
