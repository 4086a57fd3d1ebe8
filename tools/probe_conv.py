"""Developer tool (GPU): SM-clock timelines of CTA 0 of the fused Karel conv forward / backward kernels,
and event timings of the fused vs the per-layer (tensor-core) paths."""
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
print('  (frames staged for dW1 +%d, dW1 products done +%d)' % (p[12] - p[0], p[13] - p[0]))

lib.d2p_debug_set_probe(ptr(probe))
fwd()
torch.cuda.synchronize()
lib.d2p_debug_set_probe(None)
p = probe.cpu().tolist()[96 + 16:]
names = ['start', 'weights staged', 'conv1 done', 'a1 saved', 'BN1 exchanged', 'conv2 done', 'BN2 exchanged (a2 saved)',
         'conv3 done', 'BN3 exchanged (a3 saved)', 'features written']
print('forward:')
for i, n in enumerate(names):
    print('%-28s +%d cycles' % (n, p[i] - p[0]))


def timed(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


q = probe.cpu().tolist()[96:]
print('  (conv1 frames staged +%d; layer-3 exchange: statistics ready +%d, slice barrier passed +%d)' % (
    q[26] - q[16], q[27] - q[16], q[28] - q[16]))
print('fused: fwd %.1f us  bwd %.1f us' % (timed(fwd), timed(bwd)))
lib.d2p_conv_set_fused(0)
for mode in (0, 7):
    lib.d2p_conv_set_tc(mode)
    print('per-layer, tc mode %d: fwd %.1f us  bwd %.1f us' % (mode, timed(fwd), timed(bwd)))
lib.d2p_conv_set_fused(1)
lib.d2p_conv_set_tc(7)
