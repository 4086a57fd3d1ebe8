"""Developer tool (GPU): SM-clock timeline of one fused LSTM forward step CTA."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check

lib = _lib.load()
dev = 'cuda:0'
st = torch.cuda.current_stream().cuda_stream
T, R, In, H = 20, 320, 512, 512
scratch = torch.zeros(128 << 20, dtype=torch.uint8, device=dev)
cache = torch.zeros(128 << 20, dtype=torch.uint8, device=dev)
lib.d2p_tc_configure(ptr(scratch), scratch.numel(), ptr(cache), cache.numel(), 1)
X = torch.randn(T, R, In, device=dev) * 0.1
W = torch.randn(In + H, 4 * H, device=dev) * 0.05
b = torch.zeros(4 * H, device=dev)
ln = torch.full((R,), T, dtype=torch.int32, device=dev)
Y = torch.zeros(T, R, H, device=dev); hT = torch.zeros(R, H, device=dev); cT = torch.zeros(R, H, device=dev)
gates = torch.zeros(T, R, 4 * H, device=dev); cells = torch.zeros(T, R, H, device=dev)
probe = torch.zeros(64, dtype=torch.int64, device=dev)


def run():
    check(lib.d2p_lstm_seq_fwd(ptr(X), T, R, In, H, ptr(ln), None, None, ptr(W), ptr(b), 1.0, ptr(Y), ptr(hT),
                               ptr(cT), ptr(gates), ptr(cells), 3, st), 'fwd')


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.d2p_debug_set_probe(ptr(probe))
run()
torch.cuda.synchronize()
p = probe.cpu().tolist()
t0 = p[0]
print('last step CTA(0,0): prologue done +%d, accum ready +%d, epilogue done +%d, teardown +%d cycles' % (
    p[1] - t0, p[2] - t0, p[3] - t0, p[4] - t0))
print('producer issue stamps:', [p[8 + i] - t0 for i in range(16)])
print('mma full-wait done   :', [p[32 + i] - t0 for i in range(16)])
lib.d2p_debug_set_probe(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print('lstm fwd T=20: %.1f us per call' % (e0.elapsed_time(e1) * 100))
