"""Developer tool (GPU): SM-clock timeline of step 5 of the compact persistent forward kernel (CTA 0)."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check

lib = _lib.load()
dev = 'cuda:0'
st = torch.cuda.current_stream().cuda_stream
H = 512
scratch = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
cache = torch.zeros(128 << 20, dtype=torch.uint8, device=dev)
lib.d2p_tc_configure(ptr(scratch), scratch.numel(), ptr(cache), cache.numel(), 1)
NAMES = ['producer: tile top', 'producer: publishers seen', 'producer: 8 copies issued', 'mma: tile top',
         'mma: accumulator free', 'mma: first block landed', 'mma: last MMA issued', 'epi: tile top',
         'epi: accumulator ready', 'epi: cell math done', 'epi: barrier passed', 'epi: published', 'epi: stores issued']
T, R, In = 20, 320, 512
X = torch.randn(T, R, In, device=dev) * 0.1
W = torch.randn(In + H, 4 * H, device=dev) * 0.05
b = torch.zeros(4 * H, device=dev)
ln = torch.full((R,), T, dtype=torch.int32, device=dev)
z = lambda *s: torch.zeros(*s, device=dev)
Y, hT, cT, gates, cells = z(T, R, H), z(R, H), z(R, H), z(T, R, 4 * H), z(T, R, H)
probe = torch.zeros(160, dtype=torch.int64, device=dev)


def fwd(ph):
    check(lib.d2p_lstm_seq_fwd(ptr(X), T, R, In, H, ptr(ln), None, None, ptr(W), ptr(b), 1.0, ptr(Y), ptr(hT),
                               ptr(cT), ptr(gates), ptr(cells), ph, st), 'fwd')


for _ in range(2):
    fwd(3 | 8)
torch.cuda.synchronize()
lib.d2p_debug_set_probe(ptr(probe))
fwd(3 | 8)
torch.cuda.synchronize()
lib.d2p_debug_set_probe(None)
p = probe.cpu().tolist()
base = p[64]
print('== compact forward, T=%d R=%d, step 5, CTA 0: SM cycles from the producer entering tile 0' % (T, R))
for j in range(3):
    print(' tile %d' % j)
    for k, n in enumerate(NAMES):
        print('   %-30s %+8d' % (n, p[64 + 16 * j + k] - base))
