"""Developer tool (GPU): per-op device time of one eager C2 train step, warm
caches, CUDA events around every C-ABI call (multi-kernel ops include their
internal launch gaps)."""
import collections
import sys

sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

cfg = karel_config('full', batch_size=32, k=10)
eng = Engine(cfg, use_graph=False, concurrent=False)
batch = make_batch(cfg, seed=123)
eng.stage_batch(batch)
for _ in range(3):
    eng.train_step_device(True)
torch.cuda.synchronize()

records = []
orig_call = eng._call
tag = ['']


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(name, *args)
    e1.record()
    records.append((tag[0] + name, e0, e1))


eng._call = timed_call
orig_fwd, orig_bwd = eng._lstm_fwd, eng._lstm_bwd


def lf(X, Tn, Rn, In, lens, h0, c0, scope, b, phases=3):
    tag[0] = scope.split('/')[0] + ':p%d:' % phases
    orig_fwd(X, Tn, Rn, In, lens, h0, c0, scope, b, phases)
    tag[0] = ''


def lb(X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, dX):
    tag[0] = scope.split('/')[0] + ':'
    orig_bwd(X, Tn, Rn, In, lens, h0, c0, scope, b, dY, dhT, dcT, dX)
    tag[0] = ''


eng._lstm_fwd, eng._lstm_bwd = lf, lb
orig_bwd = eng._lstm_bwd_call


def lbc(*a):
    tag[0] = a[7].split('/')[0] + ':bwd%d:' % a[-1]
    orig_bwd(*a)
    tag[0] = ''


eng._lstm_bwd_call = lbc
eng._lstm_bwd = Engine._lstm_bwd.__get__(eng)
N = 5
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(N):
    eng.train_step_device(True)
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, a, b in records:
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + a.elapsed_time(b) * 1e3)
tot = sum(t for _, t in agg.values()) / N
print('eager step wall (events) %.1f us; sum of ops %.1f us' % (t0.elapsed_time(t1) * 1e3 / N, tot))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-55s calls/step %4d  us/step %9.1f  %5.1f%%' % (name, c // N, t / N, 100 * t / N / tot))
