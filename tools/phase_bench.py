"""Developer tool (GPU): device time of the forward / forward+backward / full C2 step as CUDA
graphs, with concurrent branches on and off, persistent recurrences on and off."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch
from demo2program_b200 import _lib

lib = _lib.load()
cfg = karel_config('full', batch_size=32, k=10)
batch = make_batch(cfg, seed=123)


def timed(g, n=20):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for persist in (0, 1):
    for conc in (False, True):
        lib.d2p_lstm_set_persistent(persist)
        eng = Engine(cfg, use_graph=True, concurrent=conc)
        eng.stage_batch(batch)
        torch.cuda.synchronize()
        gf, nf = eng._capture(eng.forward)
        tf = timed(gf)

        def fb():
            eng.forward()
            eng.backward()
        gfb, nfb = eng._capture(fb)
        tfb = timed(gfb)
        gs, ns = eng._capture(lambda: eng._step_body(True))
        ts = timed(gs)
        print('persistent=%d concurrent=%d: forward %.3f ms (%d launches), fwd+bwd %.3f ms (%d), full step %.3f ms (%d)'
              % (persist, conc, tf, nf, tfb, nfb, ts, ns))
        del eng, gf, gfb, gs
