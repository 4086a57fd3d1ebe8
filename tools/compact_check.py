"""Developer tool (GPU): compact (32-CTA, all row tiles per CTA) persistent recurrences against the
one-tile-per-CTA kernels: results, device time alone, and N recurrences side by side on N streams."""
import sys

sys.path.insert(0, '.')
import torch

from demo2program_b200 import _lib
from demo2program_b200._lib import check, ptr

lib = _lib.load()
dev = torch.device('cuda:0')
H = 512
COMPACT = 8


class Lstm:
    def __init__(self, T, R, In, seed, with_init=True):
        g = torch.Generator(device='cpu').manual_seed(seed)
        f = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(dev)
        self.T, self.R, self.In = T, R, In
        self.X = f(T, R, In)
        self.W = (torch.randn(In + H, 4 * H, generator=g) * 0.05).to(dev)
        self.b = f(4 * H) * 0.1
        self.lens = torch.randint(1, T + 1, (R,), generator=g, dtype=torch.int32).to(dev)
        self.lens[0] = T
        self.h0 = f(R, H) if with_init else None
        self.c0 = f(R, H) if with_init else None
        self.dY, self.dhT, self.dcT = f(T, R, H), f(R, H), f(R, H)
        z = lambda *s: torch.zeros(*s, device=dev)
        self.Y, self.hT, self.cT, self.gates, self.cells = z(T, R, H), z(R, H), z(R, H), z(T, R, 4 * H), z(T, R, H)
        self.dX, self.dW, self.db, self.dh0, self.dc0 = z(T, R, In), z(In + H, 4 * H), z(4 * H), z(R, H), z(R, H)
        self.wsb = lib.d2p_lstm_seq_bwd_ws_bytes(T, R, H)
        self.ws = torch.zeros(self.wsb, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.Stream()
        big = lib.d2p_gemm_tc_ws_bytes(T * R, 4 * H, 4 * H) + (8 << 20)
        self.arena = torch.zeros(big, dtype=torch.uint8, device=dev)
        lib.d2p_tc_bind_stream(self.stream.cuda_stream, ptr(self.arena), self.arena.numel())

    def fwd(self, ph, st=None):
        st = self.stream.cuda_stream if st is None else st
        check(lib.d2p_lstm_seq_fwd(ptr(self.X), self.T, self.R, self.In, H, ptr(self.lens), ptr(self.h0), ptr(self.c0),
                                   ptr(self.W), ptr(self.b), 1.0, ptr(self.Y), ptr(self.hT), ptr(self.cT),
                                   ptr(self.gates), ptr(self.cells), ph, st), 'fwd')

    def bwd(self, ph, st=None):
        st = self.stream.cuda_stream if st is None else st
        check(lib.d2p_lstm_seq_bwd(ptr(self.X), self.T, self.R, self.In, H, ptr(self.lens), ptr(self.h0), ptr(self.c0),
                                   ptr(self.W), ptr(self.Y), ptr(self.gates), ptr(self.cells), ptr(self.dY),
                                   ptr(self.dhT), ptr(self.dcT), ptr(self.dX), ptr(self.dW), ptr(self.db),
                                   ptr(self.dh0), ptr(self.dc0), ptr(self.ws), self.wsb, ph, st), 'bwd')

    def outputs(self):
        return {k: getattr(self, k).clone() for k in ('Y', 'hT', 'cT', 'gates', 'cells')}


def device_error():
    import ctypes
    f = ctypes.c_int(0)
    lib.d2p_device_error(ctypes.byref(f))
    return f.value


def main():
    scratch = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    cache = torch.zeros(128 << 20, dtype=torch.uint8, device=dev)
    lib.d2p_tc_configure(ptr(scratch), scratch.numel(), ptr(cache), cache.numel(), 1)
    ok = True
    for (T, R, In, init) in [(20, 320, 512, True), (7, 130, 48, False), (5, 200, 512, True), (4, 512, 512, True),
                             (6, 257, 512, False)]:
        res = {}
        for flag in (0, COMPACT):
            m = Lstm(T, R, In, seed=R + T, with_init=init)
            with torch.cuda.stream(m.stream):
                m.fwd(3 | flag)
            torch.cuda.synchronize()
            if device_error():
                print('  !! barrier time-out (flag %d)' % flag)
                ok = False
            res[flag] = m.outputs()
            gsaved = m.gates.clone()
            with torch.cuda.stream(m.stream):
                m.bwd(3 | flag)
            torch.cuda.synchronize()
            if device_error():
                print('  !! backward barrier time-out (flag %d)' % flag)
                ok = False
            res[flag].update({k: getattr(m, k).clone() for k in ('dX', 'dW', 'db', 'dh0', 'dc0')})
            res[flag]['dZ'] = m.gates.clone()
        worst = 0.0
        for k in res[0]:
            d = float((res[0][k] - res[COMPACT][k]).abs().max())
            s = float(res[0][k].abs().max()) + 1e-12
            worst = max(worst, d / s)
            if not d / s < 1e-5:
                ok = False
                print('  MISMATCH %s: %.3e (scale %.3e)' % (k, d, s))
        print('T=%d R=%d In=%d init=%d: compact vs one-tile-per-CTA worst rel diff %.2e' % (T, R, In, init, worst))
    # timing at the C2 shape: alone and N side by side
    T, R = 20, 320
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for which in ('fwd', 'bwd'):
        for flag, name in ((0, 'one tile per CTA (96 CTAs)'), (COMPACT, 'compact (32 CTAs)')):
            for n in ((1,) if flag == 0 else (1, 2, 3)):
                ms = [Lstm(T, R, 512, seed=i) for i in range(n)]
                tot = 0.0
                reps = 6
                for it in range(reps + 2):
                    for m in ms:
                        with torch.cuda.stream(m.stream):
                            m.fwd(1)
                            if which == 'bwd':
                                m.fwd(2 | flag)
                    flush.zero_()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    cur = torch.cuda.current_stream()
                    e0.record(cur)
                    for m in ms:
                        m.stream.wait_event(e0)
                        with torch.cuda.stream(m.stream):
                            if which == 'fwd':
                                m.fwd(2 | flag)
                            else:
                                m.bwd(1 | flag)
                        cur.wait_stream(m.stream)
                    e1.record(cur)
                    torch.cuda.synchronize()
                    if it >= 2:
                        tot += e0.elapsed_time(e1)
                print('%s recurrence R=320 T=20, %s, %d side by side: %.1f us' % (which, name, n, 1e3 * tot / reps))
                if device_error():
                    print('  !! barrier time-out'); ok = False
                del ms
    print('COMPACT_CHECK', 'OK' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
