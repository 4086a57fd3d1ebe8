"""Developer tool (GPU): persistent LSTM recurrence kernels vs the per-step path.

Runs d2p_lstm_seq_fwd / d2p_lstm_seq_bwd through the C ABI with
d2p_lstm_set_persistent(0) and (1) on the same random inputs, prints the largest
differences and the device time of each variant (CUDA events, warm)."""
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from demo2program_b200 import _lib
from demo2program_b200._lib import check, ptr

lib = _lib.load()
dev = torch.device('cuda:0')
H = 512


def run(T, R, In, with_init, mode, seed=0, reps=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    f = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(dev)
    X = f(T, R, In)
    W = (torch.randn(In + H, 4 * H, generator=g) * 0.05).to(dev)
    b = f(4 * H) * 0.1
    lens = torch.randint(1, T + 1, (R,), generator=g, dtype=torch.int32).to(dev)
    lens[0] = T
    h0 = f(R, H) if with_init else None
    c0 = f(R, H) if with_init else None
    dY = f(T, R, H)
    dhT, dcT = f(R, H), f(R, H)
    z = lambda *s: torch.zeros(*s, device=dev)
    Y, hT, cT, gates, cells = z(T, R, H), z(R, H), z(R, H), z(T, R, 4 * H), z(T, R, H)
    dX, dW, db, dh0, dc0 = z(T, R, In), z(In + H, 4 * H), z(4 * H), z(R, H), z(R, H)
    wsb = lib.d2p_lstm_seq_bwd_ws_bytes(T, R, H)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    big = lib.d2p_gemm_tc_ws_bytes(T * R, 4 * H, 4 * H) + (8 << 20)
    scratch = torch.zeros(big, dtype=torch.uint8, device=dev)
    cache = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)
    lib.d2p_tc_configure(ptr(scratch), scratch.numel(), ptr(cache), cache.numel(), 1)
    lib.d2p_lstm_set_persistent(mode)
    st = torch.cuda.current_stream().cuda_stream

    def fwd(ph):
        check(lib.d2p_lstm_seq_fwd(ptr(X), T, R, In, H, ptr(lens), ptr(h0), ptr(c0), ptr(W), ptr(b),
                                   1.0, ptr(Y), ptr(hT), ptr(cT), ptr(gates), ptr(cells), ph, st), 'fwd')

    def bwd(ph):
        check(lib.d2p_lstm_seq_bwd(ptr(X), T, R, In, H, ptr(lens), ptr(h0), ptr(c0), ptr(W), ptr(Y), ptr(gates),
                                   ptr(cells), ptr(dY), ptr(dhT), ptr(dcT), ptr(dX), ptr(dW), ptr(db), ptr(dh0),
                                   ptr(dc0), ptr(ws), wsb, ph, st), 'bwd')

    al = lambda n: (n + 255) // 256 * 256

    def err_word(off):
        return int(scratch[off + 63 * 4: off + 64 * 4].view(torch.int32).item())

    fwd(3)
    torch.cuda.synchronize()
    hbytes = 8 * ((R + 7) // 8) * 2048 if R <= 32 else lib.d2p_packed_bytes(R, H)
    if mode and err_word(2 * al(hbytes)):
        print('  !! forward barrier timed out')
    out = {k: v.clone() for k, v in dict(Y=Y, hT=hT, cT=cT, gates=gates, cells=cells).items()}
    gates_saved = gates.clone()
    bwd(3)
    torch.cuda.synchronize()
    out.update({k: v.clone() for k, v in dict(dZ=gates, dX=dX, dW=dW, db=db, dh0=dh0, dc0=dc0).items()})
    times = {}
    if reps:
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        tf = tb = 0.0
        for i in range(reps + 2):
            fwd(1)
            e[0].record(); fwd(2); e[1].record()
            torch.cuda.synchronize()
            gates.copy_(gates_saved)
            e[2].record(); bwd(1); e[3].record()
            torch.cuda.synchronize()
            if i >= 2:
                tf += e[0].elapsed_time(e[1]); tb += e[2].elapsed_time(e[3])
        times = {'fwd_recur_us': 1e3 * tf / reps, 'bwd_recur_plus_dx_us': 1e3 * tb / reps}
    return out, times


def main():
    ok = True
    for (T, R, In, init) in [(20, 320, 48, False), (20, 320, 512, True), (50, 32, 512, True), (7, 130, 512, True), (6, 40, 512, True), (4, 120, 512, True), (5, 56, 512, False),
                             (3, 5, 512, False)]:
        a, ta = run(T, R, In, init, 0, reps=5)
        b, tb = run(T, R, In, init, 1, reps=5)
        worst = 0.0
        for k in a:
            d = float((a[k] - b[k]).abs().max())
            s = float(a[k].abs().max()) + 1e-12
            worst = max(worst, d / s)
            if not (d / s < 2e-4):
                ok = False
                print('  MISMATCH %s: max abs diff %.3e (scale %.3e)' % (k, d, s))
        print('T=%d R=%d In=%d init=%d: worst rel diff %.2e | per-step %s | persistent %s' %
              (T, R, In, init, worst, {k: round(v, 1) for k, v in ta.items()},
               {k: round(v, 1) for k, v in tb.items()}))
    print('PERSIST_CHECK', 'OK' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
