"""Developer tool (GPU): SM-clock timeline of one step of the persistent LSTM kernels
(CTA (0,0), step index 5 of the sequence; see pstamp() in csrc/lstm_persist.cu)."""
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
lib.d2p_lstm_set_persistent(1)
FW = ['top', 'grid barrier passed', 'last bulk copy issued', 'last MMA issued', None, 'accum ready',
      'cell math done', 'packed h stores issued', 'arrive issued', 'copy-out stores issued', 'early k-block issued', 'next step top', 'end-of-step CTA barrier passed', 'last warp: copy-out stores issued']
BW = ['top', None, 'cluster partials summed (DSMEM)', 'packed dZ stores issued', 'publish (arrive) issued',
      'row-tile barrier passed', 'accum ready', 'partial tile parked in ring', 'cluster barrier passed', 'next step top']

for (T, R) in [(20, 320), (50, 32)]:
    In = 512
    X = torch.randn(T, R, In, device=dev) * 0.1
    W = torch.randn(In + H, 4 * H, device=dev) * 0.05
    b = torch.zeros(4 * H, device=dev)
    ln = torch.full((R,), T, dtype=torch.int32, device=dev)
    z = lambda *s: torch.zeros(*s, device=dev)
    Y, hT, cT, gates, cells = z(T, R, H), z(R, H), z(R, H), z(T, R, 4 * H), z(T, R, H)
    dY, dX, dW, db, dh0, dc0 = torch.randn(T, R, H, device=dev), z(T, R, In), z(In + H, 4 * H), z(4 * H), z(R, H), z(R, H)
    wsb = lib.d2p_lstm_seq_bwd_ws_bytes(T, R, H)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    probe = torch.zeros(256 + 32 * 160, dtype=torch.int64, device=dev)

    def fwd():
        check(lib.d2p_lstm_seq_fwd(ptr(X), T, R, In, H, ptr(ln), None, None, ptr(W), ptr(b), 1.0, ptr(Y), ptr(hT),
                                   ptr(cT), ptr(gates), ptr(cells), 3, st), 'fwd')

    def bwd():
        check(lib.d2p_lstm_seq_bwd(ptr(X), T, R, In, H, ptr(ln), None, None, ptr(W), ptr(Y), ptr(gates), ptr(cells),
                                   ptr(dY), None, None, ptr(dX), ptr(dW), ptr(db), ptr(dh0), ptr(dc0), ptr(ws), wsb,
                                   1, st), 'bwd')

    for _ in range(2):
        fwd(); bwd()
    torch.cuda.synchronize()
    lib.d2p_debug_set_probe(ptr(probe))
    fwd()
    torch.cuda.synchronize()
    pf = probe.cpu().tolist()[64:]
    allf = probe.cpu().numpy()[256:].reshape(-1, 32).copy()
    bwd()
    torch.cuda.synchronize()
    allb = probe.cpu().numpy()[256:].reshape(-1, 32).copy()
    lib.d2p_debug_set_probe(None)
    ncta = 32 * ((R + 127) // 128)
    import numpy as np
    def col(a, k): return (a[:ncta, k] - a[:ncta, 0]).astype(np.int64)
    print('== all %d CTAs, forward step 5 (cycles after the CTA\'s own loop top): min / median / max' % ncta)
    for k, n in ((2, 'last bulk copy issued'), (3, 'last MMA issued'), (5, 'accum ready'), (8, 'arrive issued'), (14, 'outputs staged'), (15, 'staging consumed'), (17, 'row-tile barrier passed'), (10, 'next head issued'), (11, 'next step top')):
        c = col(allf, k)
        print('   %-24s %7d %7d %7d   slowest CTA (x,y) = (%d,%d)' % (n, c.min(), np.median(c), c.max(), int(c.argmax()) % 32, int(c.argmax()) // 32))
    print('== all CTAs, backward step 5')
    for k, n in ((20, 'publish issued'), (21, 'row-tile barrier passed'), (22, 'accum ready'), (24, 'cluster barrier passed'), (25, 'next step top')):
        c = (allb[:ncta, k] - allb[:ncta, 16]).astype(np.int64)
        print('   %-24s %7d %7d %7d   slowest CTA (x,y) = (%d,%d)' % (n, c.min(), np.median(c), c.max(), int(c.argmax()) % 32, int(c.argmax()) // 32))
    p = pf[:16] + probe.cpu().tolist()[64 + 16:]
    print('== T=%d R=%d forward step 5, CTA(0,0), SM cycles from loop top' % (T, R))
    for i, n in enumerate(FW):
        if n:
            print('   %-28s +%d' % (n, p[i] - p[0]))
    print('== backward step 5')
    for i, n in enumerate(BW):
        if n:
            print('   %-34s +%d' % (n, p[16 + i] - p[16]))
