"""Developer tool (GPU): the tensor-core convolution kernels (csrc/conv_tc.cu) against the fp32 CUDA-core
kernels on the ViZDoom geometry, one d2p_conv_set_tc bit at a time (forward / dX / dW), then timing of the
encoder forward + backward at the C4 size."""
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
from demo2program_b200 import _lib
from demo2program_b200.config import vizdoom_config, karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

lib = _lib.load()


def rel(a, b):
    d = float(np.abs(a - b).max())
    s = float(np.abs(b).max())
    return d / s if s > 0 else d


def run(cfg, batch, mode):
    lib.d2p_conv_set_tc(mode)
    lib.d2p_conv_set_fused(0)
    try:
        eng = Engine(cfg, use_graph=False)
        eng.stage_batch(batch)
        eng.forward()
        eng.backward()
        torch.cuda.synchronize()
        eng.check_device()
        return eng, {'loss': float(eng.loss[0]), 'feat': eng.feat.cpu().numpy().copy(),
                     'saved': eng.conv_saved.cpu().numpy().copy(), 'grads': eng.grads.cpu().numpy().copy()}
    finally:
        lib.d2p_conv_set_tc(7)
        lib.d2p_conv_set_fused(1)


def compare(cfg, tag):
    batch = make_batch(cfg, seed=3)
    eng, ref = run(cfg, batch, 16)      # every convolution on the fp32 CUDA-core per-layer kernels
    names = [e for e in eng.pm if 'conv' in e.name.lower() or 'State_Encoder' in e.name]
    print('== %s: conv variables %s' % (tag, [e.name.split('/')[-2] + '/' + e.name.split('/')[-1] for e in names][:6]))
    for mode, label in ((0, 'rgb fwd'), (1 | 16, 'fwd'), (2 | 16, 'dx (quad form)'), (2 | 16 | 32, 'dx (per class)'), (4 | 16, 'dw'), (7, 'all')):
        try:
            _, out = run(cfg, batch, mode)
        except Exception as ex:   # noqa: BLE001
            print('mode %-20s FAILED: %s' % (label, ex))
            continue
        worst = max(((rel(out['grads'][e.offset:e.offset + e.size], ref['grads'][e.offset:e.offset + e.size]), e.name)
                     for e in names), default=(0, ''))
        print('mode %-20s loss %.6f (ref %.6f)  feat %.2e  saved %.2e  all grads %.2e  worst conv grad %.2e %s' % (
            label, out['loss'], ref['loss'], rel(out['feat'], ref['feat']), rel(out['saved'], ref['saved']),
            rel(out['grads'], ref['grads']), worst[0], worst[1]))
        if mode in (17, 0):
            # per-layer activations / statistics inside `saved`
            d = eng.conv_desc
            off, ih, iw = 0, cfg.h, cfg.w
            N = cfg.batch_size * cfg.k * cfg.max_demo_len
            for l in range(d.n_layers):
                oh, ow, c = (ih + 1) // 2, (iw + 1) // 2, d.layers[l].cout
                na, ns = N * oh * ow * c, 4 * cfg.k * c
                a0, a1 = ref['saved'][off:off + na], out['saved'][off:off + na]
                s0, s1 = ref['saved'][off + na:off + na + ns].reshape(4, -1), out['saved'][off + na:off + na + ns].reshape(4, -1)
                flips = int(((a0 > 0) != (a1 > 0)).sum() + ((a0 < 0) != (a1 < 0)).sum()) // 1
                print('      layer %d: %d of %d activations change sign (lrelu slope) between the two paths; |a| < 1e-4 max: %d' % (
                    l + 1, flips, a0.size, int((np.abs(a0) < 1e-4 * np.abs(a0).max()).sum())))
                bad = np.argwhere(np.abs(a1 - a0) > 1e-4 * np.abs(a0).max())
                print('      layer %d: act %.2e (max %.3g, %d elements off by > 1e-4 max%s)  mean %.2e rstd %.2e scale %.2e shift %.2e' % (
                    l + 1, rel(a1, a0), np.abs(a0).max(), len(bad), (', first %s' % bad[:3].ravel().tolist()) if len(bad) else '',
                    rel(s1[0], s0[0]), rel(s1[1], s0[1]), rel(s1[2], s0[2]), rel(s1[3], s0[3])))
                off += na + ns
                ih, iw = oh, ow
            gmax = np.abs(ref['grads']).max()
            for e in names:
                a, b = out['grads'][e.offset:e.offset + e.size], ref['grads'][e.offset:e.offset + e.size]
                print('      %-62s err %.2e  max|g| %.2e  (model max %.2e)' % (e.name, np.abs(a - b).max(), np.abs(b).max(), gmax))
        if mode in (20,):
            for e in names:
                if e.name.endswith('weights'):
                    print('      %-60s %.2e' % (e.name, rel(out['grads'][e.offset:e.offset + e.size],
                                                               ref['grads'][e.offset:e.offset + e.size])))


def timing(cfg, tag, modes=(16, 7)):
    """conv encoder forward / backward alone (CUDA events, warm), tensor-core kernels vs CUDA-core kernels"""
    import ctypes as C
    from demo2program_b200._lib import ptr, check
    batch = make_batch(cfg, seed=3)
    eng = Engine(cfg, use_graph=False, concurrent=False)
    eng.stage_batch(batch)
    eng.forward(); eng.backward()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream

    def fwd():
        check(lib.d2p_conv_encoder_fwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.feat), ptr(eng.conv_saved),
                                       1, ptr(eng.ws), eng.ws_bytes, st), 'conv fwd')

    def bwd():
        check(lib.d2p_conv_encoder_bwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.dfeat), ptr(eng.conv_saved),
                                       1, ptr(eng.ws), eng.ws_bytes, st), 'conv bwd')

    def t_us(fn, n=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / n

    for mode in modes:
        lib.d2p_conv_set_tc(mode)
        print('%s conv_set_tc(%d): encoder fwd %.1f us, bwd %.1f us' % (tag, mode, t_us(fwd), t_us(bwd)))
    lib.d2p_conv_set_tc(7)


if __name__ == '__main__':
    torch.cuda.set_device(0)
    compare(vizdoom_config('full', batch_size=2, k=2, max_demo_len=3, test_k=2, max_program_len=8), 'vizdoom B2 k2 T3')
    compare(vizdoom_config('full', batch_size=8, k=3, max_demo_len=8, test_k=2, max_program_len=8), 'vizdoom B8 k3 T8')
    import os
    os.environ['D2P_CONV_TC_GRID'] = '16'
    compare(vizdoom_config('full', batch_size=2, k=2, max_demo_len=3, test_k=2, max_program_len=8), 'vizdoom B2 k2 T3, 16 CTAs')
    del os.environ['D2P_CONV_TC_GRID']
    compare(karel_config('full', batch_size=8, k=3), 'karel B8 k3 (per-layer path)')
    if '--time' in sys.argv:
        timing(vizdoom_config('full', batch_size=32, k=10), 'C4')
