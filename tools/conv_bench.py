"""Developer tool (GPU): Karel conv encoder forward / backward alone, fused vs per-layer kernels
(CUDA events, warm, L2 not flushed), at C2 (6400 frames, train) and C5 (102400 frames, eval)."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check
from demo2program_b200.config import karel_config
from demo2program_b200.engine import Engine
from demo2program_b200.synthetic import make_batch

lib = _lib.load()


def t_us(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


for (B, k, train) in [(32, 10, True), (512, 10, False)]:
    cfg = karel_config('full', batch_size=B, k=k)
    eng = Engine(cfg, use_graph=False, concurrent=False, is_train=train)
    eng.stage_batch(make_batch(cfg, seed=1) if B <= 64 else {**make_batch(karel_config('full', batch_size=32, k=k), seed=1),
                                                              's_h': np.random.RandomState(0).randint(0, 2, (B, k, cfg.max_demo_len, 8, 8, 16)).astype(np.uint8),
                                                              'demo_len': np.full((B, k), cfg.max_demo_len, np.float32),
                                                              'program_len': np.full((B, 1), 10, np.float32),
                                                              'program_tokens': np.zeros((B, cfg.max_program_len), np.int32),
                                                              'a_h_tokens': np.zeros((B, k, cfg.max_demo_len), np.int32),
                                                              'per': np.zeros((B, k, cfg.max_demo_len, cfg.per_dim), np.float32)})
    st = torch.cuda.current_stream().cuda_stream
    frames = B * k * cfg.max_demo_len

    def fwd():
        check(lib.d2p_conv_encoder_fwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.feat), ptr(eng.conv_saved),
                                       int(train), ptr(eng.ws), eng.ws_bytes, st), 'conv fwd')

    def bwd():
        check(lib.d2p_conv_encoder_bwd(C.byref(eng.conv_desc), ptr(eng.d_frames), ptr(eng.dfeat), ptr(eng.conv_saved),
                                       int(train), ptr(eng.ws), eng.ws_bytes, st), 'conv bwd')

    for fused in (0, 1):
        lib.d2p_conv_set_fused(fused)
        tf = t_us(fwd)
        tb = t_us(bwd) if train else float('nan')
        alg_f, alg_fb = frames * 1216, frames * 2432
        print('B=%d k=%d train=%d frames=%d fused=%d: fwd %.1f us (%.1f GB/s of %d algorithmic bytes), bwd %.1f us'
              % (B, k, train, frames, fused, tf, alg_f / tf / 1e3, alg_f, tb))
    lib.d2p_conv_set_fused(1)
    del eng
