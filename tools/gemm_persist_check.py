"""Developer check (GPU): persistent tile-queue GEMM kernel (gemm_tc_persist_kernel) against
fp64 matmul and against the one-CTA-per-tile kernel, with timing of the packed product alone."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check

lib = _lib.load()
dev = 'cuda:0'
st = torch.cuda.current_stream().cuda_stream
scratch = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
cache = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)
lib.d2p_tc_configure(ptr(scratch), scratch.numel(), ptr(cache), cache.numel(), 1)


def run(M, N, K, mode, alpha=1.0, beta=0.0, bias=False, reps=0):
    g = torch.Generator(device='cpu').manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)          # Op_B[n, k]
    C0 = torch.randn(M, N, generator=g).to(dev)
    bv = torch.randn(N, generator=g).to(dev) if bias else None
    apk = torch.zeros(lib.d2p_packed_bytes(M, K), dtype=torch.uint8, device=dev)
    bpk = torch.zeros(lib.d2p_packed_bytes(N, K), dtype=torch.uint8, device=dev)
    check(lib.d2p_pack_bf16(ptr(A), M, K, K, 1, ptr(apk), st), 'pack A')
    check(lib.d2p_pack_bf16(ptr(B), N, K, K, 1, ptr(bpk), st), 'pack B')
    lib.d2p_gemm_set_persistent(mode)
    C = C0.clone()
    call = lambda: check(lib.d2p_gemm_tc_packed(ptr(apk), ptr(bpk), M, N, K, alpha, beta, ptr(C), N, ptr(bv), 1,
                                                None, st), 'gemm')
    call()
    torch.cuda.synchronize()
    ref = alpha * (A.double() @ B.double().t()) + beta * C0.double()
    if bias:
        ref = ref + bv.double()
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    us = 0.0
    if reps:
        beta0 = beta
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
    lib.d2p_gemm_set_persistent(1)
    return err, us, C


bad = 0
for (M, N, K) in [(6400, 2048, 512), (6400, 512, 2048), (6400, 2048, 48), (3200, 1024, 512), (6333, 2048, 512),
                  (2560, 1024, 64), (19000, 128, 192), (1600, 2048, 512)]:
    e0, t0, c0 = run(M, N, K, 0, reps=20)
    e1, t1, c1 = run(M, N, K, 1, reps=20)
    e2, _, _ = run(M, N, K, 1, alpha=0.5, beta=1.0, bias=True)
    e3, t3, c3 = run(M, N, K, 3, reps=20)                       # 128 x 256 tiles where they fit
    e4, _, _ = run(M, N, K, 3, alpha=0.5, beta=1.0, bias=True)
    d = (c0 - c1).abs().max().item()
    flag = '' if max(e1, e2, e3, e4) < 2e-5 else '   <<<< BAD'
    bad += bool(flag)
    fl = 2.0 * M * N * K
    print('M%6d N%5d K%5d  per-tile %.2e %6.1f us (%5.1f TF/s) | persistent %.2e %6.1f us (%5.1f TF/s) | beta/bias %.2e | '
          'max |diff| %.1e | wide tiles %.2e / %.2e %6.1f us (%5.1f TF/s)%s' % (
              M, N, K, e0, t0, fl / t0 / 1e6, e1, t1, fl / t1 / 1e6, e2, d, e3, e4, t3, fl / t3 / 1e6, flag))
print('GEMM_PERSIST_CHECK', 'OK' if not bad else 'FAILED')
sys.exit(1 if bad else 0)
