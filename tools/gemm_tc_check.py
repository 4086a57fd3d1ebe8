"""Developer check (GPU): tcgen05 bf16x3 GEMM vs fp64 matmul, all transposes."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr, check

lib = _lib.load()
dev = 'cuda:0'
st = torch.cuda.current_stream().cuda_stream


def run(M, N, K, ta, tb, alpha=1.0, beta=0.0, bias=False, tc=True):
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).to(dev)
    B = torch.randn((N, K) if tb else (K, N), generator=g).to(dev)
    C0 = torch.randn(M, N, generator=g).to(dev)
    bv = torch.randn(N, generator=g).to(dev) if bias else None
    C = C0.clone()
    ws_b = lib.d2p_gemm_tc_ws_bytes(M, N, K)
    ws = torch.empty(ws_b, dtype=torch.uint8, device=dev)
    if tc:
        check(lib.d2p_gemm_tc(int(ta), int(tb), M, N, K, alpha, ptr(A), A.shape[1], ptr(B), B.shape[1],
                              beta, ptr(C), N, ptr(bv), ptr(ws), ws_b, st), 'gemm_tc')
    else:
        check(lib.d2p_gemm(int(ta), int(tb), M, N, K, alpha, ptr(A), A.shape[1], ptr(B), B.shape[1],
                           beta, ptr(C), N, ptr(bv), st), 'gemm')
    torch.cuda.synchronize()
    Ad = (A.double().t() if ta else A.double())
    Bd = (B.double().t() if tb else B.double())
    ref = alpha * (Ad @ Bd) + beta * C0.double()
    if bias:
        ref = ref + bv.double()
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    return err


if __name__ == '__main__':
    bad = 0
    for (M, N, K) in [(128, 128, 32), (128, 128, 64), (128, 64, 96), (256, 256, 512), (320, 2048, 512),
                      (100, 72, 40), (6400, 2048, 48), (512, 2048, 700), (33, 50, 512), (640, 512, 2048)]:
        for ta in (0, 1):
            for tb in (0, 1):
                e = run(M, N, K, ta, tb)
                e2 = run(M, N, K, ta, tb, alpha=0.5, beta=1.0, bias=True)
                es = run(M, N, K, ta, tb, tc=False)
                flag = '' if max(e, e2) < 1e-4 else '  <<<< BAD'
                bad += bool(flag)
                print('M%5d N%5d K%5d ta%d tb%d  tc err %.2e  (beta/bias) %.2e   simt err %.2e%s' % (
                    M, N, K, ta, tb, e, e2, es, flag))
    # timing
    for (M, N, K, ta, tb) in [(6400, 2048, 512, 0, 0), (6400, 512, 2048, 0, 1), (512, 2048, 6400, 1, 0),
                              (320, 2048, 512, 0, 0), (320, 512, 2048, 0, 1), (3200, 512, 512, 0, 0)]:
        A = torch.randn((K, M) if ta else (M, K), device=dev)
        B = torch.randn((N, K) if tb else (K, N), device=dev)
        C = torch.empty(M, N, device=dev)
        ws_b = lib.d2p_gemm_tc_ws_bytes(M, N, K)
        ws = torch.empty(ws_b, dtype=torch.uint8, device=dev)
        for name, fn in (('tc', lambda: lib.d2p_gemm_tc(ta, tb, M, N, K, 1.0, ptr(A), A.shape[1], ptr(B), B.shape[1], 0.0, ptr(C), N, None, ptr(ws), ws_b, st)),
                         ('simt', lambda: lib.d2p_gemm(ta, tb, M, N, K, 1.0, ptr(A), A.shape[1], ptr(B), B.shape[1], 0.0, ptr(C), N, None, st))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print('%-5s M%5d N%5d K%5d ta%d tb%d: %.1f us  %.1f TFLOP/s (incl. split passes for tc)' % (
                name, M, N, K, ta, tb, ms * 1e3, 2.0 * M * N * K / ms / 1e9))
    print('BAD =', bad)
