"""Developer tool (GPU): the hoisted LSTM gate GEMM [T*R, H] x [H, 4H] on the
tensor-core engine with pre-packed operands (for ncu captures and event timing)."""
import sys
sys.path.insert(0, '.')
import torch
from demo2program_b200 import _lib
from demo2program_b200._lib import ptr

lib = _lib.load()
dev = 'cuda:0'
st = torch.cuda.current_stream().cuda_stream
M, N, K = 6400, 2048, 512
# the persistent tile-queue kernel keeps its tile counter in the tensor-core arena
scratch = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)
cache = torch.zeros(16 << 20, dtype=torch.uint8, device=dev)
lib.d2p_tc_configure(ptr(scratch), scratch.numel(), ptr(cache), cache.numel(), 1)
lib.d2p_gemm_set_persistent(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
apk = torch.empty(lib.d2p_packed_bytes(M, K), dtype=torch.uint8, device=dev)
bpk = torch.empty(lib.d2p_packed_bytes(N, K), dtype=torch.uint8, device=dev)
lib.d2p_pack_bf16(ptr(A), M, K, K, 1, ptr(apk), st)
lib.d2p_pack_bf16(ptr(B), N, K, K, 1, ptr(bpk), st)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    lib.d2p_gemm_tc_packed(ptr(apk), ptr(bpk), M, N, K, 1.0, 0.0, ptr(C), N, None, 1, None, st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    lib.d2p_gemm_tc_packed(ptr(apk), ptr(bpk), M, N, K, 1.0, 0.0, ptr(C), N, None, 1, None, st)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / n
print('tensor-core GEMM (persistent=%s) %dx%dx%d:' % (sys.argv[2] if len(sys.argv) > 2 else '1', M, N, K) + ' %.1f us, %.1f TFLOP/s algorithmic (x3 on the tensor pipe = %.0f)' % (
    us, 2.0 * M * N * K / us / 1e6, 6.0 * M * N * K / us / 1e6))
ref = A.double() @ B.double().t()
print('rel err vs fp64: %.2e' % ((C.double() - ref).abs().max() / ref.abs().max()).item())
