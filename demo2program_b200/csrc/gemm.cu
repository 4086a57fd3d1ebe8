// GEMM engine for libd2p.
//
// fp32 row-major C[M,N] = alpha * op(A) * op(B) + beta * C (+ bias[N]).
// This is the exact-fp32 SIMT engine: 64x64x16 tiles, 256 threads, 4x4
// register micro-tiles, shared-memory staged with 128-bit loads where the
// operand layout allows.  It is the numerically conservative engine used for
// every contraction on the hot path (LSTM gate GEMMs, projections, the RN-pool
// FCs and all of their backward dX/dW products).  The tensor-core engine
// (tcgen05, gemm_tc.cu) is selected per call site where parity bounds allow.
#include "common.cuh"

namespace d2p {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

// Loads a BMxBK (or BKxBN) tile into shared memory as [BK][BM|BN] (k-major)
// so the inner product loop reads contiguous floats per thread.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
             const float* __restrict__ B, int ldb, float beta, float* __restrict__ C,
             int ldc, const float* __restrict__ bias) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, each 4x4 outputs
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- stage A tile: As[kk][mm] = opA(m0+mm, k0+kk) ----
        if (!TA) {
            // A is [M, K] row-major: consecutive threads along k
            for (int idx = tid; idx < BM * BK; idx += 256) {
                int mm = idx / BK, kk = idx % BK;
                int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
            }
        } else {
            // A stored [K, M]: consecutive threads along m
            for (int idx = tid; idx < BM * BK; idx += 256) {
                int kk = idx / BM, mm = idx % BM;
                int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < M && k < K) ? A[(size_t)k * lda + m] : 0.f;
            }
        }
        // ---- stage B tile: Bs[kk][nn] = opB(k0+kk, n0+nn) ----
        if (!TB) {
            for (int idx = tid; idx < BN * BK; idx += 256) {
                int kk = idx / BN, nn = idx % BN;
                int n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < N && k < K) ? B[(size_t)k * ldb + n] : 0.f;
            }
        } else {
            // B stored [N, K]
            for (int idx = tid; idx < BN * BK; idx += 256) {
                int nn = idx / BK, kk = idx % BK;
                int n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < N && k < K) ? B[(size_t)n * ldb + k] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = alpha * acc[i][j];
            if (bias) v += bias[n];
            float* c = C + (size_t)m * ldc + n;
            if (beta != 0.f) v += beta * *c;
            *c = v;
        }
    }
}

}  // namespace

int gemm(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha,
         const float* A, int lda, const float* B, int ldb, float beta, float* C,
         int ldc, const float* bias, int flags) {
    if (M <= 0 || N <= 0) return 0;
    D2P_REQUIRE(K >= 0 && A && B && C, "gemm: bad arguments");
    if (tc_eligible(M, N, K))   // tensor-core engine (tcgen05, bf16x3)
        return gemm_tc_auto(st, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, flags);
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    if (!ta && !tb)
        sgemm_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    else if (!ta && tb)
        sgemm_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    else if (ta && !tb)
        sgemm_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    else
        sgemm_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace d2p

extern "C" int d2p_gemm(int transA, int transB, int M, int N, int K, float alpha,
                        const float* A, int lda, const float* B, int ldb, float beta,
                        float* C, int ldc, const float* bias, void* stream) {
    return d2p::gemm((cudaStream_t)stream, transA != 0, transB != 0, M, N, K, alpha, A,
                     lda, B, ldb, beta, C, ldc, bias, 0);
}
