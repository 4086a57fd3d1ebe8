// Shared helpers for libd2p (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/d2p.h"

namespace d2p {

// thread-local last error text, returned by d2p_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);
void count_launch();

#define D2P_CHECK_CUDA(expr)                                                     \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess)                                                   \
            return ::d2p::fail(D2P_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, \
                               #expr, cudaGetErrorString(_e));                   \
    } while (0)

// every kernel launch is followed by exactly one D2P_CHECK_LAUNCH(): it also
// feeds the launch counter behind d2p_launch_count()
#define D2P_CHECK_LAUNCH()                   \
    do {                                     \
        ::d2p::count_launch();               \
        D2P_CHECK_CUDA(cudaGetLastError());  \
    } while (0)

#define D2P_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return ::d2p::fail(D2P_ERR_ARG, __VA_ARGS__);    \
    } while (0)

#define D2P_TRY(expr)                 \
    do {                              \
        int _r = (expr);              \
        if (_r != 0) return _r;       \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// leaky relu of the reference (models/ops.py:7-11): 0.6x + 0.4|x|
__device__ __forceinline__ float lrelu_f(float x) { return 0.6f * x + 0.4f * fabsf(x); }
// derivative recovered from the activation value a = lrelu(z): sign(a) = sign(z);
// at exactly 0 the reference's abs() has gradient 0, so the slope is 0.6.
__device__ __forceinline__ float lrelu_grad_from_out(float a) {
    return a > 0.f ? 1.0f : (a < 0.f ? 0.2f : 0.6f);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
// Short-latency forms for the recurrent kernels, whose few resident warps cannot hide
// long dependent instruction chains: ex2.approx + rcp (abs. error ~1e-7, i.e. fp32 rounding level)
__device__ __forceinline__ float sigmoid_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - 2.0f * __frcp_rn(1.0f + __expf(2.0f * x)); }

// internal engine entry points (gemm.cu / gemm_tc.cu)
enum { GEMM_CONST_A = 1, GEMM_CONST_B = 2 };   // operand is a weight: pack once per step
int gemm(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha,
         const float* A, int lda, const float* B, int ldb, float beta, float* C,
         int ldc, const float* bias = nullptr, int flags = 0);
bool tc_eligible(int M, int N, int K);
int gemm_tc_auto(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha, const float* A,
                 int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
                 int flags);

}  // namespace d2p
