// Global-norm clip + Adam over the flat parameter buffer:
// tf.contrib.layers.optimize_loss(optimizer=AdamOptimizer, clip_gradients=20.0)
// (reference trainer.py:102-109):
//   g <- g * clip / max(||g||, clip)
//   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) ; theta <- theta - lr_t * m / (sqrt(v) + eps)
// (epsilon outside the bias correction, unlike torch.optim.Adam).
// One launch covers every variable because parameters, gradients and both
// slots live in single flat buffers; `grad_scale` (1/world after the NCCL sum)
// is folded in.  The step counter lives on the device so the whole train step
// can be replayed from a CUDA graph.
#include "common.cuh"

namespace d2p {
namespace {

constexpr int AD_THREADS = 256, AD_BLOCKS = 8 * kNumSMs;

__device__ __forceinline__ double block_sum_d(double a, double* sh) {
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = AD_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    return sh[0];
}

// 16-byte loads, fixed block -> element assignment and a fixed in-block tree: deterministic
__global__ void __launch_bounds__(AD_THREADS)
sqnorm_partial(const float* __restrict__ g, size_t n, double* __restrict__ partial) {
    __shared__ double sh[AD_THREADS];
    const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    double a = 0.0;
    int cnt = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = *reinterpret_cast<const float4*>(g + 4 * i);
        a0 = fmaf(x.x, x.x, a0); a1 = fmaf(x.y, x.y, a1); a2 = fmaf(x.z, x.z, a2); a3 = fmaf(x.w, x.w, a3);
        if (++cnt == 8) { a += (double)a0 + (double)a1 + (double)a2 + (double)a3; a0 = a1 = a2 = a3 = 0.f; cnt = 0; }
    }
    a += (double)a0 + (double)a1 + (double)a2 + (double)a3;
    if (blockIdx.x == 0)
        for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) a += (double)g[i] * (double)g[i];
    const double s = block_sum_d(a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// state: [0] step (as double), [1] lr_t, [2] clip scale, [3] global norm, [4] extra sq-norm,
// [5] skip flag: a persistent kernel's step barrier timed out (sticky error words e0 / e1), the
// gradients are garbage - the update is skipped and the step counter does not advance
__global__ void __launch_bounds__(AD_THREADS)
adam_prepare(const double* __restrict__ partial, int nblk, double* __restrict__ state,
             float lr, float b1, float b2, float clip, float grad_scale, int staircase_decay,
             const unsigned* __restrict__ e0, const unsigned* __restrict__ e1) {
    __shared__ double sh[AD_THREADS];
    double a = 0.0;
    for (int i = threadIdx.x; i < nblk; i += AD_THREADS) a += partial[i];
    double s = block_sum_d(a, sh);
    if (threadIdx.x != 0) return;
    if ((e0 && *e0) || (e1 && *e1)) { state[5] = 1.0; return; }
    state[5] = 0.0;
    s += state[4];
    double norm = sqrt(s) * (double)grad_scale;
    double step0 = state[0];
    double lr_eff = lr;
    if (staircase_decay > 0) lr_eff *= pow(0.5, floor(step0 / (double)staircase_decay));
    double t = step0 + 1.0;
    state[0] = t;
    state[1] = lr_eff * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t));
    state[2] = (clip > 0.f ? (double)clip / fmax(norm, (double)clip) : 1.0) * (double)grad_scale;
    state[3] = norm;
    state[4] = 0.0;
}

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float gs, float lr_t,
                                          float b1, float b2, float eps) {
    const float gi = g * gs;
    m = b1 * m + (1.f - b1) * gi;
    v = b2 * v + (1.f - b2) * gi * gi;
    p -= lr_t * m / (sqrtf(v) + eps);
}

// 16-byte loads/stores of p, g, m, v; two vectors per thread in flight
__global__ void __launch_bounds__(AD_THREADS)
adam_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, size_t n, const double* __restrict__ state,
            float b1, float b2, float eps) {
    if (state[5] != 0.0) return;   // failed step: leave parameters and both slots untouched
    const float lr_t = (float)state[1], gs = (float)state[2];
    const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        const size_t j = i + stride;
        const bool two = j < n4;
        float4 p0 = *reinterpret_cast<float4*>(p + 4 * i), g0 = *reinterpret_cast<const float4*>(g + 4 * i);
        float4 m0 = *reinterpret_cast<float4*>(m + 4 * i), v0 = *reinterpret_cast<float4*>(v + 4 * i);
        float4 p1, g1, m1, v1;
        if (two) {
            p1 = *reinterpret_cast<float4*>(p + 4 * j); g1 = *reinterpret_cast<const float4*>(g + 4 * j);
            m1 = *reinterpret_cast<float4*>(m + 4 * j); v1 = *reinterpret_cast<float4*>(v + 4 * j);
        }
        adam_elem(p0.x, g0.x, m0.x, v0.x, gs, lr_t, b1, b2, eps); adam_elem(p0.y, g0.y, m0.y, v0.y, gs, lr_t, b1, b2, eps);
        adam_elem(p0.z, g0.z, m0.z, v0.z, gs, lr_t, b1, b2, eps); adam_elem(p0.w, g0.w, m0.w, v0.w, gs, lr_t, b1, b2, eps);
        *reinterpret_cast<float4*>(p + 4 * i) = p0; *reinterpret_cast<float4*>(m + 4 * i) = m0;
        *reinterpret_cast<float4*>(v + 4 * i) = v0;
        if (two) {
            adam_elem(p1.x, g1.x, m1.x, v1.x, gs, lr_t, b1, b2, eps); adam_elem(p1.y, g1.y, m1.y, v1.y, gs, lr_t, b1, b2, eps);
            adam_elem(p1.z, g1.z, m1.z, v1.z, gs, lr_t, b1, b2, eps); adam_elem(p1.w, g1.w, m1.w, v1.w, gs, lr_t, b1, b2, eps);
            *reinterpret_cast<float4*>(p + 4 * j) = p1; *reinterpret_cast<float4*>(m + 4 * j) = m1;
            *reinterpret_cast<float4*>(v + 4 * j) = v1;
        }
    }
    if (blockIdx.x == 0)
        for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x)
            adam_elem(p[i], g[i], m[i], v[i], gs, lr_t, b1, b2, eps);
}

}  // namespace
}  // namespace d2p

namespace d2p { int device_error_addrs(unsigned** out); }
using namespace d2p;

extern "C" size_t d2p_adam_ws_bytes(void) { return (size_t)AD_BLOCKS * sizeof(double); }

// state: 8 doubles on the device, zero-initialised by the caller before step 1.
extern "C" int d2p_clip_adam_step(float* params, const float* grads, float* m, float* v, size_t n,
                                  float lr, float b1, float b2, float eps, float clip_norm,
                                  float grad_scale, int staircase_decay_steps, double* state,
                                  void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(params && grads && m && v && state && ws, "adam: null buffer");
    D2P_REQUIRE(ws_bytes >= d2p_adam_ws_bytes(), "adam: workspace too small");
    D2P_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                "adam: buffers must be 16-byte aligned");
    const int nblk = AD_BLOCKS;
    unsigned* errw[2];
    D2P_TRY(device_error_addrs(errw));
    sqnorm_partial<<<nblk, AD_THREADS, 0, st>>>(grads, n, (double*)ws);
    D2P_CHECK_LAUNCH();
    adam_prepare<<<1, AD_THREADS, 0, st>>>((const double*)ws, nblk, state, lr, b1, b2, clip_norm, grad_scale,
                                           staircase_decay_steps, errw[0], errw[1]);
    D2P_CHECK_LAUNCH();
    adam_update<<<AD_BLOCKS, AD_THREADS, 0, st>>>(params, grads, m, v, n, state, b1, b2, eps);
    D2P_CHECK_LAUNCH();
    return 0;
}
