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

__global__ void sqnorm_partial(const float* __restrict__ g, size_t n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double a = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        double v = g[i];
        a += v * v;
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// state: [0] step (as double), [1] lr_t, [2] clip scale, [3] global norm, [4] extra sq-norm
__global__ void adam_prepare(const double* __restrict__ partial, int nblk, double* __restrict__ state,
                             float lr, float b1, float b2, float clip, float grad_scale,
                             int staircase_decay) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < nblk; ++i) s += partial[i];
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

__global__ void adam_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, const double* __restrict__ state,
                            float b1, float b2, float eps) {
    const float lr_t = (float)state[1], gs = (float)state[2];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gs;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" size_t d2p_adam_ws_bytes(void) { return 4 * kNumSMs * sizeof(double); }

// state: 8 doubles on the device, zero-initialised by the caller before step 1.
extern "C" int d2p_clip_adam_step(float* params, const float* grads, float* m, float* v, size_t n,
                                  float lr, float b1, float b2, float eps, float clip_norm,
                                  float grad_scale, int staircase_decay_steps, double* state,
                                  void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(params && grads && m && v && state && ws, "adam: null buffer");
    D2P_REQUIRE(ws_bytes >= d2p_adam_ws_bytes(), "adam: workspace too small");
    const int nblk = 4 * kNumSMs;
    sqnorm_partial<<<nblk, 256, 0, st>>>(grads, n, (double*)ws);
    D2P_CHECK_LAUNCH();
    adam_prepare<<<1, 32, 0, st>>>((const double*)ws, nblk, state, lr, b1, b2, clip_norm, grad_scale,
                                   staircase_decay_steps);
    D2P_CHECK_LAUNCH();
    adam_update<<<4 * kNumSMs, 256, 0, st>>>(params, grads, m, v, n, state, b1, b2, eps);
    D2P_CHECK_LAUNCH();
    return 0;
}
