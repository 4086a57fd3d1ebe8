// Pooled Luong attention for TRAINING the induction baseline: forward that keeps the attention
// weights, and the backward pass (reference models/baselines/model_induction.py:25-53
// _compute_attention, 107-182 PoolingAttentionWrapper.call, trained by trainer.py:102-109).
//
// For query row (b, j) and every seen demonstration i:
//   score[t'] = q . keys[t', b*k+i, :]        (t' < len[b,i]; masked to -inf beyond)
//   alpha     = softmax(score)
//   ctx_i     = sum_t' alpha[t'] values[t', b*k+i, :]
// The wrapper averages the k attention vectors [h; ctx_i] W_a; W_a is shared and linear, so the
// kernels work with ctx = mean_i ctx_i (see decode.cu) and its gradient dctx:
//   dctx_i    = dctx / k
//   dalpha    = values . dctx_i,  dscore = alpha * (dalpha - sum alpha dalpha)
//   dq       += sum_i sum_t' dscore[t'] keys[t']
//   dkeys[t'] += dscore[t'] q,    dvalues[t'] += alpha[t'] dctx_i
// One CTA per (demonstration i, batch element b) serves the test_k queries of b and OWNS the rows
// (t', b*k+i) of dkeys / dvalues: the decoder steps accumulate into them launch after launch with
// plain read-modify-write, in a fixed order (deterministic, no atomics).  Per-demonstration partial
// results ([B,k,tk,H]) are combined over i by a second kernel in increasing i.
// These are fp32 SIMT kernels: the induction baseline's training step is not a benchmarked
// configuration (BASELINE.json configs[4] is its greedy decode, decode.cu).
#include "common.cuh"

namespace d2p {
namespace {

constexpr int AT_MAX_Q = 8, AT_MAX_T = 64, AT_THREADS = 128, AT_MAX_H = 4 * AT_THREADS;

// dot products of the tk vectors vs[j][:] with one memory row, by one warp
__device__ __forceinline__ void warp_dots(const float* __restrict__ row, const float* __restrict__ vs, int tk,
                                          int H, int lane, float* out /*[AT_MAX_Q]*/) {
#pragma unroll
    for (int j = 0; j < AT_MAX_Q; ++j) out[j] = 0.f;
    for (int u = lane * 4; u < H; u += 128) {
        const float4 x = *reinterpret_cast<const float4*>(row + u);
#pragma unroll
        for (int j = 0; j < AT_MAX_Q; ++j)
            if (j < tk) {
                const float4 v = *reinterpret_cast<const float4*>(vs + (size_t)j * H + u);
                out[j] += x.x * v.x + x.y * v.y + x.z * v.z + x.w * v.w;
            }
    }
#pragma unroll
    for (int j = 0; j < AT_MAX_Q; ++j)
        if (j < tk) out[j] = warp_sum(out[j]);
}

__global__ void __launch_bounds__(AT_THREADS)
attn_train_fwd_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ keys,
                      const float* __restrict__ values, const int* __restrict__ mem_len, int B, int k, int tk,
                      int T, int H, float* __restrict__ alpha /*[B,k,tk,T]*/, float* __restrict__ part /*[B,k,tk,H]*/) {
    extern __shared__ float sm[];
    float* qs = sm;                          // [tk][H]
    float* sc = sm + (size_t)tk * H;         // [tk][AT_MAX_T]
    const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = B * k, r = b * k + i, nwarp = AT_THREADS / 32;
    for (int idx = tid; idx < tk * H; idx += AT_THREADS) {
        const int j = idx / H, u = idx - j * H;
        qs[idx] = q[((size_t)b * tk + j) * ldq + u];
    }
    int len = mem_len[r];
    len = len < 0 ? 0 : (len > T ? T : len);
    __syncthreads();
    for (int t = warp; t < len; t += nwarp) {
        float d[AT_MAX_Q];
        warp_dots(keys + ((size_t)t * R + r) * H, qs, tk, H, lane, d);
        if (lane == 0)
            for (int j = 0; j < tk; ++j) sc[j * AT_MAX_T + t] = d[j];
    }
    __syncthreads();
    for (int j = warp; j < tk; j += nwarp) {
        float m = -INFINITY;
        for (int t = lane; t < len; t += 32) m = fmaxf(m, sc[j * AT_MAX_T + t]);
        m = warp_max(m);
        float z = 0.f;
        for (int t = lane; t < len; t += 32) z += expf(sc[j * AT_MAX_T + t] - m);
        z = warp_sum(z);
        float* arow = alpha + (((size_t)b * k + i) * tk + j) * T;
        for (int t = lane; t < T; t += 32) {
            const float a = t < len ? expf(sc[j * AT_MAX_T + t] - m) / z : 0.f;
            if (t < len) sc[j * AT_MAX_T + t] = a;
            arow[t] = a;
        }
    }
    __syncthreads();
    const int u = tid * 4;
    if (u < H) {
        float acc[AT_MAX_Q][4];
#pragma unroll
        for (int j = 0; j < AT_MAX_Q; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        for (int t = 0; t < len; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(values + ((size_t)t * R + r) * H + u);
#pragma unroll
            for (int j = 0; j < AT_MAX_Q; ++j)
                if (j < tk) {
                    const float a = sc[j * AT_MAX_T + t];
                    acc[j][0] = fmaf(a, v.x, acc[j][0]); acc[j][1] = fmaf(a, v.y, acc[j][1]);
                    acc[j][2] = fmaf(a, v.z, acc[j][2]); acc[j][3] = fmaf(a, v.w, acc[j][3]);
                }
        }
#pragma unroll
        for (int j = 0; j < AT_MAX_Q; ++j)
            if (j < tk)
                *reinterpret_cast<float4*>(part + (((size_t)b * k + i) * tk + j) * H + u) =
                    make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
}

// out[(b*tk+j)*ldo + u] = (accumulate ? out : 0) + scale * sum_i part[b,i,j,u], i increasing
__global__ void attn_combine_kernel(const float* __restrict__ part, int B, int k, int tk, int H, float scale,
                                    float* __restrict__ out, int ldo, int accumulate) {
    const size_t n4 = (size_t)B * tk * (H / 4);
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= n4) return;
    const int u = (int)(idx % (H / 4)) * 4;
    const size_t bj = idx / (H / 4);
    const int j = (int)(bj % tk);
    const size_t b = bj / tk;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < k; ++i) {
        const float4 p = *reinterpret_cast<const float4*>(part + ((b * k + i) * tk + j) * H + u);
        a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    float4* o = reinterpret_cast<float4*>(out + bj * ldo + u);
    float4 base = accumulate ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
    *o = make_float4(base.x + a.x * scale, base.y + a.y * scale, base.z + a.z * scale, base.w + a.w * scale);
}

__global__ void __launch_bounds__(AT_THREADS)
attn_train_bwd_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ keys,
                      const float* __restrict__ values, const int* __restrict__ mem_len,
                      const float* __restrict__ alpha, const float* __restrict__ dctx, int ldd, int B, int k,
                      int tk, int T, int H, float* __restrict__ dkeys, float* __restrict__ dvalues,
                      float* __restrict__ dqpart /*[B,k,tk,H]*/) {
    extern __shared__ float sm[];
    float* qs = sm;                           // [tk][H]
    float* dcs = qs + (size_t)tk * H;         // [tk][H]   dctx / k
    float* al = dcs + (size_t)tk * H;         // [tk][AT_MAX_T]
    float* ds = al + (size_t)tk * AT_MAX_T;   // [tk][AT_MAX_T]  dalpha, then dscore
    const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = B * k, r = b * k + i, nwarp = AT_THREADS / 32;
    const float invk = 1.f / (float)k;
    for (int idx = tid; idx < tk * H; idx += AT_THREADS) {
        const int j = idx / H, u = idx - j * H;
        qs[idx] = q[((size_t)b * tk + j) * ldq + u];
        dcs[idx] = dctx[((size_t)b * tk + j) * ldd + u] * invk;
    }
    int len = mem_len[r];
    len = len < 0 ? 0 : (len > T ? T : len);
    for (int idx = tid; idx < tk * T; idx += AT_THREADS) {
        const int j = idx / T, t = idx - j * T;
        al[j * AT_MAX_T + t] = alpha[(((size_t)b * k + i) * tk + j) * T + t];
    }
    __syncthreads();
    for (int t = warp; t < len; t += nwarp) {
        float d[AT_MAX_Q];
        warp_dots(values + ((size_t)t * R + r) * H, dcs, tk, H, lane, d);
        if (lane == 0)
            for (int j = 0; j < tk; ++j) ds[j * AT_MAX_T + t] = d[j];
    }
    __syncthreads();
    for (int j = warp; j < tk; j += nwarp) {
        float dot = 0.f;
        for (int t = lane; t < len; t += 32) dot += al[j * AT_MAX_T + t] * ds[j * AT_MAX_T + t];
        dot = warp_sum(dot);
        for (int t = lane; t < len; t += 32)
            ds[j * AT_MAX_T + t] = al[j * AT_MAX_T + t] * (ds[j * AT_MAX_T + t] - dot);
    }
    __syncthreads();
    const int u = tid * 4;
    if (u < H) {
        float dq[AT_MAX_Q][4];
#pragma unroll
        for (int j = 0; j < AT_MAX_Q; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
        for (int t = 0; t < len; ++t) {
            const size_t off = ((size_t)t * R + r) * H + u;
            const float4 kx = *reinterpret_cast<const float4*>(keys + off);
            float4 dv = *reinterpret_cast<const float4*>(dvalues + off);
            float4 dk = *reinterpret_cast<const float4*>(dkeys + off);
#pragma unroll
            for (int j = 0; j < AT_MAX_Q; ++j)
                if (j < tk) {
                    const float a = al[j * AT_MAX_T + t], s = ds[j * AT_MAX_T + t];
                    const float4 dc = *reinterpret_cast<const float4*>(dcs + (size_t)j * H + u);
                    const float4 qv = *reinterpret_cast<const float4*>(qs + (size_t)j * H + u);
                    dv.x = fmaf(a, dc.x, dv.x); dv.y = fmaf(a, dc.y, dv.y); dv.z = fmaf(a, dc.z, dv.z); dv.w = fmaf(a, dc.w, dv.w);
                    dk.x = fmaf(s, qv.x, dk.x); dk.y = fmaf(s, qv.y, dk.y); dk.z = fmaf(s, qv.z, dk.z); dk.w = fmaf(s, qv.w, dk.w);
                    dq[j][0] = fmaf(s, kx.x, dq[j][0]); dq[j][1] = fmaf(s, kx.y, dq[j][1]);
                    dq[j][2] = fmaf(s, kx.z, dq[j][2]); dq[j][3] = fmaf(s, kx.w, dq[j][3]);
                }
            *reinterpret_cast<float4*>(dvalues + off) = dv;
            *reinterpret_cast<float4*>(dkeys + off) = dk;
        }
#pragma unroll
        for (int j = 0; j < AT_MAX_Q; ++j)
            if (j < tk)
                *reinterpret_cast<float4*>(dqpart + (((size_t)b * k + i) * tk + j) * H + u) =
                    make_float4(dq[j][0], dq[j][1], dq[j][2], dq[j][3]);
    }
}

// dst[row, :F1] = src[row, c0 : c0 + F1]  (src rows of width F)
__global__ void split_cols_kernel(const float* __restrict__ src, int F, int c0, int F1, size_t rows,
                                  float* __restrict__ dst) {
    const size_t total = rows * F1;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx / F1;
        const int c = (int)(idx % F1);
        dst[idx] = src[row * F + c0 + c];
    }
}

int check_dims(int B, int k, int tk, int T, int H, int ld1, int ld2) {
    D2P_REQUIRE(B > 0 && k > 0 && tk >= 1 && tk <= AT_MAX_Q && T >= 1 && T <= AT_MAX_T && H >= 4 && H % 4 == 0 &&
                H <= AT_MAX_H, "luong attention (train): unsupported dims (test_k=%d T=%d H=%d)", tk, T, H);
    D2P_REQUIRE(ld1 >= H && ld2 >= H && ld1 % 4 == 0 && ld2 % 4 == 0, "luong attention (train): bad row strides");
    return 0;
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" size_t d2p_luong_pool_attention_train_ws_bytes(int B, int k, int tk, int H) {
    return (size_t)B * k * tk * H * sizeof(float);
}

extern "C" int d2p_luong_pool_attention_train_fwd(const float* q, int ldq, const float* keys, const float* values,
                                                  const int* mem_len, int B, int k, int tk, int T, int H,
                                                  float* ctx, int ldc, float* alpha, void* ws, size_t ws_bytes,
                                                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(q && keys && values && mem_len && ctx && alpha && ws, "luong attention train fwd: null buffer");
    D2P_TRY(check_dims(B, k, tk, T, H, ldq, ldc));
    D2P_REQUIRE(ws_bytes >= d2p_luong_pool_attention_train_ws_bytes(B, k, tk, H),
                "luong attention train fwd: workspace too small");
    D2P_REQUIRE((((uintptr_t)q | (uintptr_t)keys | (uintptr_t)values | (uintptr_t)ctx | (uintptr_t)ws) & 15) == 0,
                "luong attention train fwd: buffers must be 16-byte aligned");
    const size_t smem = ((size_t)tk * H + (size_t)tk * AT_MAX_T) * sizeof(float);
    float* part = (float*)ws;
    attn_train_fwd_kernel<<<dim3(k, B), AT_THREADS, smem, st>>>(q, ldq, keys, values, mem_len, B, k, tk, T, H, alpha,
                                                                 part);
    D2P_CHECK_LAUNCH();
    const size_t n4 = (size_t)B * tk * (H / 4);
    attn_combine_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(part, B, k, tk, H, 1.f / (float)k, ctx, ldc, 0);
    D2P_CHECK_LAUNCH();
    return 0;
}

/* dq rows (b*tk+j, stride lddq) are ACCUMULATED into; dkeys / dvalues [T,R,H] are accumulated into. */
extern "C" int d2p_luong_pool_attention_train_bwd(const float* q, int ldq, const float* keys, const float* values,
                                                  const int* mem_len, const float* alpha, const float* dctx,
                                                  int ldd, int B, int k, int tk, int T, int H, float* dq, int lddq,
                                                  float* dkeys, float* dvalues, void* ws, size_t ws_bytes,
                                                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(q && keys && values && mem_len && alpha && dctx && dq && dkeys && dvalues && ws,
                "luong attention train bwd: null buffer");
    D2P_TRY(check_dims(B, k, tk, T, H, ldq, ldd));
    D2P_REQUIRE(lddq >= H && lddq % 4 == 0, "luong attention train bwd: bad dq stride");
    D2P_REQUIRE(ws_bytes >= d2p_luong_pool_attention_train_ws_bytes(B, k, tk, H),
                "luong attention train bwd: workspace too small");
    D2P_REQUIRE((((uintptr_t)q | (uintptr_t)keys | (uintptr_t)values | (uintptr_t)dctx | (uintptr_t)dq |
                  (uintptr_t)dkeys | (uintptr_t)dvalues | (uintptr_t)ws) & 15) == 0,
                "luong attention train bwd: buffers must be 16-byte aligned");
    const size_t smem = (2 * (size_t)tk * H + 2 * (size_t)tk * AT_MAX_T) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(attn_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            64 * 1024));
        attr = true;
    }
    D2P_REQUIRE(smem <= 64 * 1024, "luong attention train bwd: shared memory budget exceeded");
    float* part = (float*)ws;
    attn_train_bwd_kernel<<<dim3(k, B), AT_THREADS, smem, st>>>(q, ldq, keys, values, mem_len, alpha, dctx, ldd, B, k,
                                                                 tk, T, H, dkeys, dvalues, part);
    D2P_CHECK_LAUNCH();
    const size_t n4 = (size_t)B * tk * (H / 4);
    attn_combine_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(part, B, k, tk, H, 1.f, dq, lddq, 1);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_split_cols(const float* src, int F, int c0, int F1, long long rows, float* dst, void* stream) {
    D2P_REQUIRE(src && dst && rows > 0 && c0 >= 0 && F1 > 0 && c0 + F1 <= F, "split_cols: bad arguments");
    const size_t total = (size_t)rows * F1;
    const size_t nb = (total + 255) / 256, cap = 8 * (size_t)kNumSMs;
    split_cols_kernel<<<(int)(nb < cap ? nb : cap), 256, 0, (cudaStream_t)stream>>>(src, F, c0, F1, (size_t)rows, dst);
    D2P_CHECK_LAUNCH();
    return 0;
}
