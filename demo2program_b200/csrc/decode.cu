// Greedy decoding: dynamic_decode(BasicDecoder(cell, GreedyEmbeddingHelper(embed,
// start_tokens, end_token), (c, h), Dense(V, no bias)), maximum_iterations = L)
// with impute_finished = False (reference models/model_full.py:424-435, 513-521;
// SURVEY A.6):
//   * first input = Emb[start_id] (start id = token_dim, an in-range row);
//   * sample = argmax(logits), lowest index wins ties; next input = Emb[sample];
//   * a row finishes when it samples end_token; finished rows KEEP stepping on their
//     own samples (outputs are not zeroed) until every row is finished or L steps ran;
//   * length = first finishing step + 1, or L; logits past the executed steps are 0
//     (the model zero-pads to L, model_full.py:476-484).
// The loop runs on the device with no host synchronisation: an "all finished" flag
// written by step t masks the work of steps > t (their logits are zero).
#include "common.cuh"
#include <climits>

namespace d2p {
namespace {

// X rows have stride ldx (>= E)
__global__ void gather_rows_kernel(const float* __restrict__ table, int vocab_rows, int E,
                                   const int* __restrict__ ids, int R, float* __restrict__ X, int ldx) {
    size_t total = (size_t)R * E;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / E), e = (int)(idx % E);
        int id = ids[r];
        X[(size_t)r * ldx + e] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E + e] : 0.f;
    }
}

// in-place BasicLSTMCell on pre-activations G [R,4H]; (c, h) updated in place; h rows have stride ldh
__global__ void lstm_cell_infer_kernel(const float* __restrict__ G, float* __restrict__ c,
                                       float* __restrict__ h, int ldh, int R, int H, float forget_bias) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * H) return;
    int r = idx / H, u = idx % H;
    const float* g = G + (size_t)r * 4 * H;
    float i = sigmoid_f(g[u]), j = tanhf(g[H + u]);
    float f = sigmoid_f(g[2 * H + u] + forget_bias), o = sigmoid_f(g[3 * H + u]);
    float cn = c[idx] * f + i * j;
    c[idx] = cn;
    h[(size_t)r * ldh + u] = tanhf(cn) * o;
}

// one warp per row: argmax (lowest index on ties), finished / length bookkeeping
__global__ void greedy_update_kernel(float* __restrict__ logits_t, int R, int V, int t, int max_len,
                                     int end_id, int* __restrict__ ids, int* __restrict__ finished,
                                     int* __restrict__ lengths, int* __restrict__ tokens_t,
                                     const int* __restrict__ all_done, int nsl) {
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (row >= R) return;
    float* x = logits_t + (size_t)row * V;
    if (all_done[row % nsl]) {   // this decoder instance's loop has ended: zero padding
        for (int v = lane; v < V; v += 32) x[v] = 0.f;
        if (lane == 0) tokens_t[row] = 0;
        return;
    }
    float best = -INFINITY; int bi = INT_MAX;
    for (int v = lane; v < V; v += 32) {
        float xv = x[v];
        if (xv > best || (xv == best && v < bi)) { best = xv; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        ids[row] = bi;
        tokens_t[row] = bi;
        int was = finished[row];
        int now = was || (bi == end_id) || (t + 1 >= max_len);
        if (!was && now) lengths[row] = t + 1;
        finished[row] = now;
    }
}

// rows r of the same decoder instance share r % nsl (the reference builds one
// dynamic_decode loop per demonstration index); one block per instance
__global__ void all_done_kernel(const int* __restrict__ finished, int R, int nsl,
                                int* __restrict__ all_done) {
    __shared__ int any_live;
    const int g = blockIdx.x;
    if (threadIdx.x == 0) any_live = 0;
    __syncthreads();
    for (int r = g + threadIdx.x * nsl; r < R; r += blockDim.x * nsl)
        if (!finished[r]) any_live = 1;   // benign race: all writers store 1
    __syncthreads();
    if (threadIdx.x == 0 && !any_live) all_done[g] = 1;
}

__global__ void greedy_init_kernel(int* ids, int* finished, int* lengths, int* all_done, int R,
                                   int nsl, int start_id) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) { ids[r] = start_id; finished[r] = 0; lengths[r] = 0; }
    if (r < nsl) all_done[r] = 0;
}

__global__ void copy_k(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---- K6: pooled Luong attention (induction baseline) ---------------------------------
// reference models/baselines/model_induction.py:25-53 (_compute_attention), 107-182
// (PoolingAttentionWrapper.call), 638-667; SURVEY A.11.  For query row (b, j) and every
// seen demo i: score = h . keys[b,i,t',:], positions >= len[b,i] masked to -inf, softmax,
// context_i = alpha . values[b,i]; the k attention vectors [h; context_i] W_a are averaged.
// W_a is shared and linear, so mean_i([h;ctx_i] W_a) = h W_a[:H] + (mean_i ctx_i) W_a[H:]:
// these kernels emit mean_i ctx_i.
// One CTA per (demo i, batch element b) serves all test_k queries, so every key/value row
// is read from HBM exactly once per decode step and B*k CTAs keep enough bytes in flight;
// the per-demo contexts go to a [B,k,tk,H] scratch and a second kernel averages over i in
// a fixed order (deterministic).
constexpr int ATT_MAX_Q = 8, ATT_MAX_T = 64, ATT_THREADS = 256;

__global__ void __launch_bounds__(ATT_THREADS)
luong_attn_partial_kernel(const float* __restrict__ q /*rows b*tk+j, stride ldq*/, int ldq,
                          const float* __restrict__ keys, const float* __restrict__ values /*[T,R,H]*/,
                          const int* __restrict__ mem_len, int B, int k, int tk, int T, int H,
                          float* __restrict__ part /*[B,k,tk,H]*/) {
    extern __shared__ float sm[];
    float* qs = sm;                         // [tk][H]
    float* sc = sm + (size_t)tk * H;        // [tk][ATT_MAX_T]
    const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = B * k, nwarp = ATT_THREADS / 32, r = b * k + i;
    for (int idx = tid; idx < tk * H; idx += ATT_THREADS) {
        const int j = idx / H, u = idx - j * H;
        qs[idx] = q[((size_t)b * tk + j) * ldq + u];
    }
    int len = mem_len[r];
    len = len < 0 ? 0 : (len > T ? T : len);
    __syncthreads();
    // scores: one warp per memory position, the row's loads issued before any use
    for (int t = warp; t < len; t += nwarp) {
        const float* krow = keys + ((size_t)t * R + r) * H;
        float part_s[ATT_MAX_Q];
#pragma unroll
        for (int j = 0; j < ATT_MAX_Q; ++j) part_s[j] = 0.f;
        for (int u0 = 0; u0 < H; u0 += 512) {
            float4 kv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = u0 + e * 128 + lane * 4;
                kv[e] = u < H ? __ldcs(reinterpret_cast<const float4*>(krow + u)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = u0 + e * 128 + lane * 4;
                if (u < H) {
#pragma unroll
                    for (int j = 0; j < ATT_MAX_Q; ++j)
                        if (j < tk) {
                            const float4 qv = *reinterpret_cast<const float4*>(qs + (size_t)j * H + u);
                            part_s[j] += kv[e].x * qv.x + kv[e].y * qv.y + kv[e].z * qv.z + kv[e].w * qv.w;
                        }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < ATT_MAX_Q; ++j)
            if (j < tk) {
                float sv = warp_sum(part_s[j]);
                if (lane == 0) sc[j * ATT_MAX_T + t] = sv;
            }
    }
    __syncthreads();
    // masked softmax over t' < len: warp j normalises query j
    for (int j = warp; j < tk; j += nwarp) {
        float m = -INFINITY;
        for (int t = lane; t < len; t += 32) m = fmaxf(m, sc[j * ATT_MAX_T + t]);
        m = warp_max(m);
        float z = 0.f;
        for (int t = lane; t < len; t += 32) z += expf(sc[j * ATT_MAX_T + t] - m);
        z = warp_sum(z);
        for (int t = lane; t < len; t += 32)
            sc[j * ATT_MAX_T + t] = expf(sc[j * ATT_MAX_T + t] - m) / z;
    }
    __syncthreads();
    // context: a thread owns 4 consecutive hidden units; when H/4 < ATT_THREADS the thread
    // groups interleave the memory positions and are combined through shared memory.
    // Value rows are independent loads, 4 positions in flight per thread.
    const int q4 = H / 4, ng = ATT_THREADS / q4 > 0 ? ATT_THREADS / q4 : 1;
    const int grp = tid / q4, u = (tid - grp * q4) * 4;
    float* red = sc + (size_t)tk * ATT_MAX_T;     // [ng][tk][H]
    float acc[ATT_MAX_Q][4];
#pragma unroll
    for (int j = 0; j < ATT_MAX_Q; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    if (grp < ng) {
        const float* vbase = values + (size_t)r * H + u;
        const size_t tstride = (size_t)R * H;
        int t = grp;
        for (; t + 3 * ng < len; t += 4 * ng) {
            float4 v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                v[e] = __ldcs(reinterpret_cast<const float4*>(vbase + (size_t)(t + e * ng) * tstride));
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int j = 0; j < ATT_MAX_Q; ++j)
                    if (j < tk) {
                        const float a = sc[j * ATT_MAX_T + t + e * ng];
                        acc[j][0] = fmaf(a, v[e].x, acc[j][0]); acc[j][1] = fmaf(a, v[e].y, acc[j][1]);
                        acc[j][2] = fmaf(a, v[e].z, acc[j][2]); acc[j][3] = fmaf(a, v[e].w, acc[j][3]);
                    }
        }
        for (; t < len; t += ng) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(vbase + (size_t)t * tstride));
#pragma unroll
            for (int j = 0; j < ATT_MAX_Q; ++j)
                if (j < tk) {
                    const float a = sc[j * ATT_MAX_T + t];
                    acc[j][0] = fmaf(a, v.x, acc[j][0]); acc[j][1] = fmaf(a, v.y, acc[j][1]);
                    acc[j][2] = fmaf(a, v.z, acc[j][2]); acc[j][3] = fmaf(a, v.w, acc[j][3]);
                }
        }
        if (grp > 0) {
#pragma unroll
            for (int j = 0; j < ATT_MAX_Q; ++j)
                if (j < tk)
                    *reinterpret_cast<float4*>(red + ((size_t)grp * tk + j) * H + u) =
                        make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        }
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
        for (int j = 0; j < ATT_MAX_Q; ++j)
            if (j < tk) {
                float4 a = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                for (int g2 = 1; g2 < ng; ++g2) {
                    const float4 p = *reinterpret_cast<const float4*>(red + ((size_t)g2 * tk + j) * H + u);
                    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
                }
                *reinterpret_cast<float4*>(part + (((size_t)b * k + i) * tk + j) * H + u) = a;
            }
    }
}

// ctx[(b*tk+j)*ldc + u] = (1/k) sum_i part[b,i,j,u], i in increasing order
__global__ void luong_attn_mean_kernel(const float* __restrict__ part, int B, int k, int tk, int H,
                                       float* __restrict__ ctx, int ldc) {
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
    const float inv = 1.f / (float)k;
    *reinterpret_cast<float4*>(ctx + bj * ldc + u) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}


// Arg-max margin guard of the tensor-core greedy path: counts the (step, row) positions that a
// row actually executed (t < lengths[row]) whose top-2 logit gap is <= rel_tol * max(1, max|logit|),
// i.e. where the ~5e-6 relative error of the bf16x3 products could flip the arg-max against an
// fp32 evaluation.  The host re-runs the decode on the exact fp32 engine when the count is not 0.
__global__ void near_tie_kernel(const float* __restrict__ logits, int Tdec, int R, int V,
                                const int* __restrict__ lengths, float rel_tol, int* __restrict__ count) {
    const int warps = blockDim.x / 32, lane = threadIdx.x % 32;
    const long long total = (long long)Tdec * R;
    for (long long idx = blockIdx.x * (long long)warps + threadIdx.x / 32; idx < total;
         idx += (long long)gridDim.x * warps) {
        const int t = (int)(idx / R), row = (int)(idx % R);
        if (t >= lengths[row]) continue;
        const float* x = logits + (size_t)idx * V;
        float b1 = -INFINITY, b2 = -INFINITY, am = 0.f;
        for (int v = lane; v < V; v += 32) {
            const float xv = x[v];
            am = fmaxf(am, fabsf(xv));
            if (xv > b1) { b2 = b1; b1 = xv; } else if (xv > b2) b2 = xv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float o1 = __shfl_xor_sync(0xffffffffu, b1, o), o2 = __shfl_xor_sync(0xffffffffu, b2, o);
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
            if (o1 > b1) { b2 = fmaxf(b1, o2); b1 = o1; } else b2 = fmaxf(b2, o1);
        }
        if (lane == 0 && V > 1 && b1 - b2 <= rel_tol * fmaxf(1.f, am)) atomicAdd(count, 1);
    }
}

// x[r2,:] = table[id] with id = (t == 0 ? start_id : tokens[r2, t-1]); out of range -> 0
__global__ void gather_step_kernel(const float* __restrict__ table, int vocab_rows, int E,
                                   const int* __restrict__ tokens, int R2, int L, int t, int start_id,
                                   float* __restrict__ X, int ldx) {
    size_t total = (size_t)R2 * E;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / E), e = (int)(idx % E);
        int id = t == 0 ? start_id : tokens[(size_t)r * L + t - 1];
        X[(size_t)r * ldx + e] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E + e] : 0.f;
    }
}

// out[(b*tk+j)*ldo + u] = S ? S[b,u] : 0
__global__ void bcast_rows_kernel(const float* __restrict__ S, int B, int tk, int H,
                                  float* __restrict__ out, int ldo) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * tk * H) return;
    const size_t row = idx / H;
    out[row * ldo + idx % H] = S ? S[(row / tk) * H + idx % H] : 0.f;
}

// out[row, :] = [A[row, :F1] ; Bm[row, :F2]]
__global__ void concat_cols_kernel(const float* __restrict__ A, int F1, const float* __restrict__ Bm,
                                   int F2, size_t rows, float* __restrict__ out) {
    const int F = F1 + F2;
    size_t total = rows * F;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t row = idx / F; int c = (int)(idx % F);
        out[idx] = c < F1 ? A[row * F1 + c] : Bm[row * F2 + (c - F1)];
    }
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" int d2p_concat_cols(const float* A, int F1, const float* Bm, int F2, long long rows,
                               float* out, void* stream) {
    D2P_REQUIRE(A && Bm && out && rows > 0, "concat_cols: bad arguments");
    size_t total = (size_t)rows * (F1 + F2);
    size_t b = (total + 255) / 256, cap = 8 * (size_t)kNumSMs;
    concat_cols_kernel<<<(int)(b < cap ? b : cap), 256, 0, (cudaStream_t)stream>>>(A, F1, Bm, F2, (size_t)rows, out);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t d2p_luong_pool_attention_ws_bytes(int B, int k, int tk, int H) {
    return (size_t)B * k * tk * H * sizeof(float);   // per-demo contexts [B,k,tk,H]
}

// q rows (b*tk + j) have stride ldq, ctx rows stride ldc (both >= H, multiples of 4 floats)
extern "C" int d2p_luong_pool_attention(const float* q, int ldq, const float* keys, const float* values,
                                        const int* mem_len, int B, int k, int tk, int T, int H,
                                        float* ctx, int ldc, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(q && keys && values && mem_len && ctx && ws, "luong attention: null buffer");
    D2P_REQUIRE(B > 0 && k > 0 && tk >= 1 && tk <= ATT_MAX_Q && T >= 1 && T <= ATT_MAX_T && H >= 4 &&
                H % 4 == 0 && H <= 4 * ATT_THREADS,
                "luong attention: unsupported dims (test_k=%d T=%d H=%d)", tk, T, H);
    D2P_REQUIRE(ldq >= H && ldc >= H && ldq % 4 == 0 && ldc % 4 == 0, "luong attention: bad row strides");
    D2P_REQUIRE(ws_bytes >= d2p_luong_pool_attention_ws_bytes(B, k, tk, H), "luong attention: workspace too small");
    D2P_REQUIRE((((uintptr_t)q | (uintptr_t)keys | (uintptr_t)values | (uintptr_t)ctx | (uintptr_t)ws) & 15) == 0,
                "luong attention: buffers must be 16-byte aligned");
    const int q4 = H / 4, ng = ATT_THREADS / q4 > 0 ? ATT_THREADS / q4 : 1;
    const size_t smem = ((size_t)tk * H + (size_t)tk * ATT_MAX_T + (ng > 1 ? (size_t)ng * tk * H : 0)) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(luong_attn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            100 * 1024));
        attr = true;
    }
    D2P_REQUIRE(smem <= 100 * 1024, "luong attention: shared memory budget exceeded");
    float* part = (float*)ws;
    luong_attn_partial_kernel<<<dim3(k, B), ATT_THREADS, smem, st>>>(q, ldq, keys, values, mem_len, B, k, tk, T,
                                                                     H, part);
    D2P_CHECK_LAUNCH();
    const size_t n4 = (size_t)B * tk * (H / 4);
    luong_attn_mean_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(part, B, k, tk, H, ctx, ldc);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t d2p_induction_decode_ws_bytes(int B, int k, int tk, int H) {
    size_t R2 = (size_t)B * tk;
    // [x|att|h|ctx] [R2,4H] | c [R2,H] | gates [R2,4H] | per-demo contexts | ids, finished [R2] | all_done[64]
    // (+ [R2,H] for the folded memory-layer query h * W_mem^T)
    return (R2 * 10 * H) * sizeof(float) + d2p_luong_pool_attention_ws_bytes(B, k, tk, H) +
           (2 * R2 + 64) * sizeof(int) + 256;
}

// Attention-LSTM action decoder of the induction baseline over R2 = B*test_k rows
// (row = b*test_k + j).  tokens != NULL: teacher forcing (TrainingHelper, <s> id out of
// range -> zero row); tokens == NULL: greedy (start id = A, end id = A-1).
// Initial cell state reproduces the reference's swap (model_induction.py:674-676):
// c := demo_h_summary, h := demo_c_summary.  logits [Tdec, R2, A] time-major.
// The step state lives in one [R2, 4H] buffer  x | attention | h | context  so that the
// cell's input contraction [x; attention_{t-1}; h_{t-1}] * W (K = 3H) and the attention
// layer [h; context] * W_a (K = 2H) are each ONE product over adjacent columns.
// keys == NULL with memory_layer [H,H] given: the LuongAttention memory layer is folded into the
// query - score_t = h . (values_t W_mem) = (W_mem h) . values_t - so the keys tensor is never
// materialised (identical up to fp32 summation order).  Measured at C5: the encoder saves the
// 54 GFLOP key product (-0.3 ms) but every decode step pays a [B*test_k, H, H] product and the
// attention kernel is latency-, not byte-bound (+0.4 ms over 20 steps): off by default.
extern "C" int d2p_induction_decode(const float* keys, const float* memory_layer, const float* values,
                                    const int* mem_len, int B,
                                    int k, int tk, int T, int H, const float* h_sum,
                                    const float* c_sum, const float* table, int A, const float* Wcell,
                                    const float* bcell, const float* Wa, const float* proj,
                                    const int* tokens, int Tdec, float* logits, int* out_tokens,
                                    int* lengths, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE((keys || memory_layer) && values && mem_len && h_sum && c_sum && table && Wcell && bcell && Wa &&
                proj && logits && ws, "induction decode: null buffer");
    D2P_REQUIRE(ws_bytes >= d2p_induction_decode_ws_bytes(B, k, tk, H), "induction decode: workspace too small");
    D2P_REQUIRE(H % 4 == 0, "induction decode: H must be a multiple of 4");
    const bool greedy = tokens == nullptr;
    if (greedy) D2P_REQUIRE(out_tokens && lengths, "induction decode: greedy needs token/length outputs");
    const int R2 = B * tk, G4 = 4 * H, LD = 4 * H;
    const size_t RH = (size_t)R2 * H;
    float* xs = (float*)ws;                       // [R2, 4H]
    float* x = xs; float* att = xs + H; float* h = xs + 2 * H; float* ctx = xs + 3 * H;
    float* c = xs + (size_t)R2 * LD;
    float* gates = c + RH;
    float* part = gates + (size_t)R2 * G4;
    const size_t part_bytes = d2p_luong_pool_attention_ws_bytes(B, k, tk, H);
    float* qp = (float*)((char*)part + part_bytes);      // [R2, H] folded query
    int* ids = (int*)(qp + RH);
    int* finished = ids + R2; int* all_done = finished + R2;
    const int eb = cdiv((long long)RH, 256);
    bcast_rows_kernel<<<eb, 256, 0, st>>>(h_sum, B, tk, H, c, H);   // swapped on purpose (F8)
    D2P_CHECK_LAUNCH();
    bcast_rows_kernel<<<eb, 256, 0, st>>>(c_sum, B, tk, H, h, LD);
    D2P_CHECK_LAUNCH();
    bcast_rows_kernel<<<eb, 256, 0, st>>>(nullptr, B, tk, H, att, LD);   // attention_0 = 0
    D2P_CHECK_LAUNCH();
    if (greedy) {
        greedy_init_kernel<<<cdiv(R2, 256), 256, 0, st>>>(ids, finished, lengths, all_done, R2, tk, A);
        D2P_CHECK_LAUNCH();
    }
    for (int t = 0; t < Tdec; ++t) {
        if (greedy) gather_rows_kernel<<<eb, 256, 0, st>>>(table, A + 1, H, ids, R2, x, LD);
        else gather_step_kernel<<<eb, 256, 0, st>>>(table, A + 1, H, tokens, R2, Tdec, t, A + 1, x, LD);
        D2P_CHECK_LAUNCH();
        // kernel rows [0,H) take the embedding, [H,2H) attention_{t-1}, [2H,3H) h_{t-1}
        D2P_TRY(gemm(st, false, false, R2, G4, 3 * H, 1.f, xs, LD, Wcell, G4, 0.f, gates, G4, bcell, GEMM_CONST_B));
        lstm_cell_infer_kernel<<<eb, 256, 0, st>>>(gates, c, h, LD, R2, H, 1.0f);
        D2P_CHECK_LAUNCH();
        if (keys) {
            D2P_TRY(d2p_luong_pool_attention(h, LD, keys, values, mem_len, B, k, tk, T, H, ctx, LD, part,
                                             part_bytes, stream));
        } else {   // q' = W_mem h, scored against the values themselves
            D2P_TRY(gemm(st, false, true, R2, H, H, 1.f, h, LD, memory_layer, H, 0.f, qp, H, nullptr, GEMM_CONST_B));
            D2P_TRY(d2p_luong_pool_attention(qp, H, values, values, mem_len, B, k, tk, T, H, ctx, LD, part,
                                             part_bytes, stream));
        }
        // attention_t = [h_t; mean context] * W_a  (reads columns [2H,4H), writes [H,2H))
        D2P_TRY(gemm(st, false, false, R2, H, 2 * H, 1.f, h, LD, Wa, H, 0.f, att, LD, nullptr, GEMM_CONST_B));
        float* lt = logits + (size_t)t * R2 * A;
        D2P_TRY(gemm(st, false, false, R2, A, H, 1.f, att, LD, proj, A, 0.f, lt, A, nullptr, GEMM_CONST_B));
        if (greedy) {
            greedy_update_kernel<<<cdiv(R2, 8), 256, 0, st>>>(lt, R2, A, t, Tdec, A - 1, ids, finished, lengths,
                                                              out_tokens + (size_t)t * R2, all_done, tk);
            D2P_CHECK_LAUNCH();
            all_done_kernel<<<tk, 256, 0, st>>>(finished, R2, tk, all_done);
            D2P_CHECK_LAUNCH();
        }
    }
    return 0;
}

extern "C" int d2p_greedy_near_ties(const float* logits, int Tdec, int R, int V, const int* lengths,
                                    float rel_tol, int* count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(logits && lengths && count && Tdec > 0 && R > 0 && V > 0, "greedy_near_ties: bad arguments");
    D2P_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
    const long long total = (long long)Tdec * R;
    long long blocks = (total + 7) / 8;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    near_tie_kernel<<<(int)blocks, 256, 0, st>>>(logits, Tdec, R, V, lengths, rel_tol, count);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t d2p_greedy_ws_bytes(int R, int H, int E) {
    // x [R,E] | gates [R,4H] | h [R,H] | c [R,H] | ids, finished [R] | all_done
    return ((size_t)R * (E + 6 * H) + 16) * sizeof(float) + (size_t)(2 * R + 64) * sizeof(int) + 256;
}

extern "C" int d2p_lstm_decoder_greedy(const float* table, int vocab_rows, int E, const float* W,
                                       const float* b, const float* proj, int R, int H, int V,
                                       int start_id, int end_id, int max_len, int nsl, const float* h0,
                                       const float* c0, float* logits, int* tokens, int* lengths,
                                       void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(table && W && b && proj && h0 && c0 && logits && tokens && lengths && ws,
                "greedy: null buffer");
    D2P_REQUIRE(ws_bytes >= d2p_greedy_ws_bytes(R, H, E), "greedy: workspace too small");
    D2P_REQUIRE(nsl >= 1 && nsl <= 64 && R % nsl == 0, "greedy: bad instance count %d", nsl);
    float* x = (float*)ws;
    float* gates = x + (size_t)R * E;
    float* h = gates + (size_t)R * 4 * H;
    float* c = h + (size_t)R * H;
    int* ids = (int*)(c + (size_t)R * H);
    int* finished = ids + R;
    int* all_done = finished + R;
    const int G4 = 4 * H;
    const float* Wx = W;
    const float* Wh = W + (size_t)E * G4;
    greedy_init_kernel<<<cdiv(R, 256), 256, 0, st>>>(ids, finished, lengths, all_done, R, nsl, start_id);
    D2P_CHECK_LAUNCH();
    copy_k<<<cdiv((long long)R * H, 256), 256, 0, st>>>(h, h0, (size_t)R * H);
    D2P_CHECK_LAUNCH();
    copy_k<<<cdiv((long long)R * H, 256), 256, 0, st>>>(c, c0, (size_t)R * H);
    D2P_CHECK_LAUNCH();
    for (int t = 0; t < max_len; ++t) {
        size_t tot = (size_t)R * E;
        gather_rows_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(table, vocab_rows, E, ids, R, x, E);
        D2P_CHECK_LAUNCH();
        D2P_TRY(gemm(st, false, false, R, G4, E, 1.f, x, E, Wx, G4, 0.f, gates, G4, b, GEMM_CONST_B));
        D2P_TRY(gemm(st, false, false, R, G4, H, 1.f, h, H, Wh, G4, 1.f, gates, G4, nullptr, GEMM_CONST_B));
        lstm_cell_infer_kernel<<<cdiv((long long)R * H, 256), 256, 0, st>>>(gates, c, h, H, R, H, 1.0f);
        D2P_CHECK_LAUNCH();
        float* lt = logits + (size_t)t * R * V;
        D2P_TRY(gemm(st, false, false, R, V, H, 1.f, h, H, proj, V, 0.f, lt, V, nullptr, GEMM_CONST_B));
        greedy_update_kernel<<<cdiv(R, 8), 256, 0, st>>>(lt, R, V, t, max_len, end_id, ids, finished,
                                                         lengths, tokens + (size_t)t * R, all_done, nsl);
        D2P_CHECK_LAUNCH();
        all_done_kernel<<<nsl, 256, 0, st>>>(finished, R, nsl, all_done);
        D2P_CHECK_LAUNCH();
    }
    return 0;
}
