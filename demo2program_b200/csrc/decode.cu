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

__global__ void gather_rows_kernel(const float* __restrict__ table, int vocab_rows, int E,
                                   const int* __restrict__ ids, int R, float* __restrict__ X) {
    size_t total = (size_t)R * E;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / E), e = (int)(idx % E);
        int id = ids[r];
        X[idx] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E + e] : 0.f;
    }
}

// in-place BasicLSTMCell on pre-activations G [R,4H]; (c, h) updated in place
__global__ void lstm_cell_infer_kernel(const float* __restrict__ G, float* __restrict__ c,
                                       float* __restrict__ h, int R, int H, float forget_bias) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * H) return;
    int r = idx / H, u = idx % H;
    const float* g = G + (size_t)r * 4 * H;
    float i = sigmoid_f(g[u]), j = tanhf(g[H + u]);
    float f = sigmoid_f(g[2 * H + u] + forget_bias), o = sigmoid_f(g[3 * H + u]);
    float cn = c[idx] * f + i * j;
    c[idx] = cn;
    h[idx] = tanhf(cn) * o;
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
// this kernel emits mean_i ctx_i.  One CTA per batch element serves all test_k queries,
// so every key/value row is read from HBM exactly once per decode step.
constexpr int ATT_MAX_Q = 8, ATT_MAX_T = 64, ATT_THREADS = 256;

__global__ void __launch_bounds__(ATT_THREADS)
luong_pool_attn_kernel(const float* __restrict__ q /*[B*tk,H]*/, const float* __restrict__ keys,
                       const float* __restrict__ values /*[T,R,H]*/, const int* __restrict__ mem_len,
                       int B, int k, int tk, int T, int H, float* __restrict__ ctx /*[B*tk,H]*/) {
    extern __shared__ float sm[];
    float* qs = sm;                         // [tk][H]
    float* sc = sm + (size_t)tk * H;        // [tk][ATT_MAX_T]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = B * k, nwarp = ATT_THREADS / 32;
    for (int idx = tid; idx < tk * H; idx += ATT_THREADS) qs[idx] = q[(size_t)b * tk * H + idx];
    float acc[ATT_MAX_Q][4];
#pragma unroll
    for (int j = 0; j < ATT_MAX_Q; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    __syncthreads();
    for (int i = 0; i < k; ++i) {
        const int r = b * k + i;
        int len = mem_len[r];
        len = len < 0 ? 0 : (len > T ? T : len);
        // scores: one warp per memory position
        for (int t = warp; t < len; t += nwarp) {
            const float* krow = keys + ((size_t)t * R + r) * H;
            float part[ATT_MAX_Q];
#pragma unroll
            for (int j = 0; j < ATT_MAX_Q; ++j) part[j] = 0.f;
            for (int u = lane * 4; u < H; u += 128) {
                float4 kv = *reinterpret_cast<const float4*>(krow + u);
#pragma unroll
                for (int j = 0; j < ATT_MAX_Q; ++j)
                    if (j < tk) {
                        const float* qj = qs + (size_t)j * H + u;
                        part[j] += kv.x * qj[0] + kv.y * qj[1] + kv.z * qj[2] + kv.w * qj[3];
                    }
            }
#pragma unroll
            for (int j = 0; j < ATT_MAX_Q; ++j)
                if (j < tk) {
                    float s = warp_sum(part[j]);
                    if (lane == 0) sc[j * ATT_MAX_T + t] = s;
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
        // context: each thread owns hidden units tid, tid+256, ...; value rows read once
        for (int t = 0; t < len; ++t) {
            const float* vrow = values + ((size_t)t * R + r) * H;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = tid + e * ATT_THREADS;
                if (u < H) {
                    const float v = vrow[u];
#pragma unroll
                    for (int j = 0; j < ATT_MAX_Q; ++j)
                        if (j < tk) acc[j][e] = fmaf(sc[j * ATT_MAX_T + t], v, acc[j][e]);
                }
            }
        }
        __syncthreads();
    }
    const float inv = 1.f / (float)k;
#pragma unroll
    for (int j = 0; j < ATT_MAX_Q; ++j)
        if (j < tk)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = tid + e * ATT_THREADS;
                if (u < H) ctx[((size_t)b * tk + j) * H + u] = acc[j][e] * inv;
            }
}

// x[r2,:] = table[id] with id = (t == 0 ? start_id : tokens[r2, t-1]); out of range -> 0
__global__ void gather_step_kernel(const float* __restrict__ table, int vocab_rows, int E,
                                   const int* __restrict__ tokens, int R2, int L, int t, int start_id,
                                   float* __restrict__ X) {
    size_t total = (size_t)R2 * E;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / E), e = (int)(idx % E);
        int id = t == 0 ? start_id : tokens[(size_t)r * L + t - 1];
        X[idx] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E + e] : 0.f;
    }
}

__global__ void bcast_rows_kernel(const float* __restrict__ S, int B, int tk, int H,
                                  float* __restrict__ out) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * tk * H) return;
    out[idx] = S[(idx / ((size_t)tk * H)) * H + idx % H];
}

__global__ void fill_zero_k(float* __restrict__ x, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = 0.f;
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

extern "C" int d2p_luong_pool_attention(const float* q, const float* keys, const float* values,
                                        const int* mem_len, int B, int k, int tk, int T, int H,
                                        float* ctx, void* stream) {
    D2P_REQUIRE(q && keys && values && mem_len && ctx, "luong attention: null buffer");
    D2P_REQUIRE(tk >= 1 && tk <= ATT_MAX_Q && T <= ATT_MAX_T && H % 4 == 0 && H <= 4 * ATT_THREADS,
                "luong attention: unsupported dims (test_k=%d T=%d H=%d)", tk, T, H);
    size_t smem = ((size_t)tk * H + (size_t)tk * ATT_MAX_T) * sizeof(float);
    luong_pool_attn_kernel<<<B, ATT_THREADS, smem, (cudaStream_t)stream>>>(q, keys, values, mem_len, B, k,
                                                                          tk, T, H, ctx);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t d2p_induction_decode_ws_bytes(int B, int tk, int H) {
    size_t R2 = (size_t)B * tk;
    // x | att | ctx | h | c [R2,H] | gates [R2,4H] | ids, finished [R2] | all_done[64]
    return (R2 * 9 * H) * sizeof(float) + (2 * R2 + 64) * sizeof(int) + 256;
}

// Attention-LSTM action decoder of the induction baseline over R2 = B*test_k rows
// (row = b*test_k + j).  tokens != NULL: teacher forcing (TrainingHelper, <s> id out of
// range -> zero row); tokens == NULL: greedy (start id = A, end id = A-1).
// Initial cell state reproduces the reference's swap (model_induction.py:674-676):
// c := demo_h_summary, h := demo_c_summary.  logits [Tdec, R2, A] time-major.
extern "C" int d2p_induction_decode(const float* keys, const float* values, const int* mem_len, int B,
                                    int k, int tk, int T, int H, const float* h_sum,
                                    const float* c_sum, const float* table, int A, const float* Wcell,
                                    const float* bcell, const float* Wa, const float* proj,
                                    const int* tokens, int Tdec, float* logits, int* out_tokens,
                                    int* lengths, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(keys && values && mem_len && h_sum && c_sum && table && Wcell && bcell && Wa && proj &&
                logits && ws, "induction decode: null buffer");
    D2P_REQUIRE(ws_bytes >= d2p_induction_decode_ws_bytes(B, tk, H), "induction decode: workspace too small");
    const bool greedy = tokens == nullptr;
    if (greedy) D2P_REQUIRE(out_tokens && lengths, "induction decode: greedy needs token/length outputs");
    const int R2 = B * tk, G4 = 4 * H;
    const size_t RH = (size_t)R2 * H;
    float* x = (float*)ws;
    float* att = x + RH; float* ctx = att + RH; float* h = ctx + RH; float* c = h + RH;
    float* gates = c + RH;
    int* ids = (int*)(gates + (size_t)R2 * G4);
    int* finished = ids + R2; int* all_done = finished + R2;
    const int eb = cdiv((long long)RH, 256);
    bcast_rows_kernel<<<eb, 256, 0, st>>>(h_sum, B, tk, H, c);   // swapped on purpose (F8)
    D2P_CHECK_LAUNCH();
    bcast_rows_kernel<<<eb, 256, 0, st>>>(c_sum, B, tk, H, h);
    D2P_CHECK_LAUNCH();
    fill_zero_k<<<eb, 256, 0, st>>>(att, RH);
    D2P_CHECK_LAUNCH();
    if (greedy) {
        greedy_init_kernel<<<cdiv(R2, 256), 256, 0, st>>>(ids, finished, lengths, all_done, R2, tk, A);
        D2P_CHECK_LAUNCH();
    }
    for (int t = 0; t < Tdec; ++t) {
        if (greedy) gather_rows_kernel<<<eb, 256, 0, st>>>(table, A + 1, H, ids, R2, x);
        else gather_step_kernel<<<eb, 256, 0, st>>>(table, A + 1, H, tokens, R2, Tdec, t, A + 1, x);
        D2P_CHECK_LAUNCH();
        // cell input = [emb ; attention_{t-1}], recurrent input h: kernel rows [0,H) [H,2H) [2H,3H)
        D2P_TRY(gemm(st, false, false, R2, G4, H, 1.f, x, H, Wcell, G4, 0.f, gates, G4, bcell, GEMM_CONST_B));
        D2P_TRY(gemm(st, false, false, R2, G4, H, 1.f, att, H, Wcell + (size_t)H * G4, G4, 1.f, gates, G4, nullptr, GEMM_CONST_B));
        D2P_TRY(gemm(st, false, false, R2, G4, H, 1.f, h, H, Wcell + (size_t)2 * H * G4, G4, 1.f, gates, G4, nullptr, GEMM_CONST_B));
        lstm_cell_infer_kernel<<<eb, 256, 0, st>>>(gates, c, h, R2, H, 1.0f);
        D2P_CHECK_LAUNCH();
        D2P_TRY(d2p_luong_pool_attention(h, keys, values, mem_len, B, k, tk, T, H, ctx, stream));
        D2P_TRY(gemm(st, false, false, R2, H, H, 1.f, h, H, Wa, H, 0.f, att, H, nullptr, GEMM_CONST_B));
        D2P_TRY(gemm(st, false, false, R2, H, H, 1.f, ctx, H, Wa + (size_t)H * H, H, 1.f, att, H, nullptr, GEMM_CONST_B));
        float* lt = logits + (size_t)t * R2 * A;
        D2P_TRY(gemm(st, false, false, R2, A, H, 1.f, att, H, proj, A, 0.f, lt, A, nullptr, GEMM_CONST_B));
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
        gather_rows_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(table, vocab_rows, E, ids, R, x);
        D2P_CHECK_LAUNCH();
        D2P_TRY(gemm(st, false, false, R, G4, E, 1.f, x, E, Wx, G4, 0.f, gates, G4, b, GEMM_CONST_B));
        D2P_TRY(gemm(st, false, false, R, G4, H, 1.f, h, H, Wh, G4, 1.f, gates, G4, nullptr, GEMM_CONST_B));
        lstm_cell_infer_kernel<<<cdiv((long long)R * H, 256), 256, 0, st>>>(gates, c, h, R, H, 1.0f);
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
