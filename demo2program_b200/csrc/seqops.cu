// Sequence-side ops of the decoders: token embedding gather with the
// reference's <s> shift, masked softmax / sigmoid cross-entropy with the
// reference's batch normaliser, the embedding-table gradient, and the small
// [B,k,H] group reductions of the summarizer.
//
//  * Teacher forcing (reference models/model_full.py:446-450): decoder input at
//    step t is Emb[token_dim+1] for t = 0 - out of range for the
//    [token_dim+1, E] table, which TF's GPU gather turns into a zero row
//     - and Emb[y_{t-1}] afterwards.
//  * Sequence_Loss (model_full.py:620-657): ce summed over positions
//    t < gt_len and divided by sum(gt_len) over the batch of that decoder
//    instance; action / per losses are then averaged over the k instances.
#include "common.cuh"

namespace d2p {
namespace {

// X[t, r, :] = table[id(t,r)] or 0 if id out of range; id = t==0 ? start_id : tokens[r, t-1]
__global__ void gather_shifted_kernel(const float* __restrict__ table, int vocab_rows, int E,
                                      const int* __restrict__ tokens, int R, int L, int start_id,
                                      float* __restrict__ X) {
    size_t total = (size_t)L * R * E;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int e = (int)(idx % E);
        size_t tr = idx / E;
        int r = (int)(tr % R), t = (int)(tr / R);
        int id = t == 0 ? start_id : tokens[(size_t)r * L + t - 1];
        X[idx] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E + e] : 0.f;
    }
}

// float4 form (E % 4 == 0, 16-byte aligned, fewer than 2^31 elements): 32-bit index math
__global__ void gather_shifted_v4_kernel(const float4* __restrict__ table, int vocab_rows, unsigned E4,
                                         const int* __restrict__ tokens, unsigned R, unsigned L, int start_id,
                                         float4* __restrict__ X) {
    const unsigned total = L * R * E4;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned tr = idx / E4, e = idx - tr * E4;
        const unsigned t = tr / R, r = tr - t * R;
        const int id = t == 0 ? start_id : tokens[(size_t)r * L + t - 1];
        X[idx] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E4 + e] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// dTable[v, e] += sum_{(t,r): id(t,r) == v} dX[t,r,e], two-stage and in a fixed
// order (deterministic): stage 1 reduces a chunk of positions per block into
// partial[chunk, v, e]; stage 2 sums the chunks.
// Block (chunk, 128-wide slice of E): one pass over the chunk's positions, every dX element is
// read once; thread e owns column e of a [vocab_rows][128] shared-memory accumulator, so the
// position order (hence the result) is fixed.
constexpr int EB_COLS = 128;
__global__ void __launch_bounds__(EB_COLS)
embedding_bwd_partial(const float* __restrict__ dX, int vocab_rows, int E,
                      const int* __restrict__ tokens, int R, int L, int start_id,
                      int pos_per_chunk, float* __restrict__ partial) {
    extern __shared__ float eb_acc[];   // [vocab_rows][EB_COLS]
    const int chunk = blockIdx.x, e = blockIdx.y * EB_COLS + threadIdx.x;
    for (int v = 0; v < vocab_rows; ++v) eb_acc[v * EB_COLS + threadIdx.x] = 0.f;
    const int p0 = chunk * pos_per_chunk;
    int p1 = p0 + pos_per_chunk;
    if (p1 > L * R) p1 = L * R;
    if (e < E) {
        int t = p0 / R, r = p0 - t * R;
        for (int p = p0; p < p1; ++p) {
            const int id = t == 0 ? start_id : tokens[(size_t)r * L + t - 1];
            if (id >= 0 && id < vocab_rows) eb_acc[id * EB_COLS + threadIdx.x] += dX[(size_t)p * E + e];
            if (++r == R) { r = 0; ++t; }
        }
        for (int v = 0; v < vocab_rows; ++v)
            partial[((size_t)chunk * vocab_rows + v) * E + e] = eb_acc[v * EB_COLS + threadIdx.x];
    }
}

__global__ void embedding_bwd_reduce(const float* __restrict__ partial, int nchunk, int n,
                                     float* __restrict__ dTable) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int c = 0; c < nchunk; ++c) acc += partial[(size_t)c * n + i];
    dTable[i] += acc;
}

// Per decoder-instance normalisers.  Rows r of the same instance share
// r % nsl.  w[r] = coef / sum_{r' in instance} len[r'];  runlen[r] = max len.
__global__ void seq_weights_kernel(const int* __restrict__ len, int R, int nsl, float coef,
                                   int max_len, float* __restrict__ w, int* __restrict__ runlen) {
    __shared__ int tot[64], mx[64];
    for (int s = threadIdx.x; s < nsl; s += blockDim.x) {
        int a = 0, m = 0;
        for (int r = s; r < R; r += nsl) {
            int l = len[r]; l = l < 0 ? 0 : (l > max_len ? max_len : l);
            a += l; m = l > m ? l : m;
        }
        tot[s] = a; mx[s] = m;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        int s = r % nsl;
        w[r] = tot[s] > 0 ? coef / (float)tot[s] : 0.f;
        if (runlen) runlen[r] = mx[s];
    }
}

// One warp per (t, r) row of logits [T,R,V].  Writes rowloss[t*R+r] and dlogits.
// Rows with t >= runlen[r] are forced to zero logits (dynamic_decode's zero pad).
__global__ void softmax_ce_kernel(float* __restrict__ logits, int T, int R, int V,
                                  const int* __restrict__ labels /*[R, T]*/,
                                  const int* __restrict__ len, const int* __restrict__ runlen,
                                  const float* __restrict__ w, float* __restrict__ rowloss,
                                  float* __restrict__ dlogits) {
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (row >= T * R) return;
    int t = row / R, r = row % R;
    float* x = logits + (size_t)row * V;
    float* dx = dlogits ? dlogits + (size_t)row * V : nullptr;
    if (runlen && t >= runlen[r]) {
        for (int v = lane; v < V; v += 32) { x[v] = 0.f; if (dx) dx[v] = 0.f; }
        if (lane == 0) rowloss[row] = 0.f;
        return;
    }
    if (t >= len[r]) {
        if (dx) for (int v = lane; v < V; v += 32) dx[v] = 0.f;
        if (lane == 0) rowloss[row] = 0.f;
        return;
    }
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, x[v]);
    m = warp_max(m);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(x[v] - m);
    s = warp_sum(s);
    float lse = m + logf(s);
    int y = labels[(size_t)r * T + t];
    float wr = w[r];
    if (lane == 0) rowloss[row] = (y >= 0 && y < V) ? wr * (lse - x[y]) : 0.f;
    if (dx)
        for (int v = lane; v < V; v += 32) {
            float p = expf(x[v] - lse);
            dx[v] = wr * (p - (v == y ? 1.f : 0.f));
        }
}

// sigmoid CE averaged over P (reference model_full.py:651-653); labels [R,T,P] float.
__global__ void sigmoid_ce_kernel(float* __restrict__ logits, int T, int R, int P,
                                  const float* __restrict__ labels, const int* __restrict__ len,
                                  const int* __restrict__ runlen, const float* __restrict__ w,
                                  float* __restrict__ rowloss, float* __restrict__ dlogits) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= T * R) return;
    int t = row / R, r = row % R;
    float* x = logits + (size_t)row * P;
    float* dx = dlogits ? dlogits + (size_t)row * P : nullptr;
    bool padded = runlen && t >= runlen[r];
    bool live = !padded && t < len[r];
    float acc = 0.f, wr = live ? w[r] / (float)P : 0.f;
    for (int p = 0; p < P; ++p) {
        if (padded) x[p] = 0.f;
        float xv = x[p];
        if (live) {
            float z = labels[((size_t)r * T + t) * P + p];
            acc += fmaxf(xv, 0.f) - xv * z + log1pf(expf(-fabsf(xv)));
            if (dx) dx[p] = wr * (sigmoid_f(xv) - z);
        } else if (dx) dx[p] = 0.f;
    }
    rowloss[row] = live ? w[r] * acc / (float)P : 0.f;
}

// out[0] (+)= sum x[0..n) in a fixed order (single block)
__global__ void sum_reduce_kernel(const float* __restrict__ x, int n, float* __restrict__ out,
                                  int accumulate) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += x[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)sh[0];
}

// out[b,:] = alpha * sum_i F[b,i,:] (+ out if accumulate)
__global__ void group_sum_kernel(const float* __restrict__ F, int B, int k, int H, float alpha,
                                 float* __restrict__ out, int accumulate) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    int b = idx / H, u = idx % H;
    float a = 0.f;
    for (int i = 0; i < k; ++i) a += F[((size_t)b * k + i) * H + u];
    out[idx] = (accumulate ? out[idx] : 0.f) + alpha * a;
}

// out[b,i,:] = alpha * S[b,:] (+ out if accumulate)
__global__ void group_bcast_kernel(const float* __restrict__ S, int B, int k, int H, float alpha,
                                   float* __restrict__ out, int accumulate) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * k * H) return;
    int u = (int)(idx % H);
    int b = (int)(idx / ((size_t)k * H));
    out[idx] = (accumulate ? out[idx] : 0.f) + alpha * S[(size_t)b * H + u];
}

// out[b,u] = max_i F[b,i,u] (lowest i among equal maxima), arg[b,u] = that i
__global__ void group_max_kernel(const float* __restrict__ F, int B, int k, int H, float* __restrict__ out,
                                 int* __restrict__ arg) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    int b = idx / H, u = idx % H;
    float m = F[(size_t)b * k * H + u];
    int am = 0;
    for (int i = 1; i < k; ++i) {
        const float v = F[((size_t)b * k + i) * H + u];
        if (v > m) { m = v; am = i; }
    }
    out[idx] = m;
    arg[idx] = am;
}

// dF[b,i,u] = (i == arg[b,u]) ? dout[b,u] : 0
__global__ void group_max_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, int B, int k,
                                     int H, float* __restrict__ dF) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * k * H) return;
    int u = (int)(idx % H);
    int i = (int)((idx / H) % k), b = (int)(idx / ((size_t)k * H));
    const size_t o = (size_t)b * H + u;
    dF[idx] = arg[o] == i ? dout[o] : 0.f;
}

// y = alpha*x + beta*y
__global__ void axpby_kernel(const float* __restrict__ x, float alpha, float* __restrict__ y,
                             float beta, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        y[i] = alpha * x[i] + (beta != 0.f ? beta * y[i] : 0.f);
}

// [T,R,V] time-major logits -> [R,V,T] (the reference's [bs, n, len] layout,
// model_full.py:486-489)
__global__ void logits_to_bvl_kernel(const float* __restrict__ X, int T, int R, int V,
                                     float* __restrict__ Y) {
    size_t total = (size_t)T * R * V;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(idx % T);
        size_t rv = idx / T;
        int v = (int)(rv % V), r = (int)(rv / V);
        Y[idx] = X[((size_t)t * R + r) * V + v];
    }
}

// per [R,T,P] (batch-major, as fed) -> [T,R,P] time-major
__global__ void rtp_to_trp_kernel(const float* __restrict__ X, int R, int T, int P,
                                  float* __restrict__ Y) {
    size_t total = (size_t)T * R * P;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int p = (int)(idx % P);
        size_t tr = idx / P;
        int r = (int)(tr % R), t = (int)(tr / R);
        Y[idx] = X[((size_t)r * T + t) * P + p];
    }
}

// lengths arrive as fp32 (reference models/model_full.py:155-171) -> int32
__global__ void len_to_int_kernel(const float* __restrict__ x, int* __restrict__ y, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (int)x[i];
}

inline int ew_blocks(size_t total) {
    size_t b = (total + 255) / 256, cap = 8 * (size_t)kNumSMs;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" int d2p_embed_shifted(const float* table, int vocab_rows, int E, const int* tokens,
                                 int R, int L, int start_id, float* X, void* stream) {
    D2P_REQUIRE(table && tokens && X && R > 0 && L > 0 && E > 0, "embed: bad arguments");
    if (E % 4 == 0 && (size_t)L * R * E < ((size_t)1 << 31) &&
        ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(X)) & 15) == 0)
        gather_shifted_v4_kernel<<<ew_blocks((size_t)L * R * E / 4), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(table), vocab_rows, (unsigned)(E / 4), tokens, (unsigned)R,
            (unsigned)L, start_id, reinterpret_cast<float4*>(X));
    else
        gather_shifted_kernel<<<ew_blocks((size_t)L * R * E), 256, 0, (cudaStream_t)stream>>>(
            table, vocab_rows, E, tokens, R, L, start_id, X);
    D2P_CHECK_LAUNCH();
    return 0;
}

static int embed_chunks(int R, int L, int* ppc) {
    int pos = R * L;
    // ~32 positions per chunk (each thread walks its chunk serially), at most 256 chunks
    int n = pos < 64 ? 1 : (pos / 32 < 256 ? pos / 32 : 256);
    *ppc = (pos + n - 1) / n;
    return (pos + *ppc - 1) / *ppc;
}

extern "C" size_t d2p_embed_shifted_bwd_ws_bytes(int vocab_rows, int E, int R, int L) {
    int ppc;
    return (size_t)embed_chunks(R, L, &ppc) * vocab_rows * E * sizeof(float);
}

extern "C" int d2p_embed_shifted_bwd(const float* dX, int vocab_rows, int E, const int* tokens,
                                     int R, int L, int start_id, float* dTable, void* ws,
                                     size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(dX && tokens && dTable && ws, "embed bwd: bad arguments");
    int ppc;
    int nchunk = embed_chunks(R, L, &ppc);
    D2P_REQUIRE(ws_bytes >= (size_t)nchunk * vocab_rows * E * sizeof(float), "embed bwd: workspace too small");
    const size_t eb_smem = (size_t)vocab_rows * EB_COLS * sizeof(float);
    D2P_REQUIRE(eb_smem <= 48 * 1024, "embed bwd: vocabulary of %d rows too large", vocab_rows);
    embedding_bwd_partial<<<dim3(nchunk, cdiv(E, EB_COLS)), EB_COLS, eb_smem, st>>>(
        dX, vocab_rows, E, tokens, R, L, start_id, ppc, (float*)ws);
    D2P_CHECK_LAUNCH();
    int n = vocab_rows * E;
    embedding_bwd_reduce<<<cdiv(n, 256), 256, 0, st>>>((const float*)ws, nchunk, n, dTable);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_seq_weights(const int* len, int R, int nsl, float coef, int max_len, float* w,
                               int* runlen, void* stream) {
    D2P_REQUIRE(len && w && R > 0 && nsl > 0 && nsl <= 64 && R % nsl == 0, "seq_weights: bad arguments");
    seq_weights_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(len, R, nsl, coef, max_len, w, runlen);
    D2P_CHECK_LAUNCH();
    return 0;
}

// loss[0] (+)= sum_rows w[r]*ce ; dlogits optional. rowloss: scratch [T*R].
extern "C" int d2p_softmax_ce(float* logits, int T, int R, int V, const int* labels, const int* len,
                              const int* runlen, const float* w, float* rowloss, float* dlogits,
                              float* loss, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(logits && labels && len && w && rowloss && loss, "softmax_ce: null buffer");
    softmax_ce_kernel<<<cdiv((long long)T * R, 8), 256, 0, st>>>(logits, T, R, V, labels, len, runlen,
                                                                w, rowloss, dlogits);
    D2P_CHECK_LAUNCH();
    sum_reduce_kernel<<<1, 256, 0, st>>>(rowloss, T * R, loss, accumulate);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_sigmoid_ce(float* logits, int T, int R, int P, const float* labels,
                              const int* len, const int* runlen, const float* w, float* rowloss,
                              float* dlogits, float* loss, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(logits && labels && len && w && rowloss && loss, "sigmoid_ce: null buffer");
    sigmoid_ce_kernel<<<cdiv((long long)T * R, 256), 256, 0, st>>>(logits, T, R, P, labels, len,
                                                                  runlen, w, rowloss, dlogits);
    D2P_CHECK_LAUNCH();
    sum_reduce_kernel<<<1, 256, 0, st>>>(rowloss, T * R, loss, accumulate);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_group_sum(const float* F, int B, int k, int H, float alpha, float* out,
                             int accumulate, void* stream) {
    D2P_REQUIRE(F && out, "group_sum: null buffer");
    group_sum_kernel<<<cdiv((long long)B * H, 256), 256, 0, (cudaStream_t)stream>>>(F, B, k, H, alpha,
                                                                                out, accumulate);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_group_bcast(const float* S, int B, int k, int H, float alpha, float* out,
                               int accumulate, void* stream) {
    D2P_REQUIRE(S && out, "group_bcast: null buffer");
    group_bcast_kernel<<<cdiv((long long)B * k * H, 256), 256, 0, (cudaStream_t)stream>>>(
        S, B, k, H, alpha, out, accumulate);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_group_max(const float* F, int B, int k, int H, float* out, int* arg, void* stream) {
    D2P_REQUIRE(F && out && arg, "group_max: null buffer");
    D2P_REQUIRE(B > 0 && k > 0 && H > 0, "group_max: empty shape");
    group_max_kernel<<<cdiv((long long)B * H, 256), 256, 0, (cudaStream_t)stream>>>(F, B, k, H, out, arg);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_group_max_bwd(const float* dout, const int* arg, int B, int k, int H, float* dF, void* stream) {
    D2P_REQUIRE(dout && arg && dF, "group_max_bwd: null buffer");
    group_max_bwd_kernel<<<cdiv((long long)B * k * H, 256), 256, 0, (cudaStream_t)stream>>>(dout, arg, B, k, H, dF);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_axpby(const float* x, float alpha, float* y, float beta, size_t n, void* stream) {
    D2P_REQUIRE(x && y, "axpby: null buffer");
    if (n == 0) return 0;
    axpby_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, alpha, y, beta, n);
    D2P_CHECK_LAUNCH();
    return 0;
}

namespace d2p { namespace {
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b,
                            const float* __restrict__ c, float* __restrict__ y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (a[i] + b[i]) + c[i];
}
} }

// y = (a + b) + c  (one pass instead of three axpby launches; same association as the chain it replaces)
extern "C" int d2p_add3(const float* a, const float* b, const float* c, float* y, size_t n, void* stream) {
    D2P_REQUIRE(a && b && c && y, "add3: null buffer");
    if (n == 0) return 0;
    d2p::add3_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, c, y, n);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_logits_to_bvl(const float* X, int T, int R, int V, float* Y, void* stream) {
    D2P_REQUIRE(X && Y, "logits_to_bvl: null buffer");
    logits_to_bvl_kernel<<<ew_blocks((size_t)T * R * V), 256, 0, (cudaStream_t)stream>>>(X, T, R, V, Y);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_rtp_to_trp(const float* X, int R, int T, int P, float* Y, void* stream) {
    D2P_REQUIRE(X && Y, "rtp_to_trp: null buffer");
    rtp_to_trp_kernel<<<ew_blocks((size_t)T * R * P), 256, 0, (cudaStream_t)stream>>>(X, R, T, P, Y);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_len_to_int(const float* x, int* y, int n, void* stream) {
    D2P_REQUIRE(x && y, "len_to_int: null buffer");
    len_to_int_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    D2P_CHECK_LAUNCH();
    return 0;
}

// developer tool: write the GPU global timer (ns) into buf[slot] in stream order - an
// in-graph timeline of the step when called between ops during capture
namespace d2p { namespace {
__global__ void stamp_kernel(unsigned long long* buf, int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[slot] = t;
}
} }
extern "C" int d2p_debug_stamp(unsigned long long* buf, int slot, void* stream) {
    D2P_REQUIRE(buf != nullptr && slot >= 0, "debug_stamp: bad arguments");
    d2p::stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(buf, slot);
    D2P_CHECK_LAUNCH();
    return 0;
}
