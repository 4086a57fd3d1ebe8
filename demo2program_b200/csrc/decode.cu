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

}  // namespace
}  // namespace d2p

using namespace d2p;

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
