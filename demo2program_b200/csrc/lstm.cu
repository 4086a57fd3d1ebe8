// LSTM sequence op: tf.nn.dynamic_rnn(BasicLSTMCell(H), sequence_length=len)
// (reference models/model_full.py:244-258, 265-277) and the teacher-forced
// decoders' cell loop (model_full.py:465-471), forward and backward.
//
// BasicLSTMCell semantics: z = [x, h] * kernel + bias;
// i, j, f, o = split(z, 4); c' = c*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j);
// h' = tanh(c')*sigmoid(o).  dynamic_rnn masking (A.5): for t >= len[r] the
// output row is zero and the state is copied through.
//
// Layout: time-major.  X [T,R,In], gates [T,R,4H], Y/cells [T,R,H].  The
// input contraction X*Wx is hoisted out of the recurrence into one GEMM over
// all T*R rows; each step then adds h_{t-1}*Wh and runs the fused gate kernel.
#include "common.cuh"

namespace d2p {

int colsum(cudaStream_t, const float*, long long, int, float*, float, void*, size_t);
size_t bn_ws_bytes(long long rows, int C, int nsl);
// tensor-core recurrence (lstm_tc.cu)
bool lstm_tc_supported(int R, int H);
int lstm_seq_fwd_tc(cudaStream_t, const float*, int, int, int, int, const int*, const float*,
                    const float*, const float*, const float*, float, float*, float*, float*, float*,
                    float*, int);
int lstm_seq_bwd_tc(cudaStream_t, const float*, int, int, int, int, const int*, const float*,
                    const float*, const float*, const float*, float*, const float*, const float*,
                    const float*, const float*, float*, float*, float*, float*, float*, void*, size_t,
                    int);

namespace {

__global__ void lstm_gates_fwd(float* __restrict__ G /*[R,4H] step t, in: preact, out: activated*/,
                               float* __restrict__ cells_t, float* __restrict__ Y_t,
                               float* __restrict__ hstate, float* __restrict__ cstate,
                               const int* __restrict__ len, int t, int R, int H, float forget_bias) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * H) return;
    int r = idx / H, u = idx % H;
    float cprev = cstate[idx];
    if (t < len[r]) {
        float* g = G + (size_t)r * 4 * H;
        float i = sigmoid_f(g[u]);
        float j = tanhf(g[H + u]);
        float f = sigmoid_f(g[2 * H + u] + forget_bias);
        float o = sigmoid_f(g[3 * H + u]);
        float c = cprev * f + i * j;
        float h = tanhf(c) * o;
        g[u] = i; g[H + u] = j; g[2 * H + u] = f; g[3 * H + u] = o;
        cells_t[idx] = c; cstate[idx] = c;
        Y_t[idx] = h; hstate[idx] = h;
    } else {
        cells_t[idx] = cprev;
        Y_t[idx] = 0.f;
    }
}

__global__ void lstm_gates_bwd(float* __restrict__ G /*in: activated gates, out: dZ*/,
                               const float* __restrict__ cells_t,
                               const float* __restrict__ cells_prev /*nullable -> c0*/,
                               const float* __restrict__ c0 /*nullable -> 0*/,
                               const float* __restrict__ dY_t /*nullable*/,
                               float* __restrict__ dhs, float* __restrict__ dcs,
                               const int* __restrict__ len, int t, int R, int H) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * H) return;
    int r = idx / H, u = idx % H;
    float* g = G + (size_t)r * 4 * H;
    if (t < len[r]) {
        float i = g[u], j = g[H + u], f = g[2 * H + u], o = g[3 * H + u];
        float c = cells_t[idx];
        float cprev = cells_prev ? cells_prev[idx] : (c0 ? c0[idx] : 0.f);
        float dh = dhs[idx] + (dY_t ? dY_t[idx] : 0.f);
        float tc = tanhf(c);
        float d_o = dh * tc * o * (1.f - o);
        float dc = dcs[idx] + dh * o * (1.f - tc * tc);
        g[u] = dc * j * i * (1.f - i);
        g[H + u] = dc * i * (1.f - j * j);
        g[2 * H + u] = dc * cprev * f * (1.f - f);
        g[3 * H + u] = d_o;
        dcs[idx] = dc * f;
        dhs[idx] = 0.f;   // the recurrent GEMM accumulates the new dh on top
    } else {
        g[u] = 0.f; g[H + u] = 0.f; g[2 * H + u] = 0.f; g[3 * H + u] = 0.f;
    }
}

__global__ void copy_or_zero(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src ? src[i] : 0.f;
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" int d2p_lstm_seq_fwd(const float* X, int T, int R, int In, int H, const int* len,
                                const float* h0, const float* c0, const float* W, const float* b,
                                float forget_bias, float* Y, float* hT, float* cT, float* gates,
                                float* cells, int phases, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(X && len && W && b && Y && hT && cT && gates && cells, "lstm fwd: null buffer");
    D2P_REQUIRE(T > 0 && R > 0 && In > 0 && H > 0, "lstm fwd: bad dims");
    if (lstm_tc_supported(R, H))
        return lstm_seq_fwd_tc(st, X, T, R, In, H, len, h0, c0, W, b, forget_bias, Y, hT, cT, gates, cells,
                               phases);
    const int G4 = 4 * H;
    const float* Wx = W;
    const float* Wh = W + (size_t)In * G4;
    size_t RH = (size_t)R * H;
    int eb = cdiv(RH, 256);
    // hoisted input contraction for all steps: gates = X*Wx + b
    if (phases & D2P_LSTM_INPUT)
        D2P_TRY(gemm(st, false, false, T * R, G4, In, 1.f, X, In, Wx, G4, 0.f, gates, G4, b, GEMM_CONST_B));
    if (!(phases & D2P_LSTM_RECUR)) return 0;
    copy_or_zero<<<eb, 256, 0, st>>>(hT, h0, RH);
    D2P_CHECK_LAUNCH();
    copy_or_zero<<<eb, 256, 0, st>>>(cT, c0, RH);
    D2P_CHECK_LAUNCH();
    for (int t = 0; t < T; ++t) {
        float* Gt = gates + (size_t)t * R * G4;
        if (t > 0 || h0 != nullptr)
            D2P_TRY(gemm(st, false, false, R, G4, H, 1.f, hT, H, Wh, G4, 1.f, Gt, G4, nullptr, GEMM_CONST_B));
        lstm_gates_fwd<<<eb, 256, 0, st>>>(Gt, cells + t * RH, Y + t * RH, hT, cT, len, t, R, H,
                                           forget_bias);
        D2P_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" size_t d2p_lstm_seq_bwd_ws_bytes(int T, int R, int H) {
    return bn_ws_bytes((long long)T * R, 4 * H, 1);
}

// gates is consumed: on return it holds dZ [T,R,4H].
extern "C" int d2p_lstm_seq_bwd(const float* X, int T, int R, int In, int H, const int* len,
                                const float* h0, const float* c0, const float* W, const float* Y,
                                float* gates, const float* cells, const float* dY,
                                const float* dhT, const float* dcT, float* dX, float* dW, float* db,
                                float* dh0, float* dc0, void* ws, size_t ws_bytes, int phases,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(X && len && W && Y && gates && cells && dW && db && dh0 && dc0, "lstm bwd: null buffer");
    if (lstm_tc_supported(R, H))
        return lstm_seq_bwd_tc(st, X, T, R, In, H, len, h0, c0, W, Y, gates, cells, dY, dhT, dcT, dX, dW,
                               db, dh0, dc0, ws, ws_bytes, phases);
    const int G4 = 4 * H;
    const float* Wx = W;
    const float* Wh = W + (size_t)In * G4;
    float* dWx = dW;
    float* dWh = dW + (size_t)In * G4;
    size_t RH = (size_t)R * H;
    int eb = cdiv(RH, 256);
    if (phases & D2P_LSTM_BWD_RECUR) {
    copy_or_zero<<<eb, 256, 0, st>>>(dh0, dhT, RH);   // dh0/dc0 double as the running dh/dc
        D2P_CHECK_LAUNCH();
        copy_or_zero<<<eb, 256, 0, st>>>(dc0, dcT, RH);
        D2P_CHECK_LAUNCH();
        for (int t = T - 1; t >= 0; --t) {
            float* Gt = gates + (size_t)t * R * G4;
            lstm_gates_bwd<<<eb, 256, 0, st>>>(Gt, cells + t * RH, t > 0 ? cells + (t - 1) * RH : nullptr,
                                               c0, dY ? dY + t * RH : nullptr, dh0, dc0, len, t, R, H);
            D2P_CHECK_LAUNCH();
            if (t > 0 || h0 != nullptr)   // dh_{t-1} += dZ_t * Wh^T
                D2P_TRY(gemm(st, false, true, R, H, G4, 1.f, Gt, G4, Wh, G4, 1.f, dh0, H, nullptr, GEMM_CONST_B));
        }
        if (dX) D2P_TRY(gemm(st, false, true, T * R, In, G4, 1.f, gates, G4, Wx, G4, 0.f, dX, In, nullptr, GEMM_CONST_B));
    }
    if (!(phases & D2P_LSTM_BWD_PARAMS)) return 0;
    // parameter gradients from the full dZ
    if (!(phases & D2P_LSTM_BWD_NO_DWX))
        D2P_TRY(gemm(st, true, false, In, G4, T * R, 1.f, X, In, gates, G4, 1.f, dWx, G4));
    if (T > 1)
        D2P_TRY(gemm(st, true, false, H, G4, (T - 1) * R, 1.f, Y, H, gates + (size_t)R * G4, G4, 1.f, dWh, G4));
    if (h0) D2P_TRY(gemm(st, true, false, H, G4, R, 1.f, h0, H, gates, G4, 1.f, dWh, G4));
    D2P_TRY(colsum(st, gates, (long long)T * R, G4, db, 1.f, ws, ws_bytes));
    return 0;
}
