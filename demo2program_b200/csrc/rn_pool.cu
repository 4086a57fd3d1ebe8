// Relation-network pooling over the k demonstrations (rn_pool, reference
// models/model_full.py:333-349): rows (b,i,j) = [f_j ; f_i] -> fc 512 -> lrelu
// -> BN -> fc 512 -> lrelu -> BN -> mean over (i,j).
//
// The first FC is factored: [f_j ; f_i] * W1 = f_j * W1[:H] + f_i * W1[H:], so
// only two [B*k, H] x [H, H] products are formed instead of materialising the
// tiled/concatenated [B*k*k, 2H] matrix (13 MB at B=32, k=10).  The second BN
// is folded into the pooled mean (it is affine per channel).
#include "common.cuh"

namespace d2p {

int bn_forward_stats(cudaStream_t, const float*, long long, int, int, int, const float*,
                     const float*, float*, float*, int, float*, void*, size_t);
int bn_apply(cudaStream_t, const float*, float*, long long, int, int, int, const float*, int, int);
int bn_backward(cudaStream_t, const float*, const float*, float*, long long, int, int, int,
                const float*, const float*, float*, float*, int, int, float*, void*, size_t, int,
                int, float*);
int colsum(cudaStream_t, const float*, long long, int, float*, float, void*, size_t);
size_t bn_ws_bytes(long long rows, int C, int nsl);

namespace {

__global__ void pair_add_lrelu(const float* __restrict__ P, const float* __restrict__ Q,
                               const float* __restrict__ bias, int B, int k, int H,
                               float* __restrict__ A) {
    size_t total = (size_t)B * k * k * H;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int u = (int)(idx % H);
        size_t row = idx / H;
        int j = (int)(row % k), i = (int)((row / k) % k), b = (int)(row / ((size_t)k * k));
        float z = P[((size_t)b * k + j) * H + u] + Q[((size_t)b * k + i) * H + u] + bias[u];
        A[idx] = lrelu_f(z);
    }
}

// float4 form of pair_add_lrelu (H % 4 == 0, fewer than 2^31 elements): 32-bit index math
__global__ void pair_add_lrelu_v4(const float4* __restrict__ P, const float4* __restrict__ Q,
                                  const float4* __restrict__ bias, unsigned B, unsigned k, unsigned H4,
                                  float4* __restrict__ A) {
    const unsigned total = B * k * k * H4;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned row = idx / H4, u = idx - row * H4;
        const unsigned bi = row / k, j = row - bi * k;      // bi = b*k + i
        const unsigned b = bi / k;
        const float4 p = P[(b * k + j) * H4 + u], q = Q[bi * H4 + u], c = bias[u];
        A[idx] = make_float4(lrelu_f(p.x + q.x + c.x), lrelu_f(p.y + q.y + c.y), lrelu_f(p.z + q.z + c.z),
                             lrelu_f(p.w + q.w + c.w));
    }
}

__global__ void lrelu_inplace(float* __restrict__ x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        x[i] = lrelu_f(x[i]);
}

// pooled[b,u] = scale[u] * mean_{ij} A[b,i,j,u] + shift[u]
__global__ void pooled_mean(const float* __restrict__ A, int B, int kk, int H,
                            const float* __restrict__ scale, const float* __restrict__ shift,
                            float* __restrict__ pooled) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    int b = idx / H, u = idx % H;
    float a = 0.f;
    for (int q = 0; q < kk; ++q) a += A[((size_t)b * kk + q) * H + u];
    pooled[idx] = scale[u] * (a / (float)kk) + shift[u];
}

// the same with the kk rows of a batch element split over 4 thread rows (fixed-order combine):
// grid (B, H / 128), block (128, 4)
__global__ void pooled_mean_v2(const float* __restrict__ A, int kk, int H, const float* __restrict__ scale,
                               const float* __restrict__ shift, float* __restrict__ pooled) {
    __shared__ float part[4][128];
    const int b = blockIdx.x, u = blockIdx.y * 128 + threadIdx.x, y = threadIdx.y;
    float a = 0.f;
    for (int q = y; q < kk; q += 4) a += A[((size_t)b * kk + q) * H + u];
    part[y][threadIdx.x] = a;
    __syncthreads();
    if (y == 0) {
        const float s = ((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x];
        pooled[(size_t)b * H + u] = scale[u] * (s / (float)kk) + shift[u];
    }
}

__global__ void bcast_pairs(const float* __restrict__ dpooled, int B, int kk, int H,
                            float* __restrict__ dY) {
    size_t total = (size_t)B * kk * H;
    float inv = 1.f / (float)kk;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int u = (int)(idx % H);
        int b = (int)(idx / ((size_t)kk * H));
        dY[idx] = dpooled[(size_t)b * H + u] * inv;
    }
}

// dP[b,j,u] = sum_i dZ[b,i,j,u] ; dQ[b,i,u] = sum_j dZ[b,i,j,u]
__global__ void pair_reduce(const float* __restrict__ dZ, int B, int k, int H,
                            float* __restrict__ dP, float* __restrict__ dQ) {
    size_t total = (size_t)B * k * H;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int u = (int)(idx % H);
        int a = (int)((idx / H) % k), b = (int)(idx / ((size_t)k * H));
        float sp = 0.f, sq = 0.f;
        for (int o = 0; o < k; ++o) {
            sp += dZ[(((size_t)b * k + o) * k + a) * H + u];
            sq += dZ[(((size_t)b * k + a) * k + o) * H + u];
        }
        dP[idx] = sp; dQ[idx] = sq;
    }
}

inline int ewb(size_t total) {
    size_t b = (total + 255) / 256, cap = 8 * (size_t)kNumSMs;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

struct Plan { size_t a, b, c, d, e, coef, part, part_bytes, total; };

Plan plan(int B, int k, int H) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t big = al((size_t)B * k * k * H * sizeof(float));
    size_t sml = al((size_t)B * k * H * sizeof(float));
    Plan p;
    p.a = 0; p.b = big; p.c = 2 * big; p.d = 3 * big; p.e = 3 * big + sml;
    p.coef = p.e + sml;
    p.part = p.coef + al(2 * (size_t)H * sizeof(float));
    p.part_bytes = al(bn_ws_bytes((long long)B * k * k, H, 1));
    p.total = p.part + p.part_bytes;
    return p;
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" size_t d2p_rn_pool_saved_floats(int B, int k, int H) {
    return 2 * (size_t)B * k * k * H + 8 * (size_t)H;
}
extern "C" size_t d2p_rn_pool_ws_bytes(int B, int k, int H) { return plan(B, k, H).total; }

extern "C" int d2p_rn_pool_fwd(const float* F, int B, int k, int H, const d2p_fc_bn* fc1,
                               const d2p_fc_bn* fc2, float* pooled, float* saved, int training,
                               void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(F && fc1 && fc2 && pooled && saved && ws, "rn_pool fwd: null buffer");
    Plan p = plan(B, k, H);
    D2P_REQUIRE(ws_bytes >= p.total, "rn_pool fwd: workspace too small");
    char* w = (char*)ws;
    const int Bk = B * k, kk = k * k;
    const long long rows = (long long)B * kk;
    float* A1 = saved; float* A2 = A1 + rows * H;
    float* st1 = A2 + rows * H; float* st2 = st1 + 4 * H;
    float* P = (float*)(w + p.d); float* Q = (float*)(w + p.e);
    float* X2 = (float*)(w + p.a);
    D2P_TRY(gemm(st, false, false, Bk, H, H, 1.f, F, H, fc1->w, H, 0.f, P, H, nullptr, GEMM_CONST_B));
    D2P_TRY(gemm(st, false, false, Bk, H, H, 1.f, F, H, fc1->w + (size_t)H * H, H, 0.f, Q, H, nullptr, GEMM_CONST_B));
    if (H % 4 == 0 && rows * H < (1LL << 31) &&
        ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(fc1->b) |
          reinterpret_cast<uintptr_t>(A1)) & 15) == 0)
        pair_add_lrelu_v4<<<ewb(rows * H / 4), 256, 0, st>>>(
            reinterpret_cast<const float4*>(P), reinterpret_cast<const float4*>(Q),
            reinterpret_cast<const float4*>(fc1->b), (unsigned)B, (unsigned)k, (unsigned)(H / 4),
            reinterpret_cast<float4*>(A1));
    else
        pair_add_lrelu<<<ewb(rows * H), 256, 0, st>>>(P, Q, fc1->b, B, k, H, A1);
    D2P_CHECK_LAUNCH();
    D2P_TRY(bn_forward_stats(st, A1, rows, H, 1, 1, fc1->gamma, fc1->beta, fc1->moving_mean,
                             fc1->moving_var, training, st1, w + p.part, p.part_bytes));
    D2P_TRY(bn_apply(st, A1, X2, rows, H, 1, 1, st1, 0, 0));
    D2P_TRY(gemm(st, false, false, (int)rows, H, H, 1.f, X2, H, fc2->w, H, 0.f, A2, H, fc2->b, GEMM_CONST_B));
    lrelu_inplace<<<ewb(rows * H), 256, 0, st>>>(A2, rows * H);
    D2P_CHECK_LAUNCH();
    D2P_TRY(bn_forward_stats(st, A2, rows, H, 1, 1, fc2->gamma, fc2->beta, fc2->moving_mean,
                             fc2->moving_var, training, st2, w + p.part, p.part_bytes));
    if (H % 128 == 0)
        pooled_mean_v2<<<dim3(B, H / 128), dim3(128, 4), 0, st>>>(A2, kk, H, st2 + 2 * H, st2 + 3 * H, pooled);
    else
        pooled_mean<<<cdiv((long long)B * H, 256), 256, 0, st>>>(A2, B, kk, H, st2 + 2 * H, st2 + 3 * H, pooled);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t d2p_rn_pool_bwd_hold_floats(int B, int k, int H) {
    return 2 * (size_t)B * k * k * H + 2 * (size_t)B * k * H;
}

// dF += d(pooled)/dF.  Parameter grads accumulate into fc1/fc2 grad pointers.
// phases / hold: with a caller-owned `hold` buffer (d2p_rn_pool_bwd_hold_floats) the call can be
// split: phases & 1 = the data path (dF, plus the BatchNorm scale/shift gradients that fall out of
// it) leaving both dZ and the pair sums in `hold`; phases & 2 = the weight / bias gradients from
// `hold` - they may run later, on another stream.  hold == NULL: one call does everything.
extern "C" int d2p_rn_pool_bwd(const float* F, int B, int k, int H, const d2p_fc_bn* fc1,
                               const d2p_fc_bn* fc2, const float* dpooled, const float* saved,
                               float* dF, int training, void* ws, size_t ws_bytes, int phases,
                               float* hold, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(F && fc1 && fc2 && dpooled && saved && dF && ws, "rn_pool bwd: null buffer");
    D2P_REQUIRE(fc1->dw && fc1->db && fc1->dgamma && fc1->dbeta && fc2->dw && fc2->db &&
                fc2->dgamma && fc2->dbeta, "rn_pool bwd: null grad buffer");
    D2P_REQUIRE(hold != nullptr || phases == 3, "rn_pool bwd: a split call needs a hold buffer");
    Plan p = plan(B, k, H);
    D2P_REQUIRE(ws_bytes >= p.total, "rn_pool bwd: workspace too small");
    char* w = (char*)ws;
    const int Bk = B * k, kk = k * k;
    const long long rows = (long long)B * kk;
    const float* A1 = saved; const float* A2 = A1 + rows * H;
    const float* st1 = A2 + rows * H; const float* st2 = st1 + 4 * H;
    float* X2 = (float*)(w + p.a); float* dY = (float*)(w + p.b); float* dZ = (float*)(w + p.c);
    float* dP = (float*)(w + p.d); float* dQ = (float*)(w + p.e);
    float* coef = (float*)(w + p.coef);
    void* part = w + p.part;
    if (hold != nullptr) {
        float* dZ2 = hold; float* dZ1 = hold + rows * H;
        float* hP = dZ1 + rows * H; float* hQ = hP + (size_t)Bk * H;
        if (phases & 1) {
            bcast_pairs<<<ewb(rows * H), 256, 0, st>>>(dpooled, B, kk, H, dY);
            D2P_CHECK_LAUNCH();
            D2P_TRY(bn_backward(st, A2, dY, dZ2, rows, H, 1, 1, fc2->gamma, st2, fc2->dgamma, fc2->dbeta,
                                training, 1, coef, part, p.part_bytes, 0, 0, nullptr));
            D2P_TRY(gemm(st, false, true, (int)rows, H, H, 1.f, dZ2, H, fc2->w, H, 0.f, dY, H, nullptr, GEMM_CONST_B));
            D2P_TRY(bn_backward(st, A1, dY, dZ1, rows, H, 1, 1, fc1->gamma, st1, fc1->dgamma, fc1->dbeta,
                                training, 1, coef, part, p.part_bytes, 0, 0, nullptr));
            pair_reduce<<<ewb((size_t)Bk * H), 256, 0, st>>>(dZ1, B, k, H, hP, hQ);
            D2P_CHECK_LAUNCH();
            D2P_TRY(gemm(st, false, true, Bk, H, H, 1.f, hP, H, fc1->w, H, 1.f, dF, H, nullptr, GEMM_CONST_B));
            D2P_TRY(gemm(st, false, true, Bk, H, H, 1.f, hQ, H, fc1->w + (size_t)H * H, H, 1.f, dF, H, nullptr, GEMM_CONST_B));
        }
        if (phases & 2) {
            D2P_TRY(colsum(st, dZ2, rows, H, fc2->db, 1.f, part, p.part_bytes));
            D2P_TRY(bn_apply(st, A1, X2, rows, H, 1, 1, st1, 0, 0));
            D2P_TRY(gemm(st, true, false, H, H, (int)rows, 1.f, X2, H, dZ2, H, 1.f, fc2->dw, H));
            D2P_TRY(colsum(st, dZ1, rows, H, fc1->db, 1.f, part, p.part_bytes));
            D2P_TRY(gemm(st, true, false, H, H, Bk, 1.f, F, H, hP, H, 1.f, fc1->dw, H));
            D2P_TRY(gemm(st, true, false, H, H, Bk, 1.f, F, H, hQ, H, 1.f, fc1->dw + (size_t)H * H, H));
        }
        return 0;
    }
    // second block
    bcast_pairs<<<ewb(rows * H), 256, 0, st>>>(dpooled, B, kk, H, dY);
    D2P_CHECK_LAUNCH();
    D2P_TRY(bn_backward(st, A2, dY, dZ, rows, H, 1, 1, fc2->gamma, st2, fc2->dgamma, fc2->dbeta,
                        training, 1, coef, part, p.part_bytes, 0, 0, nullptr));
    D2P_TRY(colsum(st, dZ, rows, H, fc2->db, 1.f, part, p.part_bytes));
    D2P_TRY(bn_apply(st, A1, X2, rows, H, 1, 1, st1, 0, 0));
    D2P_TRY(gemm(st, true, false, H, H, (int)rows, 1.f, X2, H, dZ, H, 1.f, fc2->dw, H));
    D2P_TRY(gemm(st, false, true, (int)rows, H, H, 1.f, dZ, H, fc2->w, H, 0.f, dY, H, nullptr, GEMM_CONST_B));  // dX2
    // first block
    D2P_TRY(bn_backward(st, A1, dY, dZ, rows, H, 1, 1, fc1->gamma, st1, fc1->dgamma, fc1->dbeta,
                        training, 1, coef, part, p.part_bytes, 0, 0, nullptr));
    D2P_TRY(colsum(st, dZ, rows, H, fc1->db, 1.f, part, p.part_bytes));
    pair_reduce<<<ewb((size_t)Bk * H), 256, 0, st>>>(dZ, B, k, H, dP, dQ);
    D2P_CHECK_LAUNCH();
    D2P_TRY(gemm(st, true, false, H, H, Bk, 1.f, F, H, dP, H, 1.f, fc1->dw, H));
    D2P_TRY(gemm(st, true, false, H, H, Bk, 1.f, F, H, dQ, H, 1.f, fc1->dw + (size_t)H * H, H));
    D2P_TRY(gemm(st, false, true, Bk, H, H, 1.f, dP, H, fc1->w, H, 1.f, dF, H, nullptr, GEMM_CONST_B));
    D2P_TRY(gemm(st, false, true, Bk, H, H, 1.f, dQ, H, fc1->w + (size_t)H * H, H, 1.f, dF, H, nullptr, GEMM_CONST_B));
    return 0;
}
