// Recurrent step kernels for SMALL row counts (R <= 32: the program decoder, whose
// batch is B = 32 sequences for 50 dependent steps).
//
// With 32 rows a recurrent product is 33 MFLOP: the tcgen05 path would spend its time in
// launch, pipeline fill and a 128-row MMA tile that is 3/4 padding.  Here the product runs
// in exact fp32 on the CUDA cores, spread over 128 CTAs so that each CTA owns 16 output
// columns and reads only a 32 KB weight slab: A ([R, K] rows) and the weight columns are
// staged k-contiguous in shared memory, each 4x4 output tile is computed by 8 threads that
// split K and combine with warp shuffles.
//   forward : z = h_{t-1} * Wh for the i,j,f,o columns of 4 hidden units per CTA, then the
//             BasicLSTMCell epilogue (reference models/model_full.py:498);
//   backward: partial sums of dh_{t-1} = dZ_t * Wh^T, K split over blockIdx.y; the fused
//             element-wise kernel of lstm_tc.cu combines them.
#include "common.cuh"

namespace d2p {
namespace {

constexpr int SK_ROWS = 32, SK_COLS = 16, SK_KC = 512, SK_THREADS = 256, SK_LD = SK_KC + 4;

// acc[i][j] += sum_k As[rt*4+i][k] * Bs[ct*4+j][k]; the 8 threads of a tile take k in
// interleaved groups of 4 (one LDS.128 per row/column per 4 k)
__device__ __forceinline__ void sk_tile_mac(const float* __restrict__ As, const float* __restrict__ Bs,
                                            int kc, int ks, int rt, int ct, float (&acc)[4][4]) {
    const float* ap = As + (size_t)rt * 4 * SK_LD + ks * 4;
    const float* bp = Bs + (size_t)ct * 4 * SK_LD + ks * 4;
#pragma unroll 2
    for (int kq = 0; kq < kc; kq += 32) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(ap + (size_t)i * SK_LD + kq);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(bp + (size_t)j * SK_LD + kq);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = acc[i][j];
                v = fmaf(a[i].x, b[j].x, v); v = fmaf(a[i].y, b[j].y, v);
                v = fmaf(a[i].z, b[j].z, v); v = fmaf(a[i].w, b[j].w, v);
                acc[i][j] = v;
            }
    }
}

__device__ __forceinline__ void sk_reduce8(float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[i][j];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            acc[i][j] = v;
        }
}

// dst[r][0..kc) = src[r*ld + k0 .. +kc) for r < nrows (zero for r >= valid), 16-byte moves.
// Operands that every CTA of the grid reads start at a CTA-dependent offset (rot) so the
// CTAs do not all queue on one L2 slice at a time.
__device__ __forceinline__ void sk_stage_rows(const float* __restrict__ src, size_t ld, int valid, int nrows,
                                              int k0, int kc, float* __restrict__ dst, int rot) {
    const int kq = kc >> 2, total = nrows * kq;
    rot = (rot * 32) % total;
#pragma unroll 8
    for (int idx = threadIdx.x; idx < total; idx += SK_THREADS) {
        int j = idx + rot;
        if (j >= total) j -= total;
        const int r = j / kq, q = j - r * kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < valid) v = *reinterpret_cast<const float4*>(src + (size_t)r * ld + k0 + 4 * q);
        *reinterpret_cast<float4*>(dst + (size_t)r * SK_LD + 4 * q) = v;
    }
}

// ---- forward step ----------------------------------------------------------------------
// grid = H/4 CTAs; CTA j owns hidden units 4j..4j+3, i.e. gate columns g*H + 4j + ul.
__global__ void __launch_bounds__(SK_THREADS, 1)
lstm_step_fwd_skinny(const float* __restrict__ Wh /*[H,4H]*/, int R, int H, float* __restrict__ gates_t,
                     float* __restrict__ cells_t, float* __restrict__ Y_t, const float* __restrict__ hstate,
                     float* __restrict__ hnext, float* __restrict__ cstate, const int* __restrict__ len,
                     int t, float forget_bias) {
    extern __shared__ __align__(16) float sk_smem[];
    float* As = sk_smem;                                 // [SK_ROWS][SK_LD]  h_{t-1}
    float* Bs = As + (size_t)SK_ROWS * SK_LD;            // [SK_COLS][SK_LD]  Wh columns, k-contiguous
    __shared__ float zs[SK_ROWS][SK_COLS + 1];
    const int tid = threadIdx.x, u0 = blockIdx.x * 4;
    const int ks = tid & 7, tile = tid >> 3, rt = tile & 7, ct = tile >> 3;
    // epilogue operands, fetched before the product so their latency is hidden
    const int er = tid >> 2, eu = u0 + (tid & 3);
    const bool ework = tid < SK_ROWS * 4 && er < R;
    float zx[4] = {0.f, 0.f, 0.f, 0.f}, cprev = 0.f;
    bool live = false;
    if (ework) {
        const float* g = gates_t + (size_t)er * 4 * H + eu;
        zx[0] = g[0]; zx[1] = g[H]; zx[2] = g[2 * H]; zx[3] = g[3 * H];
        cprev = cstate[(size_t)er * H + eu];
        live = t < len[er];
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < H; k0 += SK_KC) {
        const int kc = H - k0 < SK_KC ? H - k0 : SK_KC;
        sk_stage_rows(hstate, H, R, SK_ROWS, k0, kc, As, blockIdx.x);
#pragma unroll 8
        for (int idx = tid; idx < kc * 4; idx += SK_THREADS) {
            const int k = idx >> 2, g = idx & 3;
            const float4 w = *reinterpret_cast<const float4*>(Wh + (size_t)(k0 + k) * 4 * H + (size_t)g * H + u0);
            float* d = Bs + (size_t)(g * 4) * SK_LD + k;
            d[0] = w.x; d[SK_LD] = w.y; d[2 * SK_LD] = w.z; d[3 * SK_LD] = w.w;
        }
        __syncthreads();
        sk_tile_mac(As, Bs, kc, ks, rt, ct, acc);
        __syncthreads();
    }
    sk_reduce8(acc);
    if (ks == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) zs[rt * 4 + i][ct * 4 + j] = acc[i][j];
    }
    __syncthreads();
    // cell epilogue: thread = (row, unit)
    if (ework) {
        const int ul = tid & 3;
        const size_t su = (size_t)er * H + eu;
        if (live) {
            float* g = gates_t + (size_t)er * 4 * H + eu;
            const float i = sigmoid_f(zx[0] + zs[er][ul]);
            const float j = tanhf(zx[1] + zs[er][4 + ul]);
            const float f = sigmoid_f(zx[2] + zs[er][8 + ul] + forget_bias);
            const float o = sigmoid_f(zx[3] + zs[er][12 + ul]);
            const float c = cprev * f + i * j;
            const float h = tanhf(c) * o;
            g[0] = i; g[H] = j; g[2 * H] = f; g[3 * H] = o;
            cells_t[su] = c; cstate[su] = c;
            Y_t[su] = h; hnext[su] = h;
        } else {
            cells_t[su] = cprev;
            Y_t[su] = 0.f;
            hnext[su] = hstate[su];
        }
    }
}

// ---- backward step: partial[z][r][u] = sum_{k in chunk z} dZ[r][k] * Wh[u][k] -----------------
// grid = (H/16, 4H/SK_KC)
__global__ void __launch_bounds__(SK_THREADS, 1)
lstm_step_bwd_skinny(const float* __restrict__ dZ /*[R,4H]*/, const float* __restrict__ Wh /*[H,4H]*/,
                     int R, int H, float* __restrict__ partials /*[nz][R][H]*/) {
    extern __shared__ __align__(16) float sk_smem[];
    float* As = sk_smem;
    float* Bs = As + (size_t)SK_ROWS * SK_LD;
    const int tid = threadIdx.x, n0 = blockIdx.x * SK_COLS, G4 = 4 * H;
    const int k0 = blockIdx.y * SK_KC;
    const int kc = G4 - k0 < SK_KC ? G4 - k0 : SK_KC;
    const int ks = tid & 7, tile = tid >> 3, rt = tile & 7, ct = tile >> 3;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    sk_stage_rows(dZ, G4, R, SK_ROWS, k0, kc, As, blockIdx.x);
    sk_stage_rows(Wh + (size_t)n0 * G4, G4, SK_COLS, SK_COLS, k0, kc, Bs, 0);
    __syncthreads();
    sk_tile_mac(As, Bs, kc, ks, rt, ct, acc);
    sk_reduce8(acc);
    if (ks == 0) {
        float* out = partials + (size_t)blockIdx.y * R * H;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rt * 4 + i;
            if (r < R)
                *reinterpret_cast<float4*>(out + (size_t)r * H + n0 + ct * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

constexpr size_t SK_SMEM = (size_t)(SK_ROWS + SK_COLS) * SK_LD * sizeof(float);

}  // namespace

bool lstm_skinny_supported(int R, int H) { return R <= SK_ROWS && H % 32 == 0 && H >= 32; }

int lstm_skinny_nsplit(int H) { return cdiv(4 * H, SK_KC); }

int lstm_skinny_fwd_step(cudaStream_t st, const float* Wh, int R, int H, float* gates_t, float* cells_t,
                         float* Y_t, const float* hstate, float* hnext, float* cstate, const int* len, int t,
                         float forget_bias) {
    static bool attr = false;
    if (!attr) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(lstm_step_fwd_skinny, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)SK_SMEM));
        attr = true;
    }
    lstm_step_fwd_skinny<<<H / 4, SK_THREADS, SK_SMEM, st>>>(Wh, R, H, gates_t, cells_t, Y_t, hstate, hnext,
                                                            cstate, len, t, forget_bias);
    D2P_CHECK_LAUNCH();
    return 0;
}

int lstm_skinny_bwd_step(cudaStream_t st, const float* dZ, const float* Wh, int R, int H, float* partials) {
    static bool attr = false;
    if (!attr) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(lstm_step_bwd_skinny, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)SK_SMEM));
        attr = true;
    }
    lstm_step_bwd_skinny<<<dim3(H / SK_COLS, cdiv(4 * H, SK_KC)), SK_THREADS, SK_SMEM, st>>>(dZ, Wh, R, H,
                                                                                          partials);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace d2p
