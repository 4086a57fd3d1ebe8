// Demonstration-frame CNN encoder (State_Encoder, reference
// models/model_full.py:216-231 through ops.conv2d, models/ops.py:27-33):
// per layer slim.conv2d(3x3, stride 2, SAME, bias) -> lrelu(0.2) -> BatchNorm
// with train-mode batch statistics per demonstration index (the reference
// instantiates the encoder k times, model_full.py:373-376).
//
// Layout: frames arrive as stored, [B, k, T, h, w, d] (u8 for Karel's bool /
// ViZDoom's 0..255, or f32 as the reference feeds); frame n = (b*k+i)*T + t,
// slice(n) = (n / T) % k.  Activations are NHWC fp32.  The BatchNorm of layer
// l is applied on the fly when layer l+1 loads its input (scale/shift per
// slice), so each layer is one pass: conv + bias + lrelu -> a_l, then a
// two-stage statistics reduction.  The final feature is written time-major
// [T, R, F] (R = B*k) for the LSTM.
//
// TF SAME geometry: pad_before = pad_total / 2, so every
// Karel layer and ViZDoom L1-L4 pad (0 top/left, 1 bottom/right).
#include "common.cuh"
#include "conv_tc.cuh"

namespace d2p {

int bn_forward_finalize(cudaStream_t, const float2*, int, long long, int, int, const float*, const float*, float*,
                        float*, float*);
int bn_forward_stats(cudaStream_t, const float*, long long, int, int, int, const float*,
                     const float*, float*, float*, int, float*, void*, size_t);
int bn_apply(cudaStream_t, const float*, float*, long long, int, int, int, const float*, int, int);
int bn_backward(cudaStream_t, const float*, const float*, float*, long long, int, int, int,
                const float*, const float*, float*, float*, int, int, float*, void*, size_t, int,
                int, float*);
int colsum(cudaStream_t, const float*, long long, int, float*, float, void*, size_t);
int bn_backward_db(cudaStream_t, const float*, const float*, float*, long long, int, int, int, const float*,
                   const float*, float*, float*, float*, int, float*, void*, size_t);
size_t bn_ws_bytes(long long rows, int C, int nsl);
int permute_frames(cudaStream_t, const float*, float*, int, int, int, int);
int unpermute_frames(cudaStream_t, const float*, float*, int, int, int, int);
// fused single-kernel Karel forward (conv_fused.cu)
bool conv_fused_supported(const d2p_conv_desc* d, int training, size_t ws_bytes);
int conv_fused_fwd(cudaStream_t st, const d2p_conv_desc* d, const void* frames, float* feat, float* saved,
                   int training, void* ws);
bool conv_fused_bwd_supported(const d2p_conv_desc* d, int training, size_t ws_bytes);
size_t conv_fused_ws_bytes(const d2p_conv_desc* d);
size_t conv_fused_bwd_ws_bytes(const d2p_conv_desc* d);
int conv_fused_bwd(cudaStream_t st, const d2p_conv_desc* d, const void* frames, const float* dfeat,
                   const float* saved, void* ws);

namespace {

int g_conv_rgb = 1;             // d2p_conv_set_tc bit 4 clear: the direct RGB-layer forward kernel

constexpr int PX = 32;          // output pixels per tile
constexpr int MAXC = 48;        // max channels of any layer on the path

typedef ConvGeo Geo;

template <typename IN_T>
__device__ __forceinline__ float load_in(const IN_T* in, size_t idx) { return (float)in[idx]; }

#include "conv_v2.cuh"
// RGB input layer (CIN = 3, COUT = 16, u8 frames), forward: one thread per output pixel, the 27 x 16 weights as
// CONSTANT-BANK operands of the FMAs (copied device-to-device into c_rgb_w in front of the launch: `ncu --set full`
// showed the former broadcast float4 reads from shared memory filling 92 % of the L1 data pipe - a broadcast LDS.128
// still costs four wavefronts - while the FMA pipe sat at 32 %; profiles/r03s_ncu_full_c4_rgb_bn.txt),
// a = lrelu(conv + bias) written as one 64-byte row, and the
// BatchNorm (sum, sum of squares) of the block's pixels reduced in a fixed order into
// partial[(slice*nchunk + chunk)*16 + c] (a block never straddles a demonstration, i.e. a slice).
constexpr int RGB_THREADS = 256, RGB_PPT = 4;
__constant__ float c_rgb_w[27 * 16];
__constant__ float c_rgb_b[16];
__global__ void __launch_bounds__(RGB_THREADS)
conv_rgb_fwd_kernel(Geo g, const uint8_t* __restrict__ in, const float* __restrict__ W,
                    const float* __restrict__ bias, float* __restrict__ out, float2* __restrict__ partial,
                    int cpd, int nchunk) {
    __shared__ float red[RGB_THREADS / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    (void)W; (void)bias;          // staged into c_rgb_w / c_rgb_b by the launcher
    const int r = blockIdx.x / cpd, j = blockIdx.x - r * cpd;
    const int HW = g.OH * g.OW, P = g.T * HW;
    const int q0 = j * (RGB_THREADS * RGB_PPT);
    float s1[16], s2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    for (int it = 0; it < RGB_PPT; ++it) {
        const int q = q0 + it * RGB_THREADS + tid;
        if (q >= P) break;
        const int t = q / HW, rem = q - t * HW, oy = rem / g.OW, ox = rem - oy * g.OW;
        const uint8_t* fin = in + (size_t)(r * g.T + t) * g.IH * g.IW * 3;
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = c_rgb_b[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy + ky - g.PT;
            if (iy < 0 || iy >= g.IH) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox + kx - g.PL;
                if (ix < 0 || ix >= g.IW) continue;
                const uint8_t* px = fin + ((size_t)iy * g.IW + ix) * 3;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = (float)__ldg(px + ci);
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[c] = fmaf(v, c_rgb_w[((ky * 3 + kx) * 3 + ci) * 16 + c], acc[c]);
                }
            }
        }
        float4* o = reinterpret_cast<float4*>(out + ((size_t)r * P + q) * 16);
#pragma unroll
        for (int c = 0; c < 16; ++c) { acc[c] = lrelu_f(acc[c]); s1[c] += acc[c]; s2[c] = fmaf(acc[c], acc[c], s2[c]); }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) o[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
    }
    if (partial == nullptr) return;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        s1[c] = warp_sum(s1[c]); s2[c] = warp_sum(s2[c]);
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { red[warp][c] = s1[c]; red[warp][16 + c] = s2[c]; }
    }
    __syncthreads();
    if (tid < 16) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < RGB_THREADS / 32; ++w) { a += red[w][tid]; b += red[w][16 + tid]; }
        const int chunk = (r / g.k) * cpd + j, sl = r % g.k;
        partial[((size_t)sl * nchunk + chunk) * 16 + tid] = make_float2(a, b);
    }
}
inline int rgb_cpd(const Geo& g) { return cdiv((long long)g.T * g.OH * g.OW, RGB_THREADS * RGB_PPT); }
inline bool rgb_fwd_supported(const d2p_conv_desc* d, const Geo& g) {
    return d->frames_dtype == D2P_U8 && g.CIN == 3 && g.COUT == 16 && g.N % g.T == 0 && (g.N / g.T) % g.k == 0;
}
inline size_t rgb_fwd_ws_bytes(const Geo& g) {
    return ((size_t)g.k * (g.N / g.T / g.k) * rgb_cpd(g) * 16 * 2 + (size_t)g.k * 16) * sizeof(float);
}

// dW[tap, o] += sum_blk partial[blk, tap, o]
__global__ void conv_dw_reduce(const float* __restrict__ partial, int nblk, int n,
                               float* __restrict__ dW) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * n + i];
    dW[i] += (float)s;
}

inline int dw_blocks(long long npix, int* ppb) {
    long long target = 2LL * kNumSMs;
    long long per = (npix + target - 1) / target;
    per = (per + PX - 1) / PX * PX;
    if (per < 8 * PX) per = 8 * PX;
    *ppb = (int)per;
    return (int)((npix + per - 1) / per);
}

int fill_geo(const d2p_conv_desc* d, int layer, Geo* g) {
    int ih = d->h, iw = d->w, cin = d->d;
    for (int l = 0; l <= layer; ++l) {
        int oh = (ih + 1) / 2, ow = (iw + 1) / 2;
        int pth = (oh - 1) * 2 + 3 - ih; if (pth < 0) pth = 0;
        int ptw = (ow - 1) * 2 + 3 - iw; if (ptw < 0) ptw = 0;
        if (l == layer) {
            g->N = d->B * d->k * d->T; g->IH = ih; g->IW = iw; g->CIN = cin;
            g->OH = oh; g->OW = ow; g->COUT = d->layers[l].cout;
            g->PT = pth / 2; g->PL = ptw / 2; g->T = d->T; g->k = d->k;
        }
        ih = oh; iw = ow; cin = d->layers[l].cout;
    }
    return 0;
}

int check_desc(const d2p_conv_desc* d) {
    D2P_REQUIRE(d != nullptr, "conv: null descriptor");
    D2P_REQUIRE(d->n_layers >= 1 && d->n_layers <= D2P_MAX_CONV_LAYERS, "conv: n_layers=%d", d->n_layers);
    D2P_REQUIRE(d->B > 0 && d->k > 0 && d->T > 0 && d->h > 0 && d->w > 0 && d->d > 0, "conv: bad dims");
    D2P_REQUIRE(d->d <= MAXC, "conv: input depth %d > %d", d->d, MAXC);
    D2P_REQUIRE(d->frames_dtype == D2P_U8 || d->frames_dtype == D2P_F32, "conv: frames dtype");
    for (int l = 0; l < d->n_layers; ++l) {
        int c = d->layers[l].cout;
        D2P_REQUIRE(c == 16 || c == 32 || c == 48, "conv: layer %d cout=%d unsupported (16/32/48)", l, c);
    }
    return 0;
}

}  // namespace

// ---- launch helpers for the register-blocked kernels (conv_v2.cuh) ------------------------
template <typename K>
static int cv_set_smem(K kernel, size_t bytes) {
    D2P_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

template <typename IN_T>
static int launch_conv_fwd(cudaStream_t st, const Geo& g, const IN_T* in, const float* sc, const float* sh,
                           const float* W, const float* b, float* out) {
    const long long npix = (long long)g.N * g.OH * g.OW;
    const int K3 = 3 * g.CIN;
    const size_t smem = ((size_t)K3 * CV_TP + (size_t)K3 * g.COUT) * sizeof(float);
    int blocks = cdiv(npix, CV_TP);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
#define D2P_CV_FWD(C)                                                                         \
    do {                                                                                      \
        D2P_TRY(cv_set_smem(conv_fwd_v2<C, IN_T>, smem));                                     \
        conv_fwd_v2<C, IN_T><<<blocks, CV_THREADS, smem, st>>>(g, in, sc, sh, W, b, out);     \
    } while (0)
    if (g.COUT == 16) D2P_CV_FWD(16);
    else if (g.COUT == 32) D2P_CV_FWD(32);
    else if (g.COUT == 48) D2P_CV_FWD(48);
    else return fail(D2P_ERR_ARG, "conv fwd: unsupported channel count %d (16/32/48)", g.COUT);
#undef D2P_CV_FWD
    D2P_CHECK_LAUNCH();
    return 0;
}

static int launch_conv_dx(cudaStream_t st, const Geo& g, const float* dZ, const float* W, float* dX) {
    const long long nmax = (long long)g.N * ((g.IH + 1) / 2) * ((g.IW + 1) / 2);
    const size_t smem = ((size_t)g.COUT * CV_TP + (size_t)g.COUT * g.CIN) * sizeof(float);
    int blocks = cdiv(nmax, CV_TP);
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
#define D2P_CV_DX(C)                                                                          \
    do {                                                                                      \
        D2P_TRY(cv_set_smem(conv_bwd_dx_v2<C>, smem));                                        \
        conv_bwd_dx_v2<C><<<dim3(blocks, 4), CV_THREADS, smem, st>>>(g, dZ, W, dX);           \
    } while (0)
    if (g.CIN == 16) D2P_CV_DX(16);
    else if (g.CIN == 32) D2P_CV_DX(32);
    else if (g.CIN == 48) D2P_CV_DX(48);
    else return fail(D2P_ERR_ARG, "conv bwd dx: unsupported channel count %d (16/32/48)", g.CIN);
#undef D2P_CV_DX
    D2P_CHECK_LAUNCH();
    return 0;
}

template <typename IN_T>
static int launch_conv_dw(cudaStream_t st, const Geo& g, const IN_T* in, const float* sc, const float* sh,
                          const float* dZ, int ppb, int nblk, float* partial) {
    if (g.CIN == 3 && sc == nullptr && g.COUT == 16) {   // RGB input layer: register-resident accumulators
        conv_bwd_dw_cin3<IN_T, 4><<<nblk, CV_THREADS, 0, st>>>(g, in, dZ, ppb, partial);
        D2P_CHECK_LAUNCH();
        return 0;
    }
    const int K3 = 3 * g.CIN;
    const size_t smem = ((size_t)K3 * (CV_TP + 4) + CV_TP + (size_t)CV_TP * g.COUT) * sizeof(float);
#define D2P_CV_DW(C)                                                                          \
    do {                                                                                      \
        D2P_TRY(cv_set_smem(conv_bwd_dw_v2<C, IN_T>, smem));                                  \
        conv_bwd_dw_v2<C, IN_T><<<dim3(nblk, 3), CV_THREADS, smem, st>>>(g, in, sc, sh, dZ, ppb, partial); \
    } while (0)
    if (g.COUT == 16) D2P_CV_DW(16);
    else if (g.COUT == 32) D2P_CV_DW(32);
    else if (g.COUT == 48) D2P_CV_DW(48);
    else return fail(D2P_ERR_ARG, "conv bwd dw: unsupported channel count %d (16/32/48)", g.COUT);
#undef D2P_CV_DW
    D2P_CHECK_LAUNCH();
    return 0;
}

// ---- saved-tensor layout ---------------------------------------------------
// saved = for each layer l: act a_l [N,OH,OW,C] | stats [4, k, C]
static size_t act_floats(const Geo& g) { return (size_t)g.N * g.OH * g.OW * g.COUT; }
static size_t stat_floats(const Geo& g) { return (size_t)4 * g.k * g.COUT; }

// scratch carve-up shared by fwd/bwd: dZ | dY | dy_tmp | coef | partials
struct Plan {
    Geo geo[D2P_MAX_CONV_LAYERS];
    size_t off_dz, off_dy, off_tmp, off_coef, off_part, part_bytes, off_tc, tc_bytes, total;
};

static void make_plan(const d2p_conv_desc* d, Plan* p) {
    size_t max_act = 0, max_in = 0, max_kc = 0, max_part = 0, max_tc = 0;
    for (int l = 0; l < d->n_layers; ++l) {
        Geo& g = p->geo[l];
        fill_geo(d, l, &g);
        if (l > 0 && conv_tc_supported(g) && conv_tc_ws_bytes(g) > max_tc) max_tc = conv_tc_ws_bytes(g);
        if (l == 0 && rgb_fwd_supported(d, g) && rgb_fwd_ws_bytes(g) > max_tc) max_tc = rgb_fwd_ws_bytes(g);
        if (l == 0 && d->frames_dtype == D2P_U8 && conv_tc_dw3_supported(g) && conv_tc_dw3_ws_bytes(g) > max_tc)
            max_tc = conv_tc_dw3_ws_bytes(g);
        size_t a = act_floats(g);
        size_t i = (size_t)g.N * g.IH * g.IW * g.CIN;
        if (a > max_act) max_act = a;
        if (l > 0 && i > max_in) max_in = i;
        if ((size_t)g.k * g.COUT > max_kc) max_kc = (size_t)g.k * g.COUT;
        long long rows = (long long)g.N * g.OH * g.OW;
        size_t bn = bn_ws_bytes(rows, g.COUT, g.k);
        size_t cs = bn_ws_bytes(rows, g.COUT, 1);
        int ppb; int nblk = dw_blocks(rows, &ppb);
        size_t dw = (size_t)nblk * 9 * g.CIN * g.COUT * sizeof(float);
        size_t m = bn > dw ? bn : dw; if (cs > m) m = cs;
        if (m > max_part) max_part = m;
    }
    if (d->n_layers == 3 && d->h == 8 && d->w == 8 && d->d == 16 && d->k >= 1 && d->k <= kNumSMs) {
        // exchange buffers of the fused single-kernel Karel paths (conv_fused.cu)
        if (conv_fused_ws_bytes(d) > max_part) max_part = conv_fused_ws_bytes(d);
        if (conv_fused_bwd_ws_bytes(d) > max_part) max_part = conv_fused_bwd_ws_bytes(d);
    }
    size_t dyf = max_act > max_in ? max_act : max_in;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    p->off_dz = 0;
    p->off_dy = al(max_act * sizeof(float));
    p->off_tmp = p->off_dy + al(dyf * sizeof(float));
    p->off_coef = p->off_tmp + al(max_act * sizeof(float));
    p->off_part = p->off_coef + al(2 * max_kc * sizeof(float));
    p->part_bytes = al(max_part);
    p->off_tc = p->off_part + p->part_bytes;
    p->tc_bytes = al(max_tc);
    p->total = p->off_tc + p->tc_bytes;
}

namespace {
__global__ void permute_frames_kernel(const float* __restrict__ X, float* __restrict__ Y, long long N,
                                      int F, int T, int R, int inverse) {
    long long total = N * F;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long n = idx / F; int f = (int)(idx % F);
        long long r = n / T; int t = (int)(n % T);
        long long m = (long long)t * R + r;
        if (!inverse) Y[m * F + f] = X[idx]; else Y[idx] = X[m * F + f];
    }
}
}  // namespace
// frame-major [R*T, F] -> time-major [T*R, F]
int permute_frames(cudaStream_t st, const float* X, float* Y, int N, int F, int T, int R) {
    long long total = (long long)N * F;
    int blocks = (int)((total + 255) / 256); if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    permute_frames_kernel<<<blocks, 256, 0, st>>>(X, Y, N, F, T, R, 0);
    D2P_CHECK_LAUNCH();
    return 0;
}
int unpermute_frames(cudaStream_t st, const float* X, float* Y, int N, int F, int T, int R) {
    long long total = (long long)N * F;
    int blocks = (int)((total + 255) / 256); if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    permute_frames_kernel<<<blocks, 256, 0, st>>>(X, Y, N, F, T, R, 1);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace d2p

using namespace d2p;

namespace d2p { void conv_set_rgb(int on) { g_conv_rgb = on; } }

extern "C" size_t d2p_conv_encoder_saved_floats(const d2p_conv_desc* d) {
    if (check_desc(d)) return 0;
    size_t n = 0;
    for (int l = 0; l < d->n_layers; ++l) { Geo g; fill_geo(d, l, &g); n += act_floats(g) + stat_floats(g); }
    return n;
}

extern "C" size_t d2p_conv_encoder_ws_bytes(const d2p_conv_desc* d) {
    if (check_desc(d)) return 0;
    Plan p; make_plan(d, &p);
    return p.total;
}

extern "C" int d2p_conv_encoder_feature_dim(const d2p_conv_desc* d) {
    if (check_desc(d)) return -1;
    Geo g; fill_geo(d, d->n_layers - 1, &g);
    return g.OH * g.OW * g.COUT;
}

extern "C" int d2p_conv_encoder_fwd(const d2p_conv_desc* d, const void* frames, float* feat,
                                    float* saved, int training, void* ws, size_t ws_bytes,
                                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_TRY(check_desc(d));
    D2P_REQUIRE(frames && feat && saved && ws, "conv fwd: null buffer");
    Plan p; make_plan(d, &p);
    D2P_REQUIRE(ws_bytes >= p.total, "conv fwd: workspace too small (%zu < %zu)", ws_bytes, p.total);
    char* wsb = (char*)ws;
    if (conv_fused_supported(d, training, p.part_bytes))
        return conv_fused_fwd(st, d, frames, feat, saved, training, wsb + p.off_part);
    const float* prev = nullptr; const float* prev_stats = nullptr;
    float* sp = saved;
    for (int l = 0; l < d->n_layers; ++l) {
        const Geo& g = p.geo[l];
        float* act = sp; float* stats = sp + act_floats(g);
        sp = stats + stat_floats(g);
        const d2p_conv_layer& L = d->layers[l];
        D2P_REQUIRE(L.w && L.b && L.gamma && L.beta && L.moving_mean && L.moving_var, "conv fwd: layer %d params", l);
        long long npix = (long long)g.N * g.OH * g.OW;
        int blocks = cdiv(npix, PX);
        const float* sc = prev_stats ? prev_stats + 2 * (size_t)g.k * g.CIN : nullptr;
        const float* sh = prev_stats ? prev_stats + 3 * (size_t)g.k * g.CIN : nullptr;
        (void)blocks;
        bool stats_done = false;
        if (l > 0 && (conv_tc_mode() & 1) && conv_tc_supported(g)) {
            // tensor-core implicit GEMM; the BatchNorm (sum, sum of squares) partials come from its epilogue
            int nchunk = 0; float2* partial = nullptr;
            D2P_TRY(conv_tc_fwd(st, g, prev, sc, sh, L.w, L.b, act, training, &nchunk, &partial, wsb + p.off_tc,
                                p.tc_bytes));
            if (training) {
                D2P_TRY(bn_forward_finalize(st, partial, nchunk, npix / g.k, g.COUT, g.k, L.gamma, L.beta,
                                            L.moving_mean, L.moving_var, stats));
                stats_done = true;
            }
        } else if (l == 0 && g_conv_rgb && rgb_fwd_supported(d, g)) {
            // RGB input layer: direct kernel with the BatchNorm partial sums fused
            const int cpd = rgb_cpd(g), nchunk = (g.N / g.T / g.k) * cpd;
            float2* partial = training ? (float2*)(wsb + p.off_tc) : nullptr;
            D2P_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_rgb_w, L.w, sizeof(float) * 27 * 16, 0, cudaMemcpyDeviceToDevice, st));
            D2P_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_rgb_b, L.b, sizeof(float) * 16, 0, cudaMemcpyDeviceToDevice, st));
            conv_rgb_fwd_kernel<<<(g.N / g.T) * cpd, RGB_THREADS, 0, st>>>(g, (const uint8_t*)frames, L.w, L.b, act,
                                                                         partial, cpd, nchunk);
            D2P_CHECK_LAUNCH();
            if (training) {
                D2P_TRY(bn_forward_finalize(st, partial, nchunk, npix / g.k, g.COUT, g.k, L.gamma, L.beta,
                                            L.moving_mean, L.moving_var, stats));
                stats_done = true;
            }
        } else if (l == 0 && d->frames_dtype == D2P_U8)
            D2P_TRY(launch_conv_fwd<uint8_t>(st, g, (const uint8_t*)frames, nullptr, nullptr, L.w, L.b, act));
        else
            D2P_TRY(launch_conv_fwd<float>(st, g, l == 0 ? (const float*)frames : prev, sc, sh, L.w, L.b, act));
        if (!stats_done)
            D2P_TRY(bn_forward_stats(st, act, npix, g.COUT, g.T * g.OH * g.OW, g.k, L.gamma, L.beta,
                                     L.moving_mean, L.moving_var, training, stats, wsb + p.off_part,
                                     p.part_bytes));
        prev = act; prev_stats = stats;
        if (l == d->n_layers - 1) {
            // feature = BN(a_L) flattened HWC, written time-major [T, R, F]
            int F = g.OH * g.OW * g.COUT;
            if (g.OH * g.OW == 1) {
                D2P_TRY(bn_apply(st, act, feat, g.N, g.COUT, g.T, g.k, stats, g.T, d->B * d->k));
            } else {
                float* tmp = (float*)(wsb + p.off_tmp);
                D2P_TRY(bn_apply(st, act, tmp, npix, g.COUT, g.T * g.OH * g.OW, g.k, stats, 0, 0));
                D2P_TRY(permute_frames(st, tmp, feat, g.N, F, g.T, d->B * d->k));
            }
        }
    }
    return 0;
}

// Backward: dfeat [T,R,F] -> parameter grads (accumulated into layers[l].dw etc.).
extern "C" int d2p_conv_encoder_bwd(const d2p_conv_desc* d, const void* frames, const float* dfeat,
                                    const float* saved, int training, void* ws, size_t ws_bytes,
                                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_TRY(check_desc(d));
    D2P_REQUIRE(frames && dfeat && saved && ws, "conv bwd: null buffer");
    Plan p; make_plan(d, &p);
    D2P_REQUIRE(ws_bytes >= p.total, "conv bwd: workspace too small (%zu < %zu)", ws_bytes, p.total);
    char* wsb = (char*)ws;
    if (conv_fused_bwd_supported(d, training, p.part_bytes))
        return conv_fused_bwd(st, d, frames, dfeat, saved, wsb + p.off_part);
    const float* acts[D2P_MAX_CONV_LAYERS]; const float* stats[D2P_MAX_CONV_LAYERS];
    const float* sp = saved;
    for (int l = 0; l < d->n_layers; ++l) {
        acts[l] = sp; stats[l] = sp + act_floats(p.geo[l]);
        sp = stats[l] + stat_floats(p.geo[l]);
    }
    float* dZ = (float*)(wsb + p.off_dz);
    float* dY = (float*)(wsb + p.off_dy);      // grad wrt the BN output feeding layer l+1
    float* dy_tmp = (float*)(wsb + p.off_tmp);
    float* coef = (float*)(wsb + p.off_coef);
    void* part = wsb + p.off_part;

    for (int l = d->n_layers - 1; l >= 0; --l) {
        const Geo& g = p.geo[l];
        const d2p_conv_layer& L = d->layers[l];
        D2P_REQUIRE(L.dw && L.db && L.dgamma && L.dbeta, "conv bwd: layer %d grads", l);
        long long npix = (long long)g.N * g.OH * g.OW;
        const float* dy_lin = dY;
        if (l == d->n_layers - 1) {
            // dfeat is time-major [T,R,F]; bring it back to frame order
            D2P_TRY(unpermute_frames(st, dfeat, dy_tmp, g.N, g.OH * g.OW * g.COUT, g.T, d->B * d->k));
            dy_lin = dy_tmp;
        }
        D2P_TRY(bn_backward_db(st, acts[l], dy_lin, dZ, npix, g.COUT, g.T * g.OH * g.OW, g.k, L.gamma,
                               stats[l], L.dgamma, L.dbeta, L.db, training, coef, part, p.part_bytes));
        int ppb; int nblk = dw_blocks(npix, &ppb);
        const float* sc = l > 0 ? stats[l - 1] + 2 * (size_t)g.k * g.CIN : nullptr;
        const float* sh = l > 0 ? stats[l - 1] + 3 * (size_t)g.k * g.CIN : nullptr;
        const bool tc_ok = l > 0 && conv_tc_supported(g);
        if (tc_ok && (conv_tc_mode() & 4)) {
            D2P_TRY(conv_tc_dw(st, g, acts[l - 1], sc, sh, dZ, L.dw, wsb + p.off_tc, p.tc_bytes));
        } else if (l == 0 && d->frames_dtype == D2P_U8 && (conv_tc_mode() & 4) && conv_tc_dw3_supported(g)) {
            D2P_TRY(conv_tc_dw3(st, g, (const uint8_t*)frames, dZ, L.dw, wsb + p.off_tc, p.tc_bytes));
        } else {
            if (l == 0 && d->frames_dtype == D2P_U8)
                D2P_TRY(launch_conv_dw<uint8_t>(st, g, (const uint8_t*)frames, nullptr, nullptr, dZ, ppb, nblk, (float*)part));
            else
                D2P_TRY(launch_conv_dw<float>(st, g, l == 0 ? (const float*)frames : acts[l - 1], sc, sh, dZ, ppb, nblk, (float*)part));
            int nw = 9 * g.CIN * g.COUT;
            conv_dw_reduce<<<cdiv(nw, 256), 256, 0, st>>>((const float*)part, nblk, nw, L.dw);
            D2P_CHECK_LAUNCH();
        }
        if (l > 0) {
            D2P_REQUIRE(g.CIN % 4 == 0, "conv bwd: CIN %% 4");
            if (tc_ok && (conv_tc_mode() & 2))
                D2P_TRY(conv_tc_dx(st, g, dZ, L.w, dY, wsb + p.off_tc, p.tc_bytes));
            else
                D2P_TRY(launch_conv_dx(st, g, dZ, L.w, dY));
        }
    }
    return 0;
}
