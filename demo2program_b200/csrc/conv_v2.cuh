// Register-blocked implicit-GEMM kernels of the State_Encoder convolutions
// (3x3, stride 2, TF SAME padding).  Included by conv.cu after `Geo`/`load_in`.
//
// Shared structure: a CTA of 128 threads owns a tile of 128 pixels; the loader role of
// thread t gathers everything pixel t needs for one kernel row (3 taps x CIN values, read
// as contiguous NHWC channel vectors, BatchNorm of the previous layer applied on the fly)
// into shared memory; the compute role of thread t owns 8 pixels x (C/8) channels, so one
// k-step costs 2 vector smem loads of activations + 1-2 of weights for 8*C/8 FMAs.
#pragma once

constexpr int CV_TP = 128;       // pixels per tile
constexpr int CV_THREADS = 128;

// CG (2, 4 or 6) consecutive weights of a slab row; CG and the row pitch are even, so 8-byte (16-byte
// for CG = 4) shared-memory loads are aligned - one or a few vector loads instead of CG scalar ones
template <int CG>
__device__ __forceinline__ void cv_load_w(const float* __restrict__ p, float* b) {
    if (CG == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    } else if (CG % 2 == 0) {
#pragma unroll
        for (int j = 0; j < CG; j += 2) {
            const float2 v = *reinterpret_cast<const float2*>(p + j);
            b[j] = v.x; b[j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < CG; ++j) b[j] = p[j];
    }
}

template <typename IN_T>
__device__ __forceinline__ void cv_load_channels(const IN_T* __restrict__ p, int cin, float* v);
template <>
__device__ __forceinline__ void cv_load_channels<float>(const float* __restrict__ p, int cin, float* v) {
    if ((cin & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        for (int c = 0; c < cin; c += 4) {
            float4 x = *reinterpret_cast<const float4*>(p + c);
            v[c] = x.x; v[c + 1] = x.y; v[c + 2] = x.z; v[c + 3] = x.w;
        }
    } else {
        for (int c = 0; c < cin; ++c) v[c] = p[c];
    }
}
template <>
__device__ __forceinline__ void cv_load_channels<uint8_t>(const uint8_t* __restrict__ p, int cin, float* v) {
    if ((cin & 15) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        for (int c = 0; c < cin; c += 16) {
            uint4 x = *reinterpret_cast<const uint4*>(p + c);
            uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int b = 0; b < 4; ++b) v[c + q * 4 + b] = (float)((w[q] >> (8 * b)) & 0xffu);
        }
    } else {
        for (int c = 0; c < cin; ++c) v[c] = (float)p[c];
    }
}

// ---- forward: a = lrelu(conv(x) + bias) -------------------------------------------------
// smem: As[3*CIN][CV_TP] | Ws[3*CIN][COUT]
template <int COUT, typename IN_T>
__global__ void __launch_bounds__(CV_THREADS)
conv_fwd_v2(Geo g, const IN_T* __restrict__ in, const float* __restrict__ in_scale,
            const float* __restrict__ in_shift, const float* __restrict__ W,
            const float* __restrict__ bias, float* __restrict__ out) {
    extern __shared__ __align__(16) float cv_smem[];
    constexpr int CG = COUT / 8;
    const int K3 = 3 * g.CIN;
    float* As = cv_smem;
    float* Ws = cv_smem + (size_t)K3 * CV_TP;
    const int tid = threadIdx.x, pg = tid >> 3, cg = tid & 7;
    const long long npix = (long long)g.N * g.OH * g.OW;
    const long long ntiles = (npix + CV_TP - 1) / CV_TP;
    float bv[CG];
#pragma unroll
    for (int j = 0; j < CG; ++j) bv[j] = bias[cg * CG + j];
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * CV_TP + tid;
        const bool pv = p < npix;
        int ox = 0, oy = 0, sl = 0;
        long long n = 0;
        if (pv) {
            ox = (int)(p % g.OW);
            oy = (int)((p / g.OW) % g.OH);
            n = p / ((long long)g.OW * g.OH);
            sl = (int)((n / g.T) % g.k);
        }
        float acc[8][CG];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < CG; ++j) acc[i][j] = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            // loader: 3 taps x CIN values of my pixel
            const int iy = 2 * oy + ky - g.PT;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox + kx - g.PL;
                const bool ok = pv && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                float v[MAXC];
                if (ok) {
                    cv_load_channels<IN_T>(in + (((size_t)n * g.IH + iy) * g.IW + ix) * g.CIN, g.CIN, v);
                    if (in_scale)
                        for (int c = 0; c < g.CIN; ++c)
                            v[c] = v[c] * in_scale[sl * g.CIN + c] + in_shift[sl * g.CIN + c];
                }
                for (int c = 0; c < g.CIN; ++c) As[(size_t)(kx * g.CIN + c) * CV_TP + tid] = ok ? v[c] : 0.f;
            }
            for (int idx = tid; idx < K3 * COUT; idx += CV_THREADS)
                Ws[idx] = W[(size_t)ky * K3 * COUT + idx];
            __syncthreads();
            for (int kk = 0; kk < K3; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(As + (size_t)kk * CV_TP + pg * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(As + (size_t)kk * CV_TP + pg * 8 + 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float b[CG];
                cv_load_w<CG>(Ws + kk * COUT + cg * CG, b);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < CG; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long q = tile * CV_TP + pg * 8 + i;
            if (q < npix) {
#pragma unroll
                for (int j = 0; j < CG; ++j) out[q * COUT + cg * CG + j] = lrelu_f(acc[i][j] + bv[j]);
            }
        }
    }
}

// ---- backward, input gradient ---------------------------------------------------------------
// dX[n,iy,ix,ci] = sum over taps with (iy+PT-ky, ix+PL-kx) even, co: dZ[n,oy,ox,co] W[ky,kx,ci,co].
// Input pixels are processed per parity class (blockIdx.y = (iy&1)*2 + (ix&1)): within a class every
// pixel uses the same 1, 2 or 4 taps, so no multiply-by-zero work is done.
// smem: As[COUT][CV_TP] | Ws[COUT][CIN]
template <int CIN>
__global__ void __launch_bounds__(CV_THREADS)
conv_bwd_dx_v2(Geo g, const float* __restrict__ dZ, const float* __restrict__ W, float* __restrict__ dX) {
    extern __shared__ __align__(16) float cv_smem[];
    constexpr int CG = CIN / 8;
    float* As = cv_smem;
    float* Ws = cv_smem + (size_t)g.COUT * CV_TP;
    const int tid = threadIdx.x, pg = tid >> 3, cg = tid & 7;
    const int py = blockIdx.y >> 1, px = blockIdx.y & 1;
    const int HY = (g.IH - py + 1) / 2, WX = (g.IW - px + 1) / 2;   // rows / cols of this class
    if (HY <= 0 || WX <= 0) return;
    const long long npix = (long long)g.N * HY * WX;
    const long long ntiles = (npix + CV_TP - 1) / CV_TP;
    const int ky0 = (py + g.PT) & 1, kx0 = (px + g.PL) & 1;         // first valid tap, step 2
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * CV_TP + tid;
        const bool pv = p < npix;
        int ix = 0, iy = 0;
        long long n = 0;
        if (pv) {
            ix = (int)(p % WX) * 2 + px;
            iy = (int)((p / WX) % HY) * 2 + py;
            n = p / ((long long)WX * HY);
        }
        float acc[8][CG];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < CG; ++j) acc[i][j] = 0.f;
        for (int ky = ky0; ky < 3; ky += 2)
            for (int kx = kx0; kx < 3; kx += 2) {
                const int ty = iy + g.PT - ky, tx = ix + g.PL - kx;
                const int oy = ty >> 1, ox = tx >> 1;
                const bool ok = pv && ty >= 0 && tx >= 0 && oy < g.OH && ox < g.OW;
                float v[MAXC];
                if (ok) cv_load_channels<float>(dZ + (((size_t)n * g.OH + oy) * g.OW + ox) * g.COUT, g.COUT, v);
                for (int c = 0; c < g.COUT; ++c) As[(size_t)c * CV_TP + tid] = ok ? v[c] : 0.f;
                // Ws[co][ci] = W[ky,kx,ci,co]
                for (int idx = tid; idx < CIN * g.COUT; idx += CV_THREADS) {
                    const int ci = idx / g.COUT, co = idx % g.COUT;
                    Ws[co * CIN + ci] = W[((size_t)(ky * 3 + kx) * CIN + ci) * g.COUT + co];
                }
                __syncthreads();
                for (int kk = 0; kk < g.COUT; ++kk) {
                    const float4 a0 = *reinterpret_cast<const float4*>(As + (size_t)kk * CV_TP + pg * 8);
                    const float4 a1 = *reinterpret_cast<const float4*>(As + (size_t)kk * CV_TP + pg * 8 + 4);
                    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    float b[CG];
                    cv_load_w<CG>(Ws + kk * CIN + cg * CG, b);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < CG; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
                __syncthreads();
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long q = tile * CV_TP + pg * 8 + i;
            if (q < npix) {
                const int qx = (int)(q % WX) * 2 + px, qy = (int)((q / WX) % HY) * 2 + py;
                const long long qn = q / ((long long)WX * HY);
                float* dst = dX + (((size_t)qn * g.IH + qy) * g.IW + qx) * CIN + cg * CG;
#pragma unroll
                for (int j = 0; j < CG; ++j) dst[j] = acc[i][j];
            }
        }
    }
}

// ---- backward, weight gradient -----------------------------------------------------------------
// partial[(blk*3 + ky) * 3*CIN*COUT + (kx*CIN+ci)*COUT + co] = sum over the block's pixels of
// x(pixel @ tap (ky,kx), ci) * dZ(pixel, co).  grid = (nblk, 3).  A thread owns 4 output channels
// and every RL-th row of the [3*CIN, COUT] slab.
// smem: As[CV_TP][3*CIN + 1] | Zs[CV_TP][COUT]
template <int COUT, typename IN_T>
__global__ void __launch_bounds__(CV_THREADS)
conv_bwd_dw_v2(Geo g, const IN_T* __restrict__ in, const float* __restrict__ in_scale,
               const float* __restrict__ in_shift, const float* __restrict__ dZ, int pix_per_block,
               float* __restrict__ partial) {
    extern __shared__ __align__(16) float cv_smem[];
    constexpr int C4 = COUT / 4;                 // channel quads
    constexpr int RL = CV_THREADS / C4;          // row lanes
    constexpr int MAXROWS = (3 * MAXC + RL - 1) / RL;
    const int K3 = 3 * g.CIN, AST = K3 + 1;
    constexpr int APT = CV_TP + 4;               // pitch of the transposed slab [row][pixel]
    float* As = cv_smem;
    float* Zs = cv_smem + (size_t)K3 * APT + CV_TP;
    const int tid = threadIdx.x, ky = blockIdx.y;
    const int c4 = (tid % C4) * 4, rl = tid / C4;
    // few slab rows (3*CIN < RL, e.g. the RGB input layer): the spare row lanes split the
    // tile's pixels into PG groups that are summed through shared memory at the end
    const int PG = K3 < RL ? RL / K3 : 1;
    const int row0 = PG > 1 ? rl % K3 : rl;
    const int pgp = PG > 1 ? rl / K3 : 0;
    const bool active = pgp < PG;
    const long long npix = (long long)g.N * g.OH * g.OW;
    const long long p0 = (long long)blockIdx.x * pix_per_block;
    long long p1 = p0 + pix_per_block;
    if (p1 > npix) p1 = npix;
    float acc[MAXROWS][4];
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    for (long long pb = p0; pb < p1; pb += CV_TP) {
        const long long p = pb + tid;
        const bool pv = p < p1;
        int ox = 0, oy = 0, sl = 0;
        long long n = 0;
        if (pv) {
            ox = (int)(p % g.OW);
            oy = (int)((p / g.OW) % g.OH);
            n = p / ((long long)g.OW * g.OH);
            sl = (int)((n / g.T) % g.k);
        }
        const int iy = 2 * oy + ky - g.PT;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * ox + kx - g.PL;
            const bool ok = pv && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
            float v[MAXC];
            if (ok) {
                cv_load_channels<IN_T>(in + (((size_t)n * g.IH + iy) * g.IW + ix) * g.CIN, g.CIN, v);
                if (in_scale)
                    for (int c = 0; c < g.CIN; ++c)
                        v[c] = v[c] * in_scale[sl * g.CIN + c] + in_shift[sl * g.CIN + c];
            }
            if (PG == 1)   // transposed slab: a 16-byte load covers 4 pixels of one row
                for (int c = 0; c < g.CIN; ++c) As[(size_t)(kx * g.CIN + c) * APT + tid] = ok ? v[c] : 0.f;
            else
                for (int c = 0; c < g.CIN; ++c) As[(size_t)tid * AST + kx * g.CIN + c] = ok ? v[c] : 0.f;
        }
        {
            float z[MAXC];
            if (pv) cv_load_channels<float>(dZ + (size_t)p * COUT, COUT, z);
            for (int c = 0; c < COUT; ++c) Zs[tid * COUT + c] = pv ? z[c] : 0.f;
        }
        __syncthreads();
        if (PG == 1) {
            // 4 pixels per step: 4 dZ quads + one slab vector per row feed 16 FMAs per row
            for (int pp = 0; pp < CV_TP; pp += 4) {
                float4 z[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) z[e] = *reinterpret_cast<const float4*>(Zs + (pp + e) * COUT + c4);
#pragma unroll
                for (int r = 0; r < MAXROWS; ++r) {
                    const int row = row0 + r * RL;
                    if (row < K3) {
                        const float4 a = *reinterpret_cast<const float4*>(As + (size_t)row * APT + pp);
                        acc[r][0] = fmaf(a.x, z[0].x, acc[r][0]); acc[r][1] = fmaf(a.x, z[0].y, acc[r][1]);
                        acc[r][2] = fmaf(a.x, z[0].z, acc[r][2]); acc[r][3] = fmaf(a.x, z[0].w, acc[r][3]);
                        acc[r][0] = fmaf(a.y, z[1].x, acc[r][0]); acc[r][1] = fmaf(a.y, z[1].y, acc[r][1]);
                        acc[r][2] = fmaf(a.y, z[1].z, acc[r][2]); acc[r][3] = fmaf(a.y, z[1].w, acc[r][3]);
                        acc[r][0] = fmaf(a.z, z[2].x, acc[r][0]); acc[r][1] = fmaf(a.z, z[2].y, acc[r][1]);
                        acc[r][2] = fmaf(a.z, z[2].z, acc[r][2]); acc[r][3] = fmaf(a.z, z[2].w, acc[r][3]);
                        acc[r][0] = fmaf(a.w, z[3].x, acc[r][0]); acc[r][1] = fmaf(a.w, z[3].y, acc[r][1]);
                        acc[r][2] = fmaf(a.w, z[3].z, acc[r][2]); acc[r][3] = fmaf(a.w, z[3].w, acc[r][3]);
                    }
                }
            }
        } else if (active) {
            for (int pp = pgp; pp < CV_TP; pp += PG) {
                const float4 z4 = *reinterpret_cast<const float4*>(Zs + pp * COUT + c4);
#pragma unroll
                for (int r = 0; r < MAXROWS; ++r) {
                    const int row = row0 + r * RL;
                    if (row < K3) {
                        const float a = As[(size_t)pp * AST + row];
                        acc[r][0] = fmaf(a, z4.x, acc[r][0]); acc[r][1] = fmaf(a, z4.y, acc[r][1]);
                        acc[r][2] = fmaf(a, z4.z, acc[r][2]); acc[r][3] = fmaf(a, z4.w, acc[r][3]);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (PG > 1) {   // combine the pixel groups (the last tile's trailing barrier makes As reusable)
        float* red = As;   // [PG][K3][COUT] <= RL*COUT floats <= CV_TP*AST
        if (active)
            *reinterpret_cast<float4*>(red + ((size_t)pgp * K3 + row0) * COUT + c4) =
                make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
        __syncthreads();
        if (active && pgp == 0)
            for (int p2 = 1; p2 < PG; ++p2) {
                const float4 o = *reinterpret_cast<const float4*>(red + ((size_t)p2 * K3 + row0) * COUT + c4);
                acc[0][0] += o.x; acc[0][1] += o.y; acc[0][2] += o.z; acc[0][3] += o.w;
            }
    }
    if (active && pgp == 0) {
        float* dst = partial + ((size_t)blockIdx.x * 3 + ky) * K3 * COUT;
#pragma unroll
        for (int r = 0; r < MAXROWS; ++r) {
            const int row = row0 + r * RL;
            if (row < K3)
                *reinterpret_cast<float4*>(dst + (size_t)row * COUT + c4) =
                    make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    }
}


// ---- backward, weight gradient of a 3-channel input layer (ViZDoom RGB frames) ---------------------
// 9*3 = 27 slab rows: far too few for the row-lane scheme above (one FMA per two shared-memory
// loads).  Here a thread keeps ALL 27 x 4 accumulators of one output-channel quad in registers and
// streams its own pixels: 27 input bytes (all loads of a pixel issued before the first FMA; the 3x3
// windows overlap, so they are L1 hits) + one 16-byte dZ load feed 108 FMAs.  The COUT/4 quads of a
// pixel sit in adjacent lanes: they share the input loads (one transaction) and read one contiguous
// dZ row.  grid = nblk; the block's threads are combined with a fixed shuffle / shared-memory tree
// (deterministic) into the same partial layout as conv_bwd_dw_v2.
template <typename IN_T, int NQ /* COUT / 4 */>
__global__ void __launch_bounds__(CV_THREADS)
conv_bwd_dw_cin3(Geo g, const IN_T* __restrict__ in, const float* __restrict__ dZ, int pix_per_block,
                 float* __restrict__ partial) {
    constexpr int PPI = CV_THREADS / NQ;        // pixels per block iteration
    __shared__ float red[CV_THREADS / 32][NQ][108];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quad = tid % NQ, c4 = quad * 4;
    const long long npix = (long long)g.N * g.OH * g.OW;
    const long long p0 = (long long)blockIdx.x * pix_per_block;
    long long p1 = p0 + pix_per_block;
    if (p1 > npix) p1 = npix;
    float acc[27][4];
#pragma unroll
    for (int r = 0; r < 27; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    for (long long p = p0 + tid / NQ; p < p1; p += PPI) {
        const int ox = (int)(p % g.OW), oy = (int)((p / g.OW) % g.OH);
        const long long n = p / ((long long)g.OW * g.OH);
        const float4 z = *reinterpret_cast<const float4*>(dZ + (size_t)p * g.COUT + c4);
        const IN_T* base = in + (size_t)n * g.IH * g.IW * 3;
        float x[27];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy + ky - g.PT;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox + kx - g.PL;
                const bool ok = iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                const IN_T* px = base + ((size_t)(ok ? iy : 0) * g.IW + (ok ? ix : 0)) * 3;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = (float)__ldg(px + ci);
                    x[(ky * 3 + kx) * 3 + ci] = ok ? v : 0.f;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 27; ++r) {
            acc[r][0] = fmaf(x[r], z.x, acc[r][0]); acc[r][1] = fmaf(x[r], z.y, acc[r][1]);
            acc[r][2] = fmaf(x[r], z.z, acc[r][2]); acc[r][3] = fmaf(x[r], z.w, acc[r][3]);
        }
    }
    // lanes with the same quad: strides NQ, 2 NQ, ... within the warp
#pragma unroll
    for (int r = 0; r < 27; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[r][j];
#pragma unroll
            for (int o = 16; o >= NQ; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane < NQ) red[warp][lane][r * 4 + j] = v;
        }
    __syncthreads();
    for (int idx = tid; idx < NQ * 108; idx += CV_THREADS) {
        const int qd = idx / 108, e = idx - qd * 108;
        float s_ = 0.f;
#pragma unroll
        for (int w = 0; w < CV_THREADS / 32; ++w) s_ += red[w][qd][e];
        // partial[blk][ky][kx*CIN+ci][co]: 27 rows of COUT
        partial[((size_t)blockIdx.x * 27 + e / 4) * g.COUT + qd * 4 + (e & 3)] = s_;
    }
}
