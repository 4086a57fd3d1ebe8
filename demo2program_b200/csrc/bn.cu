// Per-slice BatchNorm machinery over [rows, C] matrices.
//
// Restates tf.contrib.layers.batch_norm(decay=0.9, eps=1e-3, center, scale,
// updates_collections=None) as called at reference models/ops.py:20-23.
// The reference instantiates the encoder k times with shared weights, so batch
// statistics are per demonstration index ("slice"); slice(row) =
// (row / seg) % nsl lets one launch cover all k copies for every layout used on
// the path (frame-major conv activations: seg = T*OH*OW; time-major [T,R,C]
// rows: seg = 1; RN-pool rows: nsl = 1).
//
// All reductions are two-stage (per-block partials in a caller-provided
// workspace, then a fixed-order finalize) so results are run-to-run
// deterministic.
#include "common.cuh"

namespace d2p {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ long long slice_row(long long q, int seg, int nsl, int sl) {
    // q: slice-local row index -> global row
    return ((q / seg) * nsl + sl) * (long long)seg + (q % seg);
}

// partial[(sl * nchunk + chunk) * C + c] = {sum a, sum b} where
//   MODE 0: a = x,  b = x*x           (forward statistics)
//   MODE 1: a = dy, b = dy * xhat     (backward; xhat from mean/rstd)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
colstats_partial(const float* __restrict__ X, const float* __restrict__ DY,
                 const float* __restrict__ mean, const float* __restrict__ rstd,
                 long long rows_per_slice, int C, int seg, int nsl, int rows_per_chunk,
                 float2* __restrict__ partial) {
    extern __shared__ float2 red[];  // [RL][CT]
    const int chunk = blockIdx.x, sl = blockIdx.y, nchunk = gridDim.x;
    const int CT = C < kThreads ? C : kThreads;
    const int RL = kThreads / CT;
    const int rl = threadIdx.x / CT, ct = threadIdx.x % CT;
    const bool active = rl < RL;
    long long q0 = (long long)chunk * rows_per_chunk;
    long long q1 = q0 + rows_per_chunk;
    if (q1 > rows_per_slice) q1 = rows_per_slice;
    {   // one column tile per blockIdx.z
        const int c0 = blockIdx.z * CT;
        const int c = c0 + ct;
        const bool valid = active && c < C;
        float sa = 0.f, sb = 0.f;
        if (valid) {
            float mu = 0.f, rs = 0.f;
            if (MODE == 1) { mu = mean[sl * C + c]; rs = rstd[sl * C + c]; }
            for (long long q = q0 + rl; q < q1; q += RL) {
                long long row = slice_row(q, seg, nsl, sl);
                float x = X[row * C + c];
                if (MODE == 0) { sa += x; sb += x * x; }
                else { float dy = DY[row * C + c]; sa += dy; sb += dy * (x - mu) * rs; }
            }
        }
        if (active) red[rl * CT + ct] = make_float2(sa, sb);
        __syncthreads();
        if (threadIdx.x < CT && c < C) {
            float2 acc = red[ct];
            for (int r = 1; r < RL; ++r) { acc.x += red[r * CT + ct].x; acc.y += red[r * CT + ct].y; }
            partial[((size_t)sl * nchunk + chunk) * C + c] = acc;
        }
        __syncthreads();
    }
}

// Same reduction for C % 4 == 0 and 16-byte aligned rows: a thread owns a channel quad,
// 16-byte loads, four rows in flight (the large ViZDoom activations are HBM-bound here).
// smem: float red[RL][CT4][8]
template <int MODE>
__global__ void __launch_bounds__(kThreads)
colstats_partial_v4(const float* __restrict__ X, const float* __restrict__ DY,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    long long rows_per_slice, int C, int seg, int nsl, int rows_per_chunk,
                    float2* __restrict__ partial) {
    extern __shared__ float red4[];
    const int chunk = blockIdx.x, sl = blockIdx.y, nchunk = gridDim.x;
    const int C4 = C / 4;
    const int CT = C4 < kThreads ? C4 : kThreads;
    const int RL = kThreads / CT;
    const int rl = threadIdx.x / CT, ct = threadIdx.x % CT;
    const int cq = blockIdx.z * CT + ct;
    const bool valid = rl < RL && cq < C4;
    long long q0 = (long long)chunk * rows_per_chunk;
    long long q1 = q0 + rows_per_chunk;
    if (q1 > rows_per_slice) q1 = rows_per_slice;
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
        float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu;
        if (MODE == 1) {
            mu = *reinterpret_cast<const float4*>(mean + (size_t)sl * C + cq * 4);
            rs = *reinterpret_cast<const float4*>(rstd + (size_t)sl * C + cq * 4);
        }
        for (long long q = q0 + rl; q < q1; q += 4LL * RL) {
            float4 x[4], dy[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long qe = q + (long long)e * RL;
                if (qe < q1) {
                    const size_t off = (size_t)slice_row(qe, seg, nsl, sl) * C + cq * 4;
                    x[e] = *reinterpret_cast<const float4*>(X + off);
                    if (MODE == 1) dy[e] = *reinterpret_cast<const float4*>(DY + off);
                } else {
                    x[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MODE == 1) { dy[e] = x[e]; x[e] = mu; }
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (MODE == 0) {
                    sa[0] += x[e].x; sa[1] += x[e].y; sa[2] += x[e].z; sa[3] += x[e].w;
                    sb[0] += x[e].x * x[e].x; sb[1] += x[e].y * x[e].y;
                    sb[2] += x[e].z * x[e].z; sb[3] += x[e].w * x[e].w;
                } else {
                    sa[0] += dy[e].x; sa[1] += dy[e].y; sa[2] += dy[e].z; sa[3] += dy[e].w;
                    sb[0] += dy[e].x * (x[e].x - mu.x) * rs.x; sb[1] += dy[e].y * (x[e].y - mu.y) * rs.y;
                    sb[2] += dy[e].z * (x[e].z - mu.z) * rs.z; sb[3] += dy[e].w * (x[e].w - mu.w) * rs.w;
                }
            }
        }
    }
    if (rl < RL) {
        float* d = red4 + ((size_t)rl * CT + ct) * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) { d[j] = sa[j]; d[4 + j] = sb[j]; }
    }
    __syncthreads();
    // thread (ct, j): channel cq*4 + j, summed over the row lanes in order
    for (int w = threadIdx.x; w < CT * 4; w += kThreads) {
        const int ct2 = w / 4, j = w % 4;
        const int cq2 = blockIdx.z * CT + ct2;
        if (cq2 < C4) {
            float a = 0.f, b = 0.f;
            for (int r = 0; r < RL; ++r) {
                const float* d = red4 + ((size_t)r * CT + ct2) * 8;
                a += d[j]; b += d[4 + j];
            }
            partial[((size_t)sl * nchunk + chunk) * C + cq2 * 4 + j] = make_float2(a, b);
        }
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Finalize kernels: one WARP per channel (8 channels per 256-thread block); lanes
// stride over the per-block partials and combine with a fixed shuffle tree, so the
// result is deterministic.
// Forward: slices in order (the reference updates the moving stats once per reuse
// call, i = 0..k-1).
__global__ void bn_fwd_finalize(const float2* __restrict__ partial, int nchunk, int C, int nsl,
                                double count, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ moving_mean,
                                float* __restrict__ moving_var, float eps, float decay,
                                int training, float* __restrict__ mean_out,
                                float* __restrict__ rstd_out, float* __restrict__ scale,
                                float* __restrict__ shift) {
    const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (c >= C) return;
    float g = gamma[c], b = beta[c];
    if (!training) {
        float mu = moving_mean[c], rs = rsqrtf(moving_var[c] + eps);
        for (int sl = lane; sl < nsl; sl += 32) {
            mean_out[sl * C + c] = mu; rstd_out[sl * C + c] = rs;
            scale[sl * C + c] = g * rs; shift[sl * C + c] = b - mu * g * rs;
        }
        return;
    }
    float mm = moving_mean[c], mv = moving_var[c];
    for (int sl = 0; sl < nsl; ++sl) {
        double s = 0.0, s2 = 0.0;
        for (int j = lane; j < nchunk; j += 32) {
            float2 p = partial[((size_t)sl * nchunk + j) * C + c];
            s += p.x; s2 += p.y;
        }
        s = warp_sum_d(s); s2 = warp_sum_d(s2);
        double mu = s / count;
        double var = s2 / count - mu * mu;
        if (var < 0.0) var = 0.0;
        float muf = (float)mu, varf = (float)var;
        float rs = (float)(1.0 / sqrt(var + (double)eps));
        if (lane == 0) {
            mean_out[sl * C + c] = muf; rstd_out[sl * C + c] = rs;
            scale[sl * C + c] = g * rs; shift[sl * C + c] = b - muf * g * rs;
        }
        mm -= (mm - muf) * (1.f - decay);
        mv -= (mv - varf) * (1.f - decay);
    }
    if (lane == 0) { moving_mean[c] = mm; moving_var[c] = mv; }
}

// Backward finalize: coefficients for dx and the shared-parameter grads.
//   dgamma[c] += sum_sl sum dy*xhat ; dbeta[c] += sum_sl sum dy
//   k1[sl,c] = sum(dy)/M ; k2[sl,c] = sum(dy*xhat)/M
__global__ void bn_bwd_finalize(const float2* __restrict__ partial, int nchunk, int C, int nsl,
                                double count, float* __restrict__ dgamma,
                                float* __restrict__ dbeta, float* __restrict__ k1,
                                float* __restrict__ k2, int training) {
    const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (c >= C) return;
    double tg = 0.0, tb = 0.0;
    for (int sl = 0; sl < nsl; ++sl) {
        double s = 0.0, s2 = 0.0;
        for (int j = lane; j < nchunk; j += 32) {
            float2 p = partial[((size_t)sl * nchunk + j) * C + c];
            s += p.x; s2 += p.y;
        }
        s = warp_sum_d(s); s2 = warp_sum_d(s2);
        tb += s; tg += s2;
        if (lane == 0) {
            k1[sl * C + c] = training ? (float)(s / count) : 0.f;
            k2[sl * C + c] = training ? (float)(s2 / count) : 0.f;
        }
    }
    if (lane == 0) { dgamma[c] += (float)tg; dbeta[c] += (float)tb; }
}

// y = x*scale + shift, optionally permuting rows (r*T + t) -> (t*R + r)
__global__ void bn_apply_kernel(const float* __restrict__ X, float* __restrict__ Y, long long rows,
                                int C, int seg, int nsl, const float* __restrict__ scale,
                                const float* __restrict__ shift, int permT, int permR) {
    long long total = rows * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long row = idx / C;
        int c = (int)(idx % C);
        int sl = (int)((row / seg) % nsl);
        float y = X[idx] * scale[sl * C + c] + shift[sl * C + c];
        long long orow = row;
        if (permT > 0) { long long r = row / permT; int t = (int)(row % permT); orow = (long long)t * permR + r; }
        Y[orow * C + c] = y;
    }
}

// one slice, no row permutation, C % 4 == 0, 16-byte aligned: float4, 32-bit index math
__global__ void bn_apply_v4_kernel(const float4* __restrict__ X, float4* __restrict__ Y, unsigned total4,
                                   unsigned C4, const float* __restrict__ scale,
                                   const float* __restrict__ shift) {
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += gridDim.x * blockDim.x) {
        const unsigned c = (idx % C4) * 4;
        const float4 x = X[idx];
        const float4 sc = *reinterpret_cast<const float4*>(scale + c);
        const float4 sh = *reinterpret_cast<const float4*>(shift + c);
        Y[idx] = make_float4(x.x * sc.x + sh.x, x.y * sc.y + sh.y, x.z * sc.z + sh.z, x.w * sc.w + sh.w);
    }
}

// da = gamma*rstd*(dy - k1 - xhat*k2); if ACT: dz = da * lrelu'(a) (a = post-activation
// value that was normalised).  DY may be row-permuted like bn_apply's output.
template <bool ACT>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ A, const float* __restrict__ DY,
                                    float* __restrict__ DZ, long long rows, int C, int seg, int nsl,
                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ k1,
                                    const float* __restrict__ k2, int permT, int permR) {
    long long total = rows * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long row = idx / C;
        int c = (int)(idx % C);
        int sl = (int)((row / seg) % nsl);
        long long drow = row;
        if (permT > 0) { long long r = row / permT; int t = (int)(row % permT); drow = (long long)t * permR + r; }
        float a = A[idx];
        float rs = rstd[sl * C + c];
        float xh = (a - mean[sl * C + c]) * rs;
        float da = gamma[c] * rs * (DY[drow * C + c] - k1[sl * C + c] - xh * k2[sl * C + c]);
        DZ[idx] = ACT ? da * lrelu_grad_from_out(a) : da;
    }
}

// 16-byte variant for the contiguous (unpermuted) case, C % 4 == 0, rows*C/4 < 2^32
template <bool ACT>
__global__ void bn_bwd_apply_v4(const float4* __restrict__ A, const float4* __restrict__ DY,
                                float4* __restrict__ DZ, unsigned total4, unsigned C4, unsigned seg,
                                unsigned nsl, const float4* __restrict__ gamma, const float4* __restrict__ mean,
                                const float4* __restrict__ rstd, const float4* __restrict__ k1,
                                const float4* __restrict__ k2) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total4; i0 += 2 * stride) {
        float4 a[2], dy[2];
        unsigned idx[2] = {i0, i0 + stride};
#pragma unroll
        for (int e = 0; e < 2; ++e)
            if (idx[e] < total4) { a[e] = A[idx[e]]; dy[e] = DY[idx[e]]; }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (idx[e] >= total4) continue;
            const unsigned row = idx[e] / C4, cq = idx[e] - row * C4;
            const unsigned p = ((row / seg) % nsl) * C4 + cq;
            const float4 g = gamma[cq], rs = rstd[p], mu = mean[p], q1 = k1[p], q2 = k2[p];
            float4 o;
            o.x = g.x * rs.x * (dy[e].x - q1.x - (a[e].x - mu.x) * rs.x * q2.x);
            o.y = g.y * rs.y * (dy[e].y - q1.y - (a[e].y - mu.y) * rs.y * q2.y);
            o.z = g.z * rs.z * (dy[e].z - q1.z - (a[e].z - mu.z) * rs.z * q2.z);
            o.w = g.w * rs.w * (dy[e].w - q1.w - (a[e].w - mu.w) * rs.w * q2.w);
            if (ACT) {
                o.x *= lrelu_grad_from_out(a[e].x); o.y *= lrelu_grad_from_out(a[e].y);
                o.z *= lrelu_grad_from_out(a[e].z); o.w *= lrelu_grad_from_out(a[e].w);
            }
            DZ[idx[e]] = o;
        }
    }
}

// bn_bwd_apply_v4<true> that also leaves the column sums of the dZ it writes (the bias gradient of the layer in
// front of this BatchNorm) as per-block partials: thread (rl, ct) owns channel quad ct and every RL-th row of
// the block's chunk, so its four running sums belong to fixed channels; combined over rl in a fixed order.
// smem: float red[RL][CT][4]
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_db_v4(const float4* __restrict__ A, const float4* __restrict__ DY, float4* __restrict__ DZ,
                   long long rows, int C4, int seg, int nsl, int rows_per_chunk,
                   const float4* __restrict__ gamma, const float4* __restrict__ mean,
                   const float4* __restrict__ rstd, const float4* __restrict__ k1,
                   const float4* __restrict__ k2, float2* __restrict__ partial) {
    extern __shared__ float red4[];
    const int CT = C4, RL = kThreads / CT;
    const int rl = threadIdx.x / CT, ct = threadIdx.x % CT;
    const long long q0 = (long long)blockIdx.x * rows_per_chunk;
    long long q1 = q0 + rows_per_chunk;
    if (q1 > rows) q1 = rows;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (rl < RL) {
        const float4 g = gamma[ct];
        for (long long row = q0 + rl; row < q1; row += RL) {
            const size_t idx = (size_t)row * C4 + ct;
            const float4 a = A[idx], dy = DY[idx];
            const int p = (int)((row / seg) % nsl) * C4 + ct;
            const float4 rs = rstd[p], mu = mean[p], c1 = k1[p], c2 = k2[p];
            float4 o;
            o.x = g.x * rs.x * (dy.x - c1.x - (a.x - mu.x) * rs.x * c2.x) * lrelu_grad_from_out(a.x);
            o.y = g.y * rs.y * (dy.y - c1.y - (a.y - mu.y) * rs.y * c2.y) * lrelu_grad_from_out(a.y);
            o.z = g.z * rs.z * (dy.z - c1.z - (a.z - mu.z) * rs.z * c2.z) * lrelu_grad_from_out(a.z);
            o.w = g.w * rs.w * (dy.w - c1.w - (a.w - mu.w) * rs.w * c2.w) * lrelu_grad_from_out(a.w);
            DZ[idx] = o;
            s[0] += o.x; s[1] += o.y; s[2] += o.z; s[3] += o.w;
        }
        float* d = red4 + ((size_t)rl * CT + ct) * 4;
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    }
    __syncthreads();
    for (int w = threadIdx.x; w < CT * 4; w += kThreads) {
        float acc = 0.f;
        for (int r = 0; r < RL; ++r) acc += red4[((size_t)r * CT + w / 4) * 4 + (w & 3)];
        partial[(size_t)blockIdx.x * (CT * 4) + w] = make_float2(acc, 0.f);
    }
}

__global__ void colsum_finalize(const float2* __restrict__ partial, int nchunk, int C,
                                float* __restrict__ out, float beta) {
    const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (c >= C) return;
    double s = 0.0;
    for (int j = lane; j < nchunk; j += 32) s += partial[(size_t)j * C + c].x;
    s = warp_sum_d(s);
    if (lane == 0) out[c] = (beta != 0.f ? beta * out[c] : 0.f) + (float)s;
}

// Gather rows of a slice-permuted matrix: used when DY is permuted but stats
// must be read in A's row order.
__global__ void permute_rows_kernel(const float* __restrict__ X, float* __restrict__ Y,
                                    long long rows, int C, int permT, int permR) {
    long long total = rows * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long row = idx / C;
        int c = (int)(idx % C);
        long long r = row / permT; int t = (int)(row % permT);
        Y[idx] = X[((long long)t * permR + r) * C + c];
    }
}

inline int pick_chunks(long long rows_per_slice, int nsl, int* rows_per_chunk) {
    // aim for ~4 waves of blocks overall, at least 16 rows per block (a 3200-row activation of the
    // summary pools was 50 blocks on 148 SMs with 64: latency-bound on a third of the GPU)
    long long target_blocks = 4LL * kNumSMs / (nsl > 0 ? nsl : 1);
    if (target_blocks < 1) target_blocks = 1;
    long long rpc = (rows_per_slice + target_blocks - 1) / target_blocks;
    if (rpc < 16) rpc = 16;
    *rows_per_chunk = (int)rpc;
    return (int)((rows_per_slice + rpc - 1) / rpc);
}

inline int ew_blocks(long long total) {
    long long b = (total + 255) / 256;
    long long cap = 8LL * kNumSMs;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// launches the column-statistics reduction (vectorised when the layout allows)
template <int MODE>
int launch_colstats(cudaStream_t st, const float* X, const float* DY, const float* mean, const float* rstd,
                    long long rows_per_slice, int C, int seg, int nsl, int nchunk, int rpc, float2* ws) {
    if (C % 4 == 0 && al16(X) && (MODE == 0 || (al16(DY) && al16(mean) && al16(rstd)))) {
        const int C4 = C / 4, CT = C4 < kThreads ? C4 : kThreads;
        const size_t sm = (size_t)(kThreads / CT) * CT * 8 * sizeof(float);
        colstats_partial_v4<MODE><<<dim3(nchunk, nsl, cdiv(C4, CT)), kThreads, sm, st>>>(
            X, DY, mean, rstd, rows_per_slice, C, seg, nsl, rpc, ws);
    } else {
        const int CT = C < kThreads ? C : kThreads;
        const size_t sm = (size_t)(kThreads / CT) * CT * sizeof(float2);
        colstats_partial<MODE><<<dim3(nchunk, nsl, cdiv(C, CT)), kThreads, sm, st>>>(
            X, DY, mean, rstd, rows_per_slice, C, seg, nsl, rpc, ws);
    }
    D2P_CHECK_LAUNCH();
    return 0;
}

size_t bn_ws_bytes(long long rows, int C, int nsl) {
    int rpc;
    int nchunk = pick_chunks(rows / nsl, nsl, &rpc);
    return (size_t)nsl * nchunk * C * sizeof(float2);
}

// Forward statistics + scale/shift. stats layout: mean|rstd|scale|shift, each [nsl, C].
int bn_forward_stats(cudaStream_t st, const float* X, long long rows, int C, int seg, int nsl,
                     const float* gamma, const float* beta, float* moving_mean,
                     float* moving_var, int training, float* stats, void* ws, size_t ws_bytes) {
    D2P_REQUIRE(rows % ((long long)seg * nsl) == 0, "bn: rows %lld not divisible by seg*nsl", rows);
    float* mean = stats; float* rstd = stats + (size_t)nsl * C;
    float* scale = rstd + (size_t)nsl * C; float* shift = scale + (size_t)nsl * C;
    int rpc = 0, nchunk = 0;
    if (training) {
        nchunk = pick_chunks(rows / nsl, nsl, &rpc);
        D2P_REQUIRE(ws_bytes >= (size_t)nsl * nchunk * C * sizeof(float2), "bn: workspace too small");
        D2P_TRY(launch_colstats<0>(st, X, nullptr, nullptr, nullptr, rows / nsl, C, seg, nsl, nchunk, rpc,
                                   (float2*)ws));
    }
    bn_fwd_finalize<<<cdiv(C, 8), 256, 0, st>>>((const float2*)ws, nchunk, C, nsl,
                                                 (double)(rows / nsl), gamma, beta, moving_mean,
                                                 moving_var, 1e-3f, 0.9f, training, mean, rstd,
                                                 scale, shift);
    D2P_CHECK_LAUNCH();
    return 0;
}

// one warp per (channel, slice): statistics from the partials; the moving averages follow in bn_moving_update
__global__ void bn_fwd_finalize_par(const float2* __restrict__ partial, int nchunk, int C, int nsl, double count,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                    float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ var_out) {
    const int c = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32, sl = blockIdx.y;
    if (c >= C) return;
    double s = 0.0, s2 = 0.0;
    for (int j = lane; j < nchunk; j += 32) {
        const float2 p = partial[((size_t)sl * nchunk + j) * C + c];
        s += p.x; s2 += p.y;
    }
    s = warp_sum_d(s); s2 = warp_sum_d(s2);
    const double mu = s / count;
    double var = s2 / count - mu * mu;
    if (var < 0.0) var = 0.0;
    if (lane == 0) {
        const float muf = (float)mu, rs = (float)(1.0 / sqrt(var + (double)eps)), g = gamma[c];
        mean_out[sl * C + c] = muf; rstd_out[sl * C + c] = rs;
        scale[sl * C + c] = g * rs; shift[sl * C + c] = beta[c] - muf * g * rs;
        var_out[sl * C + c] = (float)var;
    }
}
// slices in order, as the reference updates the moving statistics once per reuse call (i = 0..k-1)
__global__ void bn_moving_update(const float* __restrict__ mean, const float* __restrict__ var, int C, int nsl,
                                 float decay, float* __restrict__ moving_mean, float* __restrict__ moving_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mm = moving_mean[c], mv = moving_var[c];
    for (int sl = 0; sl < nsl; ++sl) {
        mm -= (mm - mean[sl * C + c]) * (1.f - decay);
        mv -= (mv - var[sl * C + c]) * (1.f - decay);
    }
    moving_mean[c] = mm; moving_var[c] = mv;
}

// Train-mode statistics from (sum, sum of squares) partials produced elsewhere (the tensor-core
// convolution's epilogue): partial[(sl*nchunk + chunk)*C + c], `count` rows per slice.
int bn_forward_finalize(cudaStream_t st, const float2* partial, int nchunk, long long count, int C, int nsl,
                        const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                        float* stats) {
    float* mean = stats; float* rstd = stats + (size_t)nsl * C;
    float* scale = rstd + (size_t)nsl * C; float* shift = scale + (size_t)nsl * C;
    float* var = (float*)(partial + (size_t)nsl * nchunk * C);   // [nsl, C] behind the partials
    bn_fwd_finalize_par<<<dim3(cdiv(C, 8), nsl), 256, 0, st>>>(partial, nchunk, C, nsl, (double)count, gamma, beta,
                                                                1e-3f, mean, rstd, scale, shift, var);
    D2P_CHECK_LAUNCH();
    bn_moving_update<<<cdiv(C, 128), 128, 0, st>>>(mean, var, C, nsl, 0.9f, moving_mean, moving_var);
    D2P_CHECK_LAUNCH();
    return 0;
}

int bn_apply(cudaStream_t st, const float* X, float* Y, long long rows, int C, int seg, int nsl,
             const float* stats, int permT, int permR) {
    const float* scale = stats + 2 * (size_t)nsl * C;
    const float* shift = stats + 3 * (size_t)nsl * C;
    if (nsl == 1 && permT <= 0 && C % 4 == 0 && rows * C < (1LL << 31) &&
        ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(scale) |
          reinterpret_cast<uintptr_t>(shift)) & 15) == 0)
        bn_apply_v4_kernel<<<ew_blocks(rows * C / 4), 256, 0, st>>>(
            reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(Y), (unsigned)(rows * C / 4),
            (unsigned)(C / 4), scale, shift);
    else
        bn_apply_kernel<<<ew_blocks(rows * C), 256, 0, st>>>(X, Y, rows, C, seg, nsl, scale, shift,
                                                            permT, permR);
    D2P_CHECK_LAUNCH();
    return 0;
}

// Backward through BN (and optionally the activation before it).
//   A: saved post-activation (pre-BN) values, DY: grad wrt BN output (row-permuted
//   if permT > 0), DZ: out grad wrt the pre-activation (or pre-BN value when !act).
//   coef: scratch [2, nsl, C]. dgamma/dbeta accumulate.
int bn_backward(cudaStream_t st, const float* A, const float* DY, float* DZ, long long rows, int C,
                int seg, int nsl, const float* gamma, const float* stats, float* dgamma,
                float* dbeta, int training, int act, float* coef, void* ws, size_t ws_bytes,
                int permT, int permR, float* dy_tmp) {
    const float* mean = stats; const float* rstd = stats + (size_t)nsl * C;
    const float* dy_lin = DY;
    if (permT > 0) {
        // statistics kernels read DY in A's row order
        D2P_REQUIRE(dy_tmp != nullptr, "bn_backward: permuted DY needs dy_tmp");
        permute_rows_kernel<<<ew_blocks(rows * C), 256, 0, st>>>(DY, dy_tmp, rows, C, permT, permR);
        D2P_CHECK_LAUNCH();
        dy_lin = dy_tmp;
    }
    int rpc;
    int nchunk = pick_chunks(rows / nsl, nsl, &rpc);
    D2P_REQUIRE(ws_bytes >= (size_t)nsl * nchunk * C * sizeof(float2), "bn: workspace too small");
    D2P_TRY(launch_colstats<1>(st, A, dy_lin, mean, rstd, rows / nsl, C, seg, nsl, nchunk, rpc, (float2*)ws));
    float* k1 = coef; float* k2 = coef + (size_t)nsl * C;
    bn_bwd_finalize<<<cdiv(C, 8), 256, 0, st>>>((const float2*)ws, nchunk, C, nsl,
                                                 (double)(rows / nsl), dgamma, dbeta, k1, k2,
                                                 training);
    D2P_CHECK_LAUNCH();
    const bool v4 = C % 4 == 0 && rows * C / 4 < (1LL << 32) && al16(A) && al16(dy_lin) && al16(DZ) &&
                    al16(gamma) && al16(mean) && al16(rstd) && al16(k1) && al16(k2);
    if (v4) {
        const unsigned total4 = (unsigned)(rows * C / 4);
        const int blocks = ew_blocks((long long)total4);
#define D2P_BWD_APPLY_V4(ACT_)                                                                          \
        bn_bwd_apply_v4<ACT_><<<blocks, 256, 0, st>>>(                                                  \
            (const float4*)A, (const float4*)dy_lin, (float4*)DZ, total4, (unsigned)(C / 4), (unsigned)seg, \
            (unsigned)nsl, (const float4*)gamma, (const float4*)mean, (const float4*)rstd,              \
            (const float4*)k1, (const float4*)k2)
        if (act) D2P_BWD_APPLY_V4(true); else D2P_BWD_APPLY_V4(false);
#undef D2P_BWD_APPLY_V4
    } else if (act)
        bn_bwd_apply_kernel<true><<<ew_blocks(rows * C), 256, 0, st>>>(
            A, dy_lin, DZ, rows, C, seg, nsl, gamma, mean, rstd, k1, k2, 0, 0);
    else
        bn_bwd_apply_kernel<false><<<ew_blocks(rows * C), 256, 0, st>>>(
            A, dy_lin, DZ, rows, C, seg, nsl, gamma, mean, rstd, k1, k2, 0, 0);
    D2P_CHECK_LAUNCH();
    return 0;
}

int colsum(cudaStream_t st, const float* X, long long rows, int C, float* out, float beta, void* ws,
           size_t ws_bytes);

// bn_backward(act = 1) followed by db += column sums of DZ, with the sums taken inside the apply pass (one read
// of DZ less: 655 MB for the first ViZDoom layer at C4).  Falls back to the two calls when the layout does not
// allow the 16-byte kernel.
int bn_backward_db(cudaStream_t st, const float* A, const float* DY, float* DZ, long long rows, int C, int seg,
                   int nsl, const float* gamma, const float* stats, float* dgamma, float* dbeta, float* db,
                   int training, float* coef, void* ws, size_t ws_bytes) {
    const float* mean = stats; const float* rstd = stats + (size_t)nsl * C;
    float* k1 = coef; float* k2 = coef + (size_t)nsl * C;
    int rpc_db;
    const int nchunk_db = pick_chunks(rows, 1, &rpc_db);
    const bool v4 = C % 4 == 0 && C / 4 <= kThreads && al16(A) && al16(DY) && al16(DZ) && al16(gamma) && al16(mean) &&
                    al16(rstd) && al16(k1) && al16(k2) && ws_bytes >= (size_t)nchunk_db * C * sizeof(float2);
    if (!v4) {
        D2P_TRY(bn_backward(st, A, DY, DZ, rows, C, seg, nsl, gamma, stats, dgamma, dbeta, training, 1, coef, ws,
                            ws_bytes, 0, 0, nullptr));
        return colsum(st, DZ, rows, C, db, 1.0f, ws, ws_bytes);
    }
    int rpc;
    const int nchunk = pick_chunks(rows / nsl, nsl, &rpc);
    D2P_REQUIRE(ws_bytes >= (size_t)nsl * nchunk * C * sizeof(float2), "bn: workspace too small");
    D2P_TRY(launch_colstats<1>(st, A, DY, mean, rstd, rows / nsl, C, seg, nsl, nchunk, rpc, (float2*)ws));
    bn_bwd_finalize<<<cdiv(C, 8), 256, 0, st>>>((const float2*)ws, nchunk, C, nsl, (double)(rows / nsl), dgamma, dbeta,
                                                 k1, k2, training);
    D2P_CHECK_LAUNCH();
    const int C4 = C / 4, RL = kThreads / C4;
    bn_bwd_apply_db_v4<<<nchunk_db, kThreads, (size_t)RL * C4 * 4 * sizeof(float), st>>>(
        (const float4*)A, (const float4*)DY, (float4*)DZ, rows, C4, seg, nsl, rpc_db, (const float4*)gamma,
        (const float4*)mean, (const float4*)rstd, (const float4*)k1, (const float4*)k2, (float2*)ws);
    D2P_CHECK_LAUNCH();
    colsum_finalize<<<cdiv(C, 8), 256, 0, st>>>((const float2*)ws, nchunk_db, C, db, 1.0f);
    D2P_CHECK_LAUNCH();
    return 0;
}

// out[c] = beta*out[c] + sum_rows X[row, c]
int colsum(cudaStream_t st, const float* X, long long rows, int C, float* out, float beta, void* ws,
           size_t ws_bytes) {
    int rpc;
    int nchunk = pick_chunks(rows, 1, &rpc);
    D2P_REQUIRE(ws_bytes >= (size_t)nchunk * C * sizeof(float2), "colsum: workspace too small");
    D2P_TRY(launch_colstats<0>(st, X, nullptr, nullptr, nullptr, rows, C, 1, 1, nchunk, rpc, (float2*)ws));
    colsum_finalize<<<cdiv(C, 8), 256, 0, st>>>((const float2*)ws, nchunk, C, out, beta);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace d2p

// ---- fully-connected -> (lrelu) -> BN, the reference's ops.fc (models/ops.py:149-155) ----
using namespace d2p;

static size_t fc_bn_part_bytes(long long rows, int Cout, int nsl) {
    size_t a = bn_ws_bytes(rows, Cout, nsl), b = bn_ws_bytes(rows, Cout, 1);
    return ((a > b ? a : b) + 255) & ~(size_t)255;
}

extern "C" size_t d2p_fc_bn_saved_floats(long long rows, int Cout, int nsl) {
    return (size_t)rows * Cout + 4 * (size_t)nsl * Cout;
}

extern "C" size_t d2p_fc_bn_ws_bytes(long long rows, int Cout, int nsl) {
    size_t dz = (((size_t)rows * Cout * sizeof(float)) + 255) & ~(size_t)255;
    size_t coef = ((2 * (size_t)nsl * Cout * sizeof(float)) + 255) & ~(size_t)255;
    return dz + coef + fc_bn_part_bytes(rows, Cout, nsl);
}

namespace d2p { namespace {
__global__ void lrelu_inplace_k(float* __restrict__ x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        x[i] = lrelu_f(x[i]);
}
} }

extern "C" int d2p_fc_bn_fwd(const float* X, long long rows, int Cin, int Cout, int seg, int nsl,
                             int act, const d2p_fc_bn* p, float* Y, float* saved, int training,
                             void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(X && p && Y && saved && ws, "fc_bn fwd: null buffer");
    D2P_REQUIRE(ws_bytes >= d2p_fc_bn_ws_bytes(rows, Cout, nsl), "fc_bn fwd: workspace too small");
    float* A = saved; float* stats = saved + (size_t)rows * Cout;
    D2P_TRY(gemm(st, false, false, (int)rows, Cout, Cin, 1.f, X, Cin, p->w, Cout, 0.f, A, Cout, p->b));
    if (act) {
        lrelu_inplace_k<<<ew_blocks(rows * Cout), 256, 0, st>>>(A, (size_t)rows * Cout);
        D2P_CHECK_LAUNCH();
    }
    D2P_TRY(bn_forward_stats(st, A, rows, Cout, seg, nsl, p->gamma, p->beta, p->moving_mean,
                             p->moving_var, training, stats, ws, ws_bytes));
    D2P_TRY(bn_apply(st, A, Y, rows, Cout, seg, nsl, stats, 0, 0));
    return 0;
}

extern "C" int d2p_fc_bn_bwd(const float* X, long long rows, int Cin, int Cout, int seg, int nsl,
                             int act, const d2p_fc_bn* p, const float* dY, const float* saved,
                             float* dX, int training, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    D2P_REQUIRE(X && p && dY && saved && ws, "fc_bn bwd: null buffer");
    D2P_REQUIRE(p->dw && p->db && p->dgamma && p->dbeta, "fc_bn bwd: null grad buffer");
    D2P_REQUIRE(ws_bytes >= d2p_fc_bn_ws_bytes(rows, Cout, nsl), "fc_bn bwd: workspace too small");
    const float* A = saved; const float* stats = saved + (size_t)rows * Cout;
    char* w = (char*)ws;
    size_t dzb = (((size_t)rows * Cout * sizeof(float)) + 255) & ~(size_t)255;
    size_t coefb = ((2 * (size_t)nsl * Cout * sizeof(float)) + 255) & ~(size_t)255;
    float* dZ = (float*)w; float* coef = (float*)(w + dzb);
    void* part = w + dzb + coefb; size_t part_bytes = ws_bytes - dzb - coefb;
    D2P_TRY(bn_backward(st, A, dY, dZ, rows, Cout, seg, nsl, p->gamma, stats, p->dgamma, p->dbeta,
                        training, act, coef, part, part_bytes, 0, 0, nullptr));
    D2P_TRY(colsum(st, dZ, rows, Cout, p->db, 1.f, part, part_bytes));
    D2P_TRY(gemm(st, true, false, Cin, Cout, (int)rows, 1.f, X, Cin, dZ, Cout, 1.f, p->dw, Cout));
    if (dX) D2P_TRY(gemm(st, false, true, (int)rows, Cin, Cout, 1.f, dZ, Cout, p->w, Cout, 0.f, dX, Cin));
    return 0;
}
