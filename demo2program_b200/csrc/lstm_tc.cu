// Tensor-core LSTM recurrence (tcgen05) with the cell non-linearity fused into
// the GEMM epilogue.
//
// Forward step t (one launch): z = h_{t-1} * Wh (+ the hoisted X*Wx + b already
// in gates[t]); the recurrent weight is packed ONCE per sequence with its
// columns permuted so that every 64-wide output tile holds the i, j, f, o
// columns of 16 hidden units.  The epilogue thread that owns accumulator row r
// therefore sees all four gates of its units, applies BasicLSTMCell
// (reference models/model_full.py:244) with dynamic_rnn masking
// (A.5), and writes h_t both as fp32 (Y, state) and directly in the packed bf16
// hi/lo operand format the next step's MMA consumes - no separate gate kernel,
// no separate split/pack pass.
//
// Backward step t (two launches): a fused element-wise kernel combines the
// split-K partial sums of dh (fixed order), back-propagates through the cell
// and writes dZ_t as fp32 (for the dW / dX products) and as the packed A
// operand; then one split-K tcgen05 GEMM produces the partial sums of
// dh_{t-1} = dZ_t * Wh^T over all SMs.
#include "tc_common.cuh"

namespace d2p {

using namespace tc;

bool tc_available();
int gemm_tc_nsplit(int K, int ksplit);
int colsum(cudaStream_t, const float*, long long, int, float*, float, void*, size_t);
// fp32 CUDA-core step kernels for R <= 32 (lstm_skinny.cu)
bool lstm_skinny_supported(int R, int H);
int lstm_skinny_nsplit(int H);
int lstm_skinny_fwd_step(cudaStream_t, const float*, int, int, float*, float*, float*, const float*, float*, float*,
                         const int*, int, float);
int lstm_skinny_bwd_step(cudaStream_t, const float*, const float*, int, int, float*);
// persistent weight-stationary recurrences (lstm_persist.cu)
bool lstm_persist_supported(int R, int H);
int lstm_persist_fwd(cudaStream_t, int, int, int, const int*, const float*, const float*, const float*, float,
                     float*, float*, float*, float*, float*, bool compact, bool wide);
int lstm_persist_bwd(cudaStream_t, int, int, int, const int*, const float*, const float*, const float*, float*,
                     const float*, const float*, const float*, const float*, float*, float*, float*, const void**,
                     size_t*, bool wide);

namespace {

constexpr int LBN = 64, LSTAGES = 3, UPT = LBN / 4;   // 16 hidden units per tile
constexpr int LTHREADS = 512;                         // 2 role warps + 14 helper warps
constexpr int ZROW = LBN + 4, SROW = UPT + 4;         // padded smem rows (floats)
constexpr size_t L_PIPE_BYTES = tc_smem_bytes<LBN, LSTAGES>();
constexpr size_t L_SMEM_BYTES = L_PIPE_BYTES + (size_t)BM * (ZROW + 2 * SROW) * sizeof(float);

// Epilogue inputs of one 128-row x 16-unit tile, fetched into smem with coalesced
// 16-byte loads by the helper warps WHILE the main loop runs.
struct StepPrefetch {
    const float* gates_t; const float* cstate; const float* hstate;
    float* zs; float* cs; float* hs;
    int m0, u0, R, H;
    __device__ __forceinline__ void operator()(int w, int nw) const {
        for (int idx = w; idx < BM * LBN / 4; idx += nw) {          // (row, gate) segments of 64 B
            const int f4 = (idx & 3) * 4, seg = idx >> 2;
            const int gate = seg & 3, row = seg >> 2;
            if (m0 + row < R)
                *reinterpret_cast<float4*>(zs + (size_t)row * ZROW + gate * UPT + f4) =
                    *reinterpret_cast<const float4*>(gates_t + (size_t)(m0 + row) * 4 * H + gate * H + u0 + f4);
        }
        for (int idx = w; idx < BM * UPT / 4; idx += nw) {
            const int f4 = (idx & 3) * 4, row = idx >> 2;
            if (m0 + row < R) {
                const size_t g = (size_t)(m0 + row) * H + u0 + f4;
                *reinterpret_cast<float4*>(cs + (size_t)row * SROW + f4) = *reinterpret_cast<const float4*>(cstate + g);
                *reinterpret_cast<float4*>(hs + (size_t)row * SROW + f4) = *reinterpret_cast<const float4*>(hstate + g);
            }
        }
    }
};

__global__ void __launch_bounds__(LTHREADS)
lstm_step_fwd_kernel(Packed A, Packed B, int R, int H, float* __restrict__ gates_t,
                     float* __restrict__ cells_t, float* __restrict__ Y_t,
                     float* __restrict__ hstate, float* __restrict__ cstate,
                     const int* __restrict__ len, int t, float forget_bias,
                     uint8_t* __restrict__ hpk_next, int hpk_mgp) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * LBN;
    const int u0 = blockIdx.x * UPT;
    float* zs = reinterpret_cast<float*>(smem + L_PIPE_BYTES);      // [128][ZROW] pre-activations
    float* cs = zs + (size_t)BM * ZROW;                             // [128][SROW] c_{t-1}
    float* hs = cs + (size_t)BM * SROW;                             // [128][SROW] h_{t-1}
    StepPrefetch pf{gates_t, cstate, hstate, zs, cs, hs, m0, u0, R, H};
    const uint32_t tmem_d = tc_mainloop<LBN, LSTAGES>(A, B, m0, n0, 0, (H + BK - 1) / BK, smem, pf);

    __syncthreads();   // prefetched tiles (helper warps) visible to every warp

    // ---- epilogue, all 16 warps ----
    // (1) z += accumulator: warp w adds gate (w / 4) for the 32 rows of TMEM lane quarter (w % 4)
    {
        const int q = warp & 3, g = warp >> 2;
        uint32_t v[UPT];
        tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * UPT), v);
        tmem_ld_wait();
        float* zr = zs + (size_t)(q * 32 + lane) * ZROW + g * UPT;
#pragma unroll
        for (int j = 0; j < UPT; j += 4) {
            float4 z = *reinterpret_cast<float4*>(zr + j);
            z.x += __uint_as_float(v[j]);     z.y += __uint_as_float(v[j + 1]);
            z.z += __uint_as_float(v[j + 2]); z.w += __uint_as_float(v[j + 3]);
            *reinterpret_cast<float4*>(zr + j) = z;
        }
    }
    __syncthreads();
    // (2) cell math: thread = (row, 4 hidden units); 64-byte segments per 4 lanes -> coalesced
    {
        const int row = tid >> 2, uq = (tid & 3) * 4;
        const int r = m0 + row;
        if (r < R) {
            const bool live = t < len[r];
            const float* zr = zs + (size_t)row * ZROW + uq;
            const float4 hp = *reinterpret_cast<const float4*>(hs + (size_t)row * SROW + uq);
            const size_t gu = (size_t)r * H + u0 + uq;
            float hn[4] = {hp.x, hp.y, hp.z, hp.w};
            if (live) {
                const float4 zi = *reinterpret_cast<const float4*>(zr);
                const float4 zj = *reinterpret_cast<const float4*>(zr + UPT);
                const float4 zf = *reinterpret_cast<const float4*>(zr + 2 * UPT);
                const float4 zo = *reinterpret_cast<const float4*>(zr + 3 * UPT);
                const float4 cp = *reinterpret_cast<const float4*>(cs + (size_t)row * SROW + uq);
                const float pi[4] = {zi.x, zi.y, zi.z, zi.w}, pj[4] = {zj.x, zj.y, zj.z, zj.w};
                const float pf4[4] = {zf.x, zf.y, zf.z, zf.w}, po[4] = {zo.x, zo.y, zo.z, zo.w};
                const float pc[4] = {cp.x, cp.y, cp.z, cp.w};
                float gi[4], gj[4], gf[4], go[4], cn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    gi[e] = sigmoid_fast(pi[e]);
                    gj[e] = tanh_fast(pj[e]);
                    gf[e] = sigmoid_fast(pf4[e] + forget_bias);
                    go[e] = sigmoid_fast(po[e]);
                    cn[e] = pc[e] * gf[e] + gi[e] * gj[e];
                    hn[e] = tanh_fast(cn[e]) * go[e];
                }
                float* grow = gates_t + (size_t)r * 4 * H + u0 + uq;
                *reinterpret_cast<float4*>(grow) = make_float4(gi[0], gi[1], gi[2], gi[3]);
                *reinterpret_cast<float4*>(grow + H) = make_float4(gj[0], gj[1], gj[2], gj[3]);
                *reinterpret_cast<float4*>(grow + 2 * H) = make_float4(gf[0], gf[1], gf[2], gf[3]);
                *reinterpret_cast<float4*>(grow + 3 * H) = make_float4(go[0], go[1], go[2], go[3]);
                const float4 c4 = make_float4(cn[0], cn[1], cn[2], cn[3]);
                const float4 h4 = make_float4(hn[0], hn[1], hn[2], hn[3]);
                *reinterpret_cast<float4*>(cells_t + gu) = c4;
                *reinterpret_cast<float4*>(cstate + gu) = c4;
                *reinterpret_cast<float4*>(Y_t + gu) = h4;
                *reinterpret_cast<float4*>(hstate + gu) = h4;
            } else {
                // t >= len: output row is zero, (c, h) are copied through (dynamic_rnn, A.5)
                *reinterpret_cast<float4*>(cells_t + gu) =
                    *reinterpret_cast<const float4*>(cs + (size_t)row * SROW + uq);
                *reinterpret_cast<float4*>(Y_t + gu) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // h_t (or the copied-through h_{t-1}) in packed operand format for step t+1
            store_packed4(hpk_next, hpk_mgp, r, u0 + uq, hn);
        }
    }
    tc_teardown<LBN>(tmem_d);
}

__device__ __forceinline__ void ld4(const float* __restrict__ p, float* x) {
    float4 a = *reinterpret_cast<const float4*>(p);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
}
__device__ __forceinline__ void st4(float* __restrict__ p, const float* x) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
}

// Element-wise backward through the cell at step t for 4 hidden units of one row.
//   dh_in = dhc + sum_s partial_s  (partial sums of dZ_{t+1} * Wh^T) + dY_t
__global__ void __launch_bounds__(256)
lstm_bwd_point_kernel(float* __restrict__ G /*[R,4H] in: gates, out: dZ*/,
                      const float* __restrict__ cells_t, const float* __restrict__ cells_prev,
                      const float* __restrict__ c0, const float* __restrict__ dY_t,
                      const float* __restrict__ partials, int nsplit, float* __restrict__ dhc,
                      float* __restrict__ dcs, const int* __restrict__ len, int t, int R, int H,
                      uint8_t* __restrict__ dzpk, int dz_mgp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ug = H / 4;
    if (idx >= R * ug) return;
    const int r = idx / ug, u0 = (idx % ug) * 4;
    const size_t su = (size_t)r * H + u0;
    float dh[4], tmp[4];
    ld4(dhc + su, dh);
    for (int s = 0; s < nsplit; ++s) {
        ld4(partials + (size_t)s * R * H + su, tmp);
#pragma unroll
        for (int e = 0; e < 4; ++e) dh[e] += tmp[e];
    }
    float* g = G + (size_t)r * 4 * H + u0;
    float di[4], dj[4], df[4], dq[4];
    if (t < len[r]) {
        float gi[4], gj[4], gf[4], go[4], c[4], cp[4], dy[4], dc[4];
        ld4(g, gi); ld4(g + H, gj); ld4(g + 2 * H, gf); ld4(g + 3 * H, go);
        ld4(cells_t + su, c);
        if (cells_prev) ld4(cells_prev + su, cp);
        else if (c0) ld4(c0 + su, cp);
        else {
#pragma unroll
            for (int e = 0; e < 4; ++e) cp[e] = 0.f;
        }
        if (dY_t) ld4(dY_t + su, dy);
        else {
#pragma unroll
            for (int e = 0; e < 4; ++e) dy[e] = 0.f;
        }
        ld4(dcs + su, dc);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float dht = dh[e] + dy[e];
            float tcn = tanh_fast(c[e]);
            dq[e] = dht * tcn * go[e] * (1.f - go[e]);
            float dct = dc[e] + dht * go[e] * (1.f - tcn * tcn);
            di[e] = dct * gj[e] * gi[e] * (1.f - gi[e]);
            dj[e] = dct * gi[e] * (1.f - gj[e] * gj[e]);
            df[e] = dct * cp[e] * gf[e] * (1.f - gf[e]);
            dc[e] = dct * gf[e];
            tmp[e] = 0.f;
        }
        st4(dcs + su, dc);
        st4(dhc + su, tmp);   // consumed: the recurrent GEMM's partial sums carry the new dh
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) di[e] = dj[e] = df[e] = dq[e] = 0.f;
        st4(dhc + su, dh);    // state copied through: gradient passes unchanged
    }
    st4(g, di); st4(g + H, dj); st4(g + 2 * H, df); st4(g + 3 * H, dq);
    if (!dzpk) return;   // fp32 (small-R) recurrence reads dZ from G
    store_packed4(dzpk, dz_mgp, r, u0, di);
    store_packed4(dzpk, dz_mgp, r, H + u0, dj);
    store_packed4(dzpk, dz_mgp, r, 2 * H + u0, df);
    store_packed4(dzpk, dz_mgp, r, 3 * H + u0, dq);
}

// dst += sum_s partial_s
__global__ void add_partials_kernel(float* __restrict__ dst, const float* __restrict__ partials,
                                    int nsplit, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = dst[i];
    for (int s = 0; s < nsplit; ++s) a += partials[(size_t)s * n + i];
    dst[i] = a;
}

__global__ void copy_or_zero_k(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src ? src[i] : 0.f;
}

}  // namespace

int lstm_tc_set_probe(long long* buf) {
    D2P_CHECK_CUDA(cudaMemcpyToSymbol(tc::g_tc_dbg, &buf, sizeof(buf)));
    return 0;
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

bool lstm_tc_supported(int R, int H) {
    return tc_available() && H % 64 == 0 && H >= 64 && R >= 1;
}

int lstm_seq_fwd_tc(cudaStream_t st, const float* X, int T, int R, int In, int H, const int* len,
                    const float* h0, const float* c0, const float* W, const float* b,
                    float forget_bias, float* Y, float* hT, float* cT, float* gates, float* cells,
                    int phases) {
    const int G4 = 4 * H;
    const float* Wx = W;
    const float* Wh = W + (size_t)In * G4;
    const size_t RH = (size_t)R * H;
    const int eb = cdiv(RH, 256);
    // phase 1: hoisted input contraction for all steps: gates = X*Wx + b
    if (phases & D2P_LSTM_INPUT)
        D2P_TRY(gemm(st, false, false, T * R, G4, In, 1.f, X, In, Wx, G4, 0.f, gates, G4, b, GEMM_CONST_B));
    if (!(phases & D2P_LSTM_RECUR)) return 0;
    // phase 2: recurrence (the arena is reused from offset 0; stream order makes that safe)
    size_t off = 0;
    if (lstm_persist_supported(R, H) && aligned16(Wh) && aligned16(gates) && aligned16(hT) && aligned16(cT) &&
        (!h0 || aligned16(h0)) && (!c0 || aligned16(c0)))
        return lstm_persist_fwd(st, T, R, H, len, h0, c0, Wh, forget_bias, Y, hT, cT, gates, cells,
                                (phases & D2P_LSTM_COMPACT) != 0, (phases & D2P_LSTM_WIDE) != 0);
    if (lstm_skinny_supported(R, H) && aligned16(Wh) && aligned16(hT) && aligned16(gates)) {
        // every CTA reads all of h_{t-1}: ping-pong between hT and a scratch copy
        float* hb[2] = {hT, (float*)tc_scratch_alloc(st, &off, RH * sizeof(float))};
        D2P_REQUIRE(hb[1], "lstm fwd: scratch arena too small");
        copy_or_zero_k<<<eb, 256, 0, st>>>(hT, h0, RH);
        D2P_CHECK_LAUNCH();
        copy_or_zero_k<<<eb, 256, 0, st>>>(cT, c0, RH);
        D2P_CHECK_LAUNCH();
        for (int t = 0; t < T; ++t)
            D2P_TRY(lstm_skinny_fwd_step(st, Wh, R, H, gates + (size_t)t * R * G4, cells + t * RH, Y + t * RH,
                                         hb[t & 1], hb[(t + 1) & 1], cT, len, t, forget_bias));
        if (T & 1) D2P_CHECK_CUDA(cudaMemcpyAsync(hT, hb[1], RH * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const size_t hbytes = packed_bytes(R, H);
    uint8_t* hpk[2];
    hpk[0] = (uint8_t*)tc_scratch_alloc(st, &off, hbytes);
    hpk[1] = (uint8_t*)tc_scratch_alloc(st, &off, hbytes);
    D2P_REQUIRE(hpk[0] && hpk[1], "lstm fwd: tensor-core scratch arena too small");
    const void* whpk;
    D2P_TRY(get_packed(st, Wh, G4, H, G4, false, true, &off, &whpk, LBN, H));
    if (h0) D2P_TRY(pack_bf16(st, h0, R, H, H, true, hpk[0]));
    else D2P_CHECK_CUDA(cudaMemsetAsync(hpk[0], 0, hbytes, st));
    D2P_CHECK_CUDA(cudaMemsetAsync(hpk[1], 0, hbytes, st));
    copy_or_zero_k<<<eb, 256, 0, st>>>(hT, h0, RH);
    D2P_CHECK_LAUNCH();
    copy_or_zero_k<<<eb, 256, 0, st>>>(cT, c0, RH);
    D2P_CHECK_LAUNCH();
    constexpr size_t smem = L_SMEM_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(lstm_step_fwd_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int mgp_h = mgp_of(R);
    dim3 grid(G4 / LBN, cdiv(R, BM));
    for (int t = 0; t < T; ++t) {
        Packed A{hpk[t & 1], mgp_h};
        Packed B{(const uint8_t*)whpk, mgp_of(G4)};
        lstm_step_fwd_kernel<<<grid, LTHREADS, smem, st>>>(A, B, R, H, gates + (size_t)t * R * G4,
                                                      cells + t * RH, Y + t * RH, hT, cT, len, t,
                                                      forget_bias, hpk[(t + 1) & 1], mgp_h);
        D2P_CHECK_LAUNCH();
    }
    return 0;
}

int lstm_seq_bwd_tc(cudaStream_t st, const float* X, int T, int R, int In, int H, const int* len,
                    const float* h0, const float* c0, const float* W, const float* Y, float* gates,
                    const float* cells, const float* dY, const float* dhT, const float* dcT,
                    float* dX, float* dW, float* db, float* dh0, float* dc0, void* ws,
                    size_t ws_bytes, int phases) {
    const int G4 = 4 * H;
    const float* Wx = W;
    const float* Wh = W + (size_t)In * G4;
    float* dWx = dW;
    float* dWh = dW + (size_t)In * G4;
    const size_t RH = (size_t)R * H;
    const int eb = cdiv(RH, 256);
    const bool persist = lstm_persist_supported(R, H) && aligned16(Wh) && aligned16(gates) && aligned16(dh0) &&
                         aligned16(dc0) && (!dhT || aligned16(dhT)) && (!dcT || aligned16(dcT)) &&
                         (!dY || aligned16(dY)) && (!c0 || aligned16(c0));
    if ((phases & D2P_LSTM_BWD_RECUR) && persist) {
        const void* dzfull = nullptr;
        size_t off = 0;
        D2P_TRY(lstm_persist_bwd(st, T, R, H, len, h0, c0, Wh, gates, cells, dY, dhT, dcT, dh0, dc0, db,
                                 dX ? &dzfull : nullptr, &off, (phases & D2P_LSTM_WIDE) != 0));
        if (dX && dzfull && tc_eligible(T * R, In, G4)) {   // dX = dZ * Wx^T from the operand the kernel packed
            const void* wxpk;
            D2P_TRY(get_packed(st, Wx, In, G4, G4, true, true, &off, &wxpk));
            D2P_TRY(gemm_tc_packed_auto(st, dzfull, wxpk, T * R, In, G4, 1.f, 0.f, dX, In, &off));
        } else if (dX) {
            D2P_TRY(gemm(st, false, true, T * R, In, G4, 1.f, gates, G4, Wx, G4, 0.f, dX, In, nullptr, GEMM_CONST_B));
        }
    } else if (phases & D2P_LSTM_BWD_RECUR) {
    // split-K factor of the per-step dh GEMM [R, H] = dZ_t [R, 4H] * Wh^T
        long long tiles64 = (long long)cdiv(H, 64) * cdiv(R, BM);
        int ks = (int)(144 / (tiles64 < 1 ? 1 : tiles64));
        const int nkb = cdiv(G4, BK);
        if (ks > nkb / 4) ks = nkb / 4;
        if (ks > 8) ks = 8;
        if (ks < 1) ks = 1;
        const bool skinny = lstm_skinny_supported(R, H) && aligned16(Wh) && aligned16(gates);
        const int nsplit = skinny ? lstm_skinny_nsplit(H) : gemm_tc_nsplit(G4, ks);

        size_t off = 0;
        const size_t zbytes = packed_bytes(R, G4);
        uint8_t* dzpk = (uint8_t*)tc_scratch_alloc(st, &off, zbytes);
        float* partials = (float*)tc_scratch_alloc(st, &off, (size_t)nsplit * RH * sizeof(float));
        D2P_REQUIRE(dzpk && partials, "lstm bwd: tensor-core scratch arena too small");
        const void* whpk = nullptr;   // Op_B[n = hidden unit, k = gate column] = Wh[n, k]
        if (!skinny) {
            D2P_TRY(get_packed(st, Wh, H, G4, G4, true, true, &off, &whpk));
            D2P_CHECK_CUDA(cudaMemsetAsync(dzpk, 0, zbytes, st));
        }
        copy_or_zero_k<<<eb, 256, 0, st>>>(dh0, dhT, RH);   // dh0/dc0 double as the running carries
        D2P_CHECK_LAUNCH();
        copy_or_zero_k<<<eb, 256, 0, st>>>(dc0, dcT, RH);
        D2P_CHECK_LAUNCH();
        const int mgp_z = mgp_of(R);
        const int pb = cdiv((long long)R * (H / 4), 256);
        bool have_partials = false;
        for (int t = T - 1; t >= 0; --t) {
            float* Gt = gates + (size_t)t * R * G4;
            lstm_bwd_point_kernel<<<pb, 256, 0, st>>>(Gt, cells + t * RH, t > 0 ? cells + (t - 1) * RH : nullptr,
                                                      c0, dY ? dY + t * RH : nullptr, partials,
                                                      have_partials ? nsplit : 0, dh0, dc0, len, t, R, H,
                                                      skinny ? nullptr : dzpk, mgp_z);
            D2P_CHECK_LAUNCH();
            if (t > 0 || h0 != nullptr) {   // partial sums of dh_{t-1} = dZ_t * Wh^T
                if (skinny) D2P_TRY(lstm_skinny_bwd_step(st, Gt, Wh, R, H, partials));
                else D2P_TRY(gemm_tc_packed(st, dzpk, whpk, R, H, G4, 1.f, 0.f, nullptr, H, nullptr, ks, partials));
                have_partials = true;
            } else {
                have_partials = false;
            }
        }
        if (have_partials) {   // dh0 = carry + last partial sums
            add_partials_kernel<<<eb, 256, 0, st>>>(dh0, partials, nsplit, RH);
            D2P_CHECK_LAUNCH();
        }
        // input gradient from the full dZ (arena reused from offset 0)
        if (dX) D2P_TRY(gemm(st, false, true, T * R, In, G4, 1.f, gates, G4, Wx, G4, 0.f, dX, In, nullptr, GEMM_CONST_B));
    }
    if (!(phases & D2P_LSTM_BWD_PARAMS)) return 0;
    // parameter gradients from the full dZ
    const size_t shared_need = al256(packed_bytes(G4, T * R)) + al256(packed_bytes(In, T * R)) +
                               al256(packed_bytes(H, T * R)) + 2 * al256((size_t)8 * (In > H ? In : H) * G4 * 4);
    if (tc_available() && R % BK == 0 && T > 1 && tc_eligible(In, G4, T * R) && tc_eligible(H, G4, (T - 1) * R) &&
        shared_need <= tc_scratch_capacity(st)) {
        // dWx = X^T dZ and dWh = Y[0..T-2]^T dZ[1..T-1] contract over the (time, row) axis of the
        // same dZ: pack it ONCE (k-blocks are whole time-row groups when R % 64 == 0; the
        // recurrent product starts R / 64 k-blocks in)
        size_t off = 0;
        const void *zpk, *xpk, *ypk;
        D2P_TRY(get_packed(st, gates, G4, T * R, G4, false, false, &off, &zpk));
        if (!(phases & D2P_LSTM_BWD_NO_DWX)) {
            D2P_TRY(get_packed(st, X, In, T * R, In, false, false, &off, &xpk));
            D2P_TRY(gemm_tc_packed_auto(st, xpk, zpk, In, G4, T * R, 1.f, 1.f, dWx, G4, &off));
        }
        D2P_TRY(get_packed(st, Y, H, (T - 1) * R, H, false, false, &off, &ypk));
        const uint8_t* zsub = (const uint8_t*)zpk + (size_t)(R / BK) * mgp_of(G4) * 2048;
        D2P_TRY(gemm_tc_packed_auto(st, ypk, zsub, H, G4, (T - 1) * R, 1.f, 1.f, dWh, G4, &off));
    } else {
        if (!(phases & D2P_LSTM_BWD_NO_DWX))
            D2P_TRY(gemm(st, true, false, In, G4, T * R, 1.f, X, In, gates, G4, 1.f, dWx, G4));
        if (T > 1)
            D2P_TRY(gemm(st, true, false, H, G4, (T - 1) * R, 1.f, Y, H, gates + (size_t)R * G4, G4, 1.f, dWh, G4));
    }
    if (h0) D2P_TRY(gemm(st, true, false, H, G4, R, 1.f, h0, H, gates, G4, 1.f, dWh, G4));
    // (the persistent recurrence kernel has already accumulated db while it produced dZ)
    if (!persist) D2P_TRY(colsum(st, gates, (long long)T * R, G4, db, 1.f, ws, ws_bytes));
    return 0;
}

}  // namespace d2p
