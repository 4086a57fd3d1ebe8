// Tensor-core (tcgen05 / TMEM) implicit-GEMM kernels of the State_Encoder convolutions
// (reference models/ops.py:27-33 slim.conv2d 3x3 / stride 2 / SAME -> lrelu -> batch_norm, called from
// models/model_full.py:219-229): every layer whose input has 16 / 32 / 48 channels, i.e. ViZDoom
// conv2-5 and Karel conv2-3 of the per-layer path.  No im2col buffer exists anywhere:
//
//  * conv_tc_gather_kernel - forward AND input gradient.  Rows of the implicit GEMM are destination
//    pixels (128 per tile, tiles never straddle a demonstration so a tile belongs to ONE BatchNorm
//    slice); the K dimension is walked tap by tap.  Eight producer warps gather, for one tap, the
//    CSRC contiguous NHWC channels of each row's source pixel (BatchNorm of the previous layer applied
//    on the fly, zero outside the image), split them into bf16 hi / lo and store them as UMMA K-major
//    core matrices into a shared-memory ring (generic stores + fence.proxy.async -> mbarrier); one
//    lane issues `Ahi x [Whi|Wlo]` as a single N = 2*NOUT tcgen05.mma plus `Alo x Whi` (the bf16x3
//    split of tc_common.cuh: fp32-equivalent products) against the whole layer's packed weights, which
//    stay resident in shared memory; the accumulator is double-buffered in TMEM so that four epilogue
//    warps (tcgen05.ld -> bias -> lrelu -> staging tile -> coalesced stores + the per-tile BatchNorm
//    (sum, sum of squares) partials) overlap the next tile's gathers and MMAs.
//    The input gradient is the same kernel: destination = input pixels of one parity class (the taps
//    that reach a pixel depend only on the parity of its coordinates), source = dZ, weights transposed.
//  * conv_tc_dw_kernel - weight gradient dW[tap, ci, co] = sum_pixels in[pixel @ tap, ci] dZ[pixel, co]:
//    M = 9*CIN rows (two to four 128-row accumulators in TMEM), N = COUT, K = pixels.  Both operands
//    are MN-major in memory (channels contiguous per pixel), so the producers store them as MN-major
//    core matrices (8 pixels x 8 channels, 16 bytes per pixel) and the instruction descriptor carries
//    a_major = b_major = MN.  Each CTA owns a contiguous pixel range; per-CTA partial tiles are summed
//    in a fixed order by conv_tc_dw_reduce (deterministic).
#include "tc_common.cuh"
#include "conv_tc.cuh"
#include <cstdlib>

namespace d2p {

bool tc_available();
void conv_set_rgb(int on);

namespace {
using namespace tc;

constexpr int kRows = 128;                    // rows (pixels) of an accumulator tile
constexpr int kProd = 384;                    // producer threads (warps 5..16): kGroups groups of four warps (17 warps are allocated as 20: 96 registers each)
constexpr int kGroups = kProd / 128;
constexpr int kEpi = 128;                     // epilogue threads (warps 0..3 = the four TMEM lane quarters)
constexpr int kThreadsTc = kEpi + 32 + kProd; // warp 4: MMA issue + TMEM allocation
constexpr int kMaxStages = 8;
constexpr size_t kSmemLimit = 227 * 1024 - 1024;   // dynamic shared memory we allow ourselves

int g_conv_tc_mode = 7;
int g_conv_tc_swap = 0;   // developer switch: exchange LBO / SBO of the MN-major descriptors
int g_conv_tc_quad = 1;   // d2p_conv_set_tc bit 5 SET: input gradient per parity class instead of the quad form

__device__ __forceinline__ void proxy_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// two fp32 -> packed bf16x2 hi and lo (one cvt.rn.bf16x2.f32 each)
__device__ __forceinline__ void split2p(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (16 bytes each)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    split2(x[0], x[1], hi.x, lo.x);
    split2(x[2], x[3], hi.y, lo.y);
    split2(x[4], x[5], hi.z, lo.z);
    split2(x[6], x[7], hi.w, lo.w);
}
__device__ __forceinline__ void st_shared16(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// forward / input-gradient gather kernel
// ------------------------------------------------------------------------------------------------
struct TapClass {
    int tile_begin;          // first tile of this class
    int tpd;                 // tiles per demonstration
    int DH, DW;              // grid of destination pixels per frame in this class
    int doff_y, doff_x;      // destination pixel = dmul * (y, x) + doff
    int ntaps;
    signed char offy[9], offx[9], wtap[9];   // source pixel = smul * (y, x) + off; index of the weight image
};
struct GatherParams {
    const float* src;        // [R*T, SH, SW, CSRC]
    float* out;              // [R*T, OH, OW, NOUT]
    const uint8_t* wimg;     // 9 packed weight images
    const float* bias;       // [NOUT] or null
    const float* scale;      // [k, CSRC] affine on the source (BatchNorm of the previous layer) or null
    const float* shift;
    float2* partial;         // [k, nchunk, NOUT] (sum, sum of squares) of the outputs, or null
    int R, T, k;             // demonstrations (B*k), frames per demonstration, BatchNorm slices
    int SH, SW, OH, OW;
    int smul, dmul;
    int act;                 // epilogue: + bias, lrelu
    int quad;                // > 0: NOUT = 4*quad columns are the four parity classes (py, px) x `quad` channels of the
                             // 2x2 destination block at (2y, 2x): the input gradient as ONE stride-1 product
    int nw;                  // weight images (9 taps, or 4 in quad mode)
    int nclass, ntiles, nchunk, stages;
    int ngroups;             // active producer groups: <= kGroups and ngroups * D <= stages (see the producers)
    TapClass cls[4];
};

template <int CSRC, int NOUT>
struct GatherCfg {
    static constexpr int G = CSRC / 8;                 // 16-byte channel groups (8 x bf16) per tap
    static constexpr int D = G == 2 ? 2 : 1;           // (tile, tap) items per producer batch
    static constexpr uint32_t STAGE = CSRC * 512;      // one tap: 128 rows x CSRC x (hi, lo)
    static constexpr uint32_t WTAP = CSRC * NOUT * 4;  // one tap of weights: NOUT x CSRC x (hi, lo)
    static constexpr int OST = NOUT + 4;               // staging row pitch in floats (conflict-free float4 rows)
    static constexpr uint32_t OST_BYTES = kRows * OST * 4;
    static constexpr uint32_t TCOLS = NOUT <= 16 ? 64 : (NOUT <= 32 ? 128 : (NOUT <= 64 ? 256 : 512));   // 2 buffers x 2*NOUT columns
    static_assert(4 * NOUT <= 512, "TMEM columns");
    static constexpr uint32_t B_LBO = NOUT * 32;       // bytes between k-groups of the weight image
};

struct RowCoord {
    bool valid;
    int r, j, y, x, sl;
    long long frame;        // r*T + t
};

__device__ __forceinline__ int find_class(const GatherParams& p, int tile) {
    int c = 0;
    while (c + 1 < p.nclass && tile >= p.cls[c + 1].tile_begin) ++c;
    return c;
}
__device__ __forceinline__ RowCoord row_coord(const GatherParams& p, const TapClass& cl, int tile, int row) {
    RowCoord rc;
    const int local = tile - cl.tile_begin;
    rc.r = local / cl.tpd;
    rc.j = local - rc.r * cl.tpd;
    const int q = rc.j * kRows + row;
    const int HW = cl.DH * cl.DW;
    rc.valid = q < p.T * HW;
    const int t = q / HW, rem = q - t * HW;
    rc.y = rem / cl.DW;
    rc.x = rem - rc.y * cl.DW;
    rc.frame = (long long)rc.r * p.T + t;
    rc.sl = rc.r % p.k;
    return rc;
}

template <int CSRC, int NOUT>
__global__ void __launch_bounds__(kThreadsTc, 1)
conv_tc_gather_kernel(const __grid_constant__ GatherParams p) {
    using C = GatherCfg<CSRC, NOUT>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 5];   // full[8] empty[8] tfull[2] tempty[2] wbar
    __shared__ uint32_t tmem_slot;
    __shared__ float2 red[NOUT];                                 // second half-tile's column sums

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.stages;
    uint8_t* wsm = smem;
    uint8_t* ring = wsm + (size_t)p.nw * C::WTAP;
    float* ost = reinterpret_cast<float*>(ring + (size_t)S * C::STAGE);
    long long* rowdst = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(ost) + C::OST_BYTES);
    int* rowflag = reinterpret_cast<int*>(rowdst + kRows);     // quad mode: bit 0 / 1 = row 2y+1 / column 2x+1 exists
    float* aff = reinterpret_cast<float*>(rowflag + kRows);    // [2][k*CSRC] scale | shift
    const int kc = p.k * CSRC;

    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]),
                   tfull0 = smem_u32(&bars[2 * kMaxStages]), tempty0 = smem_u32(&bars[2 * kMaxStages + 2]),
                   wbar = smem_u32(&bars[2 * kMaxStages + 4]);
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full0 + 8 * s, 4); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, kEpi / 32); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)), "r"(C::TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (p.scale != nullptr)
        for (int i = tid; i < kc; i += kThreadsTc) { aff[i] = p.scale[i]; aff[kc + i] = p.shift[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp >= 5) {
        // ======================= producers: gather -> bf16 hi/lo core matrices =======================
        // Up to kGroups groups of four warps, one thread per row.  The flat sequence of (tile, tap) items is cut into
        // batches of D items; a group owns every ngroups-th batch, issues all of its loads, then converts - so
        // the other groups' loads are in flight while one converts and stores.  ngroups * D <= S (ring depth):
        // a group must not get a whole ring round ahead of the group that fills the same slot one round earlier,
        // or its parity wait on the slot's empty barrier aliases (it would see the phase before last as complete).
        const int ptid = tid - (kEpi + 32);
        const int prow = ptid & (kRows - 1), grp = ptid >> 7;
        const uint32_t ring_u = smem_u32(ring);
        int tile = blockIdx.x, tap = 0;
        int cls_tile = -1, rc_tile = -1;
        const TapClass* cl = &p.cls[0];
        RowCoord rc = {};
        int y0 = 0, x0 = 0;
        auto advance = [&]() {   // to the next (tile, tap) item of this CTA
            if (cls_tile != tile) { cl = &p.cls[find_class(p, tile)]; cls_tile = tile; }
            if (++tap >= cl->ntaps) { tap = 0; tile += gridDim.x; }
        };
        for (int d = 0; d < grp * C::D; ++d)
            if (tile < p.ntiles) advance();
        const float4* src4 = reinterpret_cast<const float4*>(p.src);
        const uint32_t row_off = (uint32_t)(prow >> 3) * 128u + (uint32_t)(prow & 7) * 16u;
        int stage = (grp * C::D) % S, round = (grp * C::D) / S;     // ring position of this group's next item
        const int NG = p.ngroups;
        if (grp >= NG) tile = p.ntiles;                            // this group is not used (shallow ring)
        while (tile < p.ntiles) {
            float4 v[C::D][C::G][2];
            bool act[C::D], ok[C::D];
            int sl[C::D];
#pragma unroll
            for (int d = 0; d < C::D; ++d) {
                act[d] = tile < p.ntiles;
                ok[d] = false; sl[d] = 0;
                if (act[d]) {
                    if (cls_tile != tile) { cl = &p.cls[find_class(p, tile)]; cls_tile = tile; }
                    if (rc_tile != tile) {
                        rc = row_coord(p, *cl, tile, prow);
                        rc_tile = tile;
                        y0 = p.smul * rc.y; x0 = p.smul * rc.x;
                    }
                    const int sy = y0 + cl->offy[tap], sx = x0 + cl->offx[tap];
                    ok[d] = rc.valid && sy >= 0 && sy < p.SH && sx >= 0 && sx < p.SW;
                    sl[d] = rc.sl;
                    if (ok[d]) {
                        const float4* q = src4 + ((size_t)rc.frame * p.SH * p.SW + (size_t)sy * p.SW + sx) * (CSRC / 4);
#pragma unroll
                        for (int g = 0; g < C::G; ++g) {
                            v[d][g][0] = __ldg(q + 2 * g);
                            v[d][g][1] = __ldg(q + 2 * g + 1);
                        }
                    }
                    advance();
                }
            }
#pragma unroll
            for (int d = 0; d < C::D; ++d) {
                if (act[d]) {
                    if (round > 0) {
                        if (lane == 0) mbar_wait(empty0 + 8 * stage, (round - 1) & 1);
                        __syncwarp();
                    }
                    const uint32_t sbase = ring_u + (uint32_t)stage * C::STAGE + row_off;
                    const float* scp = aff + sl[d] * CSRC;
#pragma unroll
                    for (int g = 0; g < C::G; ++g) {
                        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                        if (ok[d]) {
                            float4 a = v[d][g][0], b = v[d][g][1];
                            if (p.scale != nullptr) {
                                const float4 s0 = *reinterpret_cast<const float4*>(scp + g * 8);
                                const float4 s1 = *reinterpret_cast<const float4*>(scp + g * 8 + 4);
                                const float4 h0 = *reinterpret_cast<const float4*>(scp + kc + g * 8);
                                const float4 h1 = *reinterpret_cast<const float4*>(scp + kc + g * 8 + 4);
                                a.x = fmaf(a.x, s0.x, h0.x); a.y = fmaf(a.y, s0.y, h0.y);
                                a.z = fmaf(a.z, s0.z, h0.z); a.w = fmaf(a.w, s0.w, h0.w);
                                b.x = fmaf(b.x, s1.x, h1.x); b.y = fmaf(b.y, s1.y, h1.y);
                                b.z = fmaf(b.z, s1.z, h1.z); b.w = fmaf(b.w, s1.w, h1.w);
                            }
                            split2p(a.x, a.y, hi.x, lo.x); split2p(a.z, a.w, hi.y, lo.y);
                            split2p(b.x, b.y, hi.z, lo.z); split2p(b.z, b.w, hi.w, lo.w);
                        }
                        st_shared16(sbase + (uint32_t)g * 4096u, hi);
                        st_shared16(sbase + (uint32_t)g * 4096u + 2048u, lo);
                    }
                    proxy_fence_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * stage);
                    if (++stage == S) { stage = 0; ++round; }
                }
            }
            stage += (NG - 1) * C::D;                            // the other groups' batches
            while (stage >= S) { stage -= S; ++round; }
            for (int d = 0; d < (NG - 1) * C::D; ++d)
                if (tile < p.ntiles) advance();
        }
    } else if (warp == 4) {
        if (lane == 0) {
            // ======================= MMA issuer =======================
            mbar_expect_tx(wbar, (uint32_t)p.nw * C::WTAP);
            for (int t = 0; t < p.nw; ++t)
                bulk_copy(smem_u32(wsm) + t * C::WTAP, p.wimg + (size_t)t * C::WTAP, C::WTAP, wbar);
            mbar_wait(wbar, 0);
            // f32 accumulator, bf16 x bf16, K-major A and B, M = 128
            constexpr uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kRows >> 4) << 24);
            constexpr uint32_t idesc2 = idesc_base | ((uint32_t)((2 * NOUT) >> 3) << 17);
            constexpr uint32_t idesc1 = idesc_base | ((uint32_t)(NOUT >> 3) << 17);
            const uint32_t ring_u = smem_u32(ring), w_u = smem_u32(wsm);
            int stage = 0, phase = 0, ti = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
                const TapClass& cl = p.cls[find_class(p, tile)];
                const int buf = ti & 1, use = ti >> 1;
                if (use > 0) mbar_wait(tempty0 + 8 * buf, (use - 1) & 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)buf * 2 * NOUT;
                for (int t = 0; t < cl.ntaps; ++t) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = ring_u + (uint32_t)stage * C::STAGE;
                    const uint32_t sb = w_u + (uint32_t)cl.wtap[t] * C::WTAP;
#pragma unroll
                    for (int kk = 0; kk < CSRC / 16; ++kk) {
                        const uint64_t ahi = make_desc(sa + kk * 8192, 4096, 128);
                        const uint64_t alo = make_desc(sa + kk * 8192 + 2048, 4096, 128);
                        const uint64_t b = make_desc(sb + kk * 2 * C::B_LBO, C::B_LBO, 128);
                        umma_bf16(acc, ahi, b, idesc2, (t > 0 || kk > 0) ? 1u : 0u);
                        umma_bf16(acc, alo, b, idesc1, 1u);
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull0 + 8 * buf);
            }
        }
    } else {
        // ======================= epilogue: TMEM -> bias / lrelu -> staging -> global =======================
        const int etid = tid;
        float* orow = ost + etid * C::OST;
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
            const TapClass& cl = p.cls[find_class(p, tile)];
            const RowCoord rc = row_coord(p, cl, tile, etid);
            const int buf = ti & 1, use = ti >> 1;
            mbar_wait(tfull0 + 8 * buf, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * 2 * NOUT;
#pragma unroll
            for (int c0 = 0; c0 < NOUT; c0 += 16) {
                uint32_t a[16], b[16];
                tmem_ld16(taddr + c0, a);
                tmem_ld16(taddr + NOUT + c0, b);
                tmem_ld_wait();
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = __uint_as_float(a[j]) + __uint_as_float(b[j]);
                    if (p.act) x = lrelu_f(x + __ldg(p.bias + c0 + j));
                    o[j] = rc.valid ? x : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            rowdst[etid] = rc.valid ? ((rc.frame * p.OH + (p.dmul * rc.y + cl.doff_y)) * p.OW +
                                       (p.dmul * rc.x + cl.doff_x))
                                    : -1;
            rowflag[etid] = ((2 * rc.y + 1 < p.OH) ? 1 : 0) | ((2 * rc.x + 1 < p.OW) ? 2 : 0);
            epi_bar();
            if (p.quad == 0) {
                for (int idx = etid; idx < kRows * (NOUT / 4); idx += kEpi) {
                    const int row = idx / (NOUT / 4), c4 = idx - row * (NOUT / 4);
                    const long long d = rowdst[row];
                    if (d >= 0)
                        *reinterpret_cast<float4*>(p.out + (size_t)d * NOUT + c4 * 4) =
                            *reinterpret_cast<const float4*>(ost + row * C::OST + c4 * 4);
                }
            } else {
                // column block cls = (py, px) of a row goes to pixel (2y + py, 2x + px), `quad` channels each
                const int QC = p.quad, q4 = QC / 4;
                for (int idx = etid; idx < kRows * (NOUT / 4); idx += kEpi) {
                    const int row = idx / (NOUT / 4), c4 = idx - row * (NOUT / 4);
                    const int cls = c4 / q4, cc = (c4 - cls * q4) * 4, py = cls >> 1, px = cls & 1;
                    const long long d = rowdst[row];
                    const int fl = rowflag[row];
                    if (d >= 0 && (!py || (fl & 1)) && (!px || (fl & 2)))
                        *reinterpret_cast<float4*>(p.out + (size_t)(d + py * p.OW + px) * QC + cc) =
                            *reinterpret_cast<const float4*>(ost + row * C::OST + c4 * 4);
                }
            }
            float s = 0.f, s2 = 0.f;
            const int c = etid % NOUT, half = etid / NOUT;
            if (p.partial != nullptr && etid < 2 * NOUT) {
                const float* col = ost + (half * 64) * C::OST + c;
#pragma unroll 8
                for (int rr = 0; rr < 64; ++rr) { const float x = col[rr * C::OST]; s += x; s2 = fmaf(x, x, s2); }
                if (half == 1) red[c] = make_float2(s, s2);
            }
            epi_bar();
            if (p.partial != nullptr && etid < NOUT) {
                const float2 o = red[c];
                const int chunk = (rc.r / p.k) * cl.tpd + rc.j;
                p.partial[((size_t)rc.sl * p.nchunk + chunk) * NOUT + c] = make_float2(s + o.x, s2 + o.y);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TCOLS));
    }
}

// weight images: fwd image (tap, k = ci, n = co) and input-gradient image (tap, k = co, n = ci); for
// each (tap, k-group of 8) the n-groups of the hi parts are followed by those of the lo parts, so that
// [Whi | Wlo] is ONE N = 2*NOUT operand:  off = (((tap*KG + k/8) * (2*NG) + hl*NG + n/8) * 128 + (n%8)*16 + (k%8)*2
__global__ void conv_tc_pack_w(const float* __restrict__ W, int CIN, int COUT, uint8_t* __restrict__ img_f,
                               uint8_t* __restrict__ img_x) {
    const int total = 9 * CIN * COUT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int o = idx % COUT, c = (idx / COUT) % CIN, t = idx / (COUT * CIN);
        const float w = W[idx];
        const bf16 h = __float2bfloat16_rn(w);
        const bf16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        {
            const size_t base = ((size_t)(t * (CIN / 8) + c / 8) * (2 * (COUT / 8)) + o / 8) * 128 + (o % 8) * 16 +
                                (c % 8) * 2;
            *reinterpret_cast<bf16*>(img_f + base) = h;
            *reinterpret_cast<bf16*>(img_f + base + (size_t)(COUT / 8) * 128) = l;
        }
        {
            const size_t base = ((size_t)(t * (COUT / 8) + o / 8) * (2 * (CIN / 8)) + c / 8) * 128 + (c % 8) * 16 +
                                (o % 8) * 2;
            *reinterpret_cast<bf16*>(img_x + base) = h;
            *reinterpret_cast<bf16*>(img_x + base + (size_t)(CIN / 8) * 128) = l;
        }
    }
}

// Weight images of the quad input-gradient form: tap (iy, ix) in {0,1}^2 <-> source offset (dys[iy], dxs[ix]) on the
// dZ grid, k = co, n = (py*2 + px)*CIN + ci with W[ky, kx, ci, co] for ky = py + PT - 2*dy, kx = px + PL - 2*dx
// (zero when that tap does not reach the class); same core-matrix layout as conv_tc_pack_w.
__global__ void conv_tc_pack_wq(const float* __restrict__ W, int CIN, int COUT, int PT, int PL,
                                uint8_t* __restrict__ img) {
    const int NOUT = 4 * CIN, total = 4 * COUT * NOUT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx % NOUT, o = (idx / NOUT) % COUT, ti = idx / (NOUT * COUT);
        const int cls = n / CIN, c = n - cls * CIN, py = cls >> 1, px = cls & 1;
        const int dy = (ti >> 1) ? (PT ? 1 : -1) : 0, dx = (ti & 1) ? (PL ? 1 : -1) : 0;
        const int ky = py + PT - 2 * dy, kx = px + PL - 2 * dx;
        float w = 0.f;
        if (ky >= 0 && ky < 3 && kx >= 0 && kx < 3) w = W[((size_t)(ky * 3 + kx) * CIN + c) * COUT + o];
        const bf16 h = __float2bfloat16_rn(w);
        const bf16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        const size_t base = ((size_t)(ti * (COUT / 8) + o / 8) * (2 * (NOUT / 8)) + n / 8) * 128 + (n % 8) * 16 + (o % 8) * 2;
        *reinterpret_cast<bf16*>(img + base) = h;
        *reinterpret_cast<bf16*>(img + base + (size_t)(NOUT / 8) * 128) = l;
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
struct DwParams {
    const float* in;         // [N, IH, IW, CIN]
    const float* dz;         // [N, OH, OW, COUT]
    const float* scale;      // [k, CIN] or null
    const float* shift;
    float* partial;          // [grid, 9*CIN, COUT]
    int T, k, IH, IW, OH, OW, PT, PL;
    long long npix;
    int pix_per_cta, stages, swap_ls;
    int ngroups;             // active producer groups (<= kGroups, <= stages)
};

template <int CIN, int COUT>
struct DwCfg {
    static constexpr int G = CIN / 8;
    static constexpr int MG = 9 * G;                    // mn-groups of the im2col^T operand
    static constexpr int M = 9 * CIN;
    static constexpr int MT = (M + 127) / 128;          // accumulator tiles
    static constexpr int KS = CIN == 16 ? 2 : 1;        // k16 steps per stage
    static constexpr int PXS = 16 * KS;                 // pixels per stage
    static constexpr int KG = PXS / 8;
    static constexpr int NSUB = 128 / PXS;              // producer threads per pixel (one group of 4 warps per stage)
    static constexpr int NCOMBO = (MG + NSUB - 1) / NSUB;
    static_assert(NSUB >= COUT / 8, "one producer sub-lane per dZ channel group");
    static constexpr uint32_t LBO_A = MG * 128;         // bytes between k-groups (8 pixels)
    static constexpr uint32_t A_HALF = KG * LBO_A;      // hi (or lo) part of the A stage
    static constexpr uint32_t LBO_B = 2 * (COUT / 8) * 128;
    static constexpr uint32_t B_BYTES = KG * LBO_B;
    static constexpr uint32_t STAGE = 2 * A_HALF + B_BYTES;
    static constexpr uint32_t TCOLS = MT * 2 * COUT <= 128 ? 128 : (MT * 2 * COUT <= 256 ? 256 : 512);
    static_assert((MT * 16 - MG) * 128 <= (int)B_BYTES, "the last M tile's over-read stays inside the stage");
    static_assert(MT * 2 * COUT <= 512, "TMEM columns");
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreadsTc, 1)
conv_tc_dw_kernel(const __grid_constant__ DwParams p) {
    using C = DwCfg<CIN, COUT>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];   // full[8] empty[8] accum
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.stages;
    uint8_t* ring = smem;
    float* aff = reinterpret_cast<float*>(ring + (size_t)S * C::STAGE);
    const int kc = p.k * CIN;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]),
                   accum = smem_u32(&bars[2 * kMaxStages]);
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full0 + 8 * s, 4); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)), "r"(C::TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (p.scale != nullptr)
        for (int i = tid; i < kc; i += kThreadsTc) { aff[i] = p.scale[i]; aff[kc + i] = p.shift[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const long long pix0 = (long long)blockIdx.x * p.pix_per_cta;
    long long pix_end = pix0 + p.pix_per_cta;
    if (pix_end > p.npix) pix_end = p.npix;
    const int nst = (int)((pix_end - pix0 + C::PXS - 1) / C::PXS);

    if (warp != 4) {
        // ======================= producers =======================
        // kGroups + 1 groups of four warps (the epilogue warps have nothing to do until the last stage has been
        // multiplied, so they produce too, as group kGroups); a group owns every ngroups-th stage (the others'
        // loads fly while it converts)
        const int grp = warp >= 5 ? (tid - (kEpi + 32)) >> 7 : kGroups, gt = warp >= 5 ? (tid - (kEpi + 32)) & 127 : tid;
        const int pl = gt % C::PXS, sub = gt / C::PXS;
        const int kg = pl >> 3, p8 = pl & 7;
        const float4* in4 = reinterpret_cast<const float4*>(p.in);
        long long pix = pix0 + (long long)grp * C::PXS + pl;
        int ox = (int)(pix % p.OW), oy = (int)((pix / p.OW) % p.OH);
        int n = (int)(pix / ((long long)p.OW * p.OH));
        const uint32_t ring_u = smem_u32(ring);
        const uint32_t a_off = (uint32_t)kg * C::LBO_A + (uint32_t)p8 * 16u;
        const uint32_t b_off = 2 * C::A_HALF + (uint32_t)kg * C::LBO_B + (uint32_t)sub * 128u + (uint32_t)p8 * 16u;
        int stage = grp % S, round = grp / S;
        const int NG = p.ngroups;
        for (int s = grp < NG ? grp : nst; s < nst; s += NG) {
            const bool valid = pix < pix_end;
            const int sl = (n / p.T) % p.k;
            const float4* fin = in4 + (size_t)n * p.IH * p.IW * (CIN / 4);
            float4 v[C::NCOMBO][2];
            bool ok[C::NCOMBO];
#pragma unroll
            for (int i = 0; i < C::NCOMBO; ++i) {
                const int c = sub + i * C::NSUB;
                ok[i] = false;
                if (c < C::MG) {
                    const int tap = c / C::G, g = c % C::G;
                    const int iy = 2 * oy + tap / 3 - p.PT, ix = 2 * ox + tap % 3 - p.PL;
                    ok[i] = valid && iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
                    if (ok[i]) {
                        const float4* q = fin + ((size_t)iy * p.IW + ix) * (CIN / 4) + g * 2;
                        v[i][0] = __ldg(q);
                        v[i][1] = __ldg(q + 1);
                    }
                }
            }
            float4 z[2];
            const bool zok = valid && sub < COUT / 8;
            if (zok) {
                const float4* q = reinterpret_cast<const float4*>(p.dz + (size_t)pix * COUT + sub * 8);
                z[0] = __ldg(q);
                z[1] = __ldg(q + 1);
            }
            if (round > 0) {
                if (lane == 0) mbar_wait(empty0 + 8 * stage, (round - 1) & 1);
                __syncwarp();
            }
            const uint32_t sbase = ring_u + (uint32_t)stage * C::STAGE;
#pragma unroll
            for (int i = 0; i < C::NCOMBO; ++i) {
                const int c = sub + i * C::NSUB;
                if (c < C::MG) {
                    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                    if (ok[i]) {
                        const int g = c % C::G;
                        float4 a = v[i][0], b = v[i][1];
                        if (p.scale != nullptr) {
                            const float4* scp = reinterpret_cast<const float4*>(aff + sl * CIN + g * 8);
                            const float4* shp = reinterpret_cast<const float4*>(aff + kc + sl * CIN + g * 8);
                            const float4 s0 = scp[0], s1 = scp[1], h0 = shp[0], h1 = shp[1];
                            a.x = fmaf(a.x, s0.x, h0.x); a.y = fmaf(a.y, s0.y, h0.y);
                            a.z = fmaf(a.z, s0.z, h0.z); a.w = fmaf(a.w, s0.w, h0.w);
                            b.x = fmaf(b.x, s1.x, h1.x); b.y = fmaf(b.y, s1.y, h1.y);
                            b.z = fmaf(b.z, s1.z, h1.z); b.w = fmaf(b.w, s1.w, h1.w);
                        }
                        split2p(a.x, a.y, hi.x, lo.x); split2p(a.z, a.w, hi.y, lo.y);
                        split2p(b.x, b.y, hi.z, lo.z); split2p(b.z, b.w, hi.w, lo.w);
                    }
                    st_shared16(sbase + a_off + (uint32_t)c * 128u, hi);
                    st_shared16(sbase + C::A_HALF + a_off + (uint32_t)c * 128u, lo);
                }
            }
            if (sub < COUT / 8) {
                uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                if (zok) {
                    split2p(z[0].x, z[0].y, hi.x, lo.x); split2p(z[0].z, z[0].w, hi.y, lo.y);
                    split2p(z[1].x, z[1].y, hi.z, lo.z); split2p(z[1].z, z[1].w, hi.w, lo.w);
                }
                st_shared16(sbase + b_off, hi);
                st_shared16(sbase + b_off + (COUT / 8) * 128u, lo);
            }
            proxy_fence_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * stage);
            stage += NG;
            while (stage >= S) { stage -= S; ++round; }
            pix += NG * C::PXS;
            ox += NG * C::PXS;
            while (ox >= p.OW) { ox -= p.OW; ++oy; }
            while (oy >= p.OH) { oy -= p.OH; ++n; }
        }
    }
    if (warp == 4) {
        if (lane == 0) {
            // ======================= MMA issuer: MN-major A (im2col^T) and B (dZ^T) =======================
            constexpr uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                            ((uint32_t)(kRows >> 4) << 24);
            constexpr uint32_t idesc2 = idesc_base | ((uint32_t)((2 * COUT) >> 3) << 17);
            constexpr uint32_t idesc1 = idesc_base | ((uint32_t)(COUT >> 3) << 17);
            const uint32_t ring_u = smem_u32(ring);
            const uint32_t a_l = p.swap_ls ? 128u : C::LBO_A, a_s = p.swap_ls ? C::LBO_A : 128u;
            const uint32_t b_l = p.swap_ls ? 128u : C::LBO_B, b_s = p.swap_ls ? C::LBO_B : 128u;
            int stage = 0, phase = 0;
            for (int s = 0; s < nst; ++s) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = ring_u + (uint32_t)stage * C::STAGE, sb = sa + 2 * C::A_HALF;
#pragma unroll
                for (int ks = 0; ks < C::KS; ++ks) {
                    const uint64_t b = make_desc(sb + ks * 2 * C::LBO_B, b_l, b_s);
#pragma unroll
                    for (int mt = 0; mt < C::MT; ++mt) {
                        const uint64_t ahi = make_desc(sa + ks * 2 * C::LBO_A + mt * 2048, a_l, a_s);
                        const uint64_t alo = make_desc(sa + C::A_HALF + ks * 2 * C::LBO_A + mt * 2048, a_l, a_s);
                        const uint32_t acc = tmem_base + (uint32_t)mt * 2 * COUT;
                        umma_bf16(acc, ahi, b, idesc2, (s > 0 || ks > 0) ? 1u : 0u);
                        umma_bf16(acc, alo, b, idesc1, 1u);
                    }
                }
                umma_commit(empty0 + 8 * stage);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum);
        }
    } else if (warp < 4) {
        // ======================= epilogue: this CTA's partial dW tile =======================
        mbar_wait(accum, 0);
        tc_fence_after();
        float* part = p.partial + (size_t)blockIdx.x * C::M * COUT;
#pragma unroll
        for (int mt = 0; mt < C::MT; ++mt) {
            const int m = mt * 128 + tid;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)mt * 2 * COUT;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                uint32_t a[16], b[16];
                tmem_ld16(taddr + c0, a);
                tmem_ld16(taddr + COUT + c0, b);
                tmem_ld_wait();
                if (m < C::M) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(part + (size_t)m * COUT + c0 + j) = make_float4(
                            __uint_as_float(a[j]) + __uint_as_float(b[j]),
                            __uint_as_float(a[j + 1]) + __uint_as_float(b[j + 1]),
                            __uint_as_float(a[j + 2]) + __uint_as_float(b[j + 2]),
                            __uint_as_float(a[j + 3]) + __uint_as_float(b[j + 3]));
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TCOLS));
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient of the RGB input layer (CIN = 3, u8 frames): dW[27, COUT] = sum_pixels patch[27] x dZ[COUT]
// ------------------------------------------------------------------------------------------------
// Same structure as conv_tc_dw_kernel with M = 27 patch elements (4 mn-groups of one 128-row accumulator; the
// other 12 groups of the descriptor read whatever follows in the stage and produce rows nobody reads).  The
// frame bytes 0..255 are exact in bf16, so A has no lo part and one tcgen05.mma `A x [dZhi|dZlo]` per k16 step
// is the full fp32-equivalent product.  Stage = 32 pixels; thread (pixel, s) of a producer group gathers the
// eight patch bytes 8s..8s+7 of its pixel and the s-th 16-byte chunk of its dZ row.
struct Dw3Params {
    const uint8_t* in;       // [N, IH, IW, 3]
    const float* dz;         // [N, OH, OW, COUT]
    float* partial;          // [grid, 27, COUT]
    int IH, IW, OH, OW, PT, PL;
    long long npix;
    int pix_per_cta, stages;
};
template <int COUT>
struct Dw3Cfg {
    static constexpr int PXS = 32, KG = 4, KS = 2;
    static constexpr uint32_t LBO_A = 4 * 128;                     // 4 mn-groups (32 patch rows) per k-group
    static constexpr uint32_t A_BYTES = KG * LBO_A;
    static constexpr uint32_t LBO_B = 2 * (COUT / 8) * 128;
    static constexpr uint32_t B_BYTES = KG * LBO_B;
    static constexpr uint32_t STAGE = A_BYTES + B_BYTES;
    static constexpr uint32_t TCOLS = 2 * COUT <= 32 ? 32 : 64;
    static_assert(COUT == 16, "four 16-byte dZ chunks per pixel, one per producer sub-lane");
    static_assert((16 - 4) * 128 <= (int)B_BYTES, "the accumulator tile's over-read stays inside the stage");
};

template <int COUT>
__global__ void __launch_bounds__(kThreadsTc, 1)
conv_tc_dw3_kernel(const __grid_constant__ Dw3Params p) {
    using C = Dw3Cfg<COUT>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.stages;
    uint8_t* ring = smem;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kMaxStages]),
                   accum = smem_u32(&bars[2 * kMaxStages]);
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full0 + 8 * s, 4); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)), "r"(C::TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // the whole ring starts as zeros: rows 27..31 of A stay zero forever, stale bits never reach a valid row
    for (int i = tid; i < (int)(S * C::STAGE / 16); i += kThreadsTc)
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    proxy_fence_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const long long pix0 = (long long)blockIdx.x * p.pix_per_cta;
    long long pix_end = pix0 + p.pix_per_cta;
    if (pix_end > p.npix) pix_end = p.npix;
    const int nst = (int)((pix_end - pix0 + C::PXS - 1) / C::PXS);

    constexpr int NG3 = kGroups + 1;   // the epilogue warps produce too (group kGroups) until the last stage
    if (warp != 4) {
        const int grp = warp >= 5 ? (tid - (kEpi + 32)) >> 7 : kGroups, gt = warp >= 5 ? (tid - (kEpi + 32)) & 127 : tid;
        const int pl = gt & 31, sub = gt >> 5;          // a warp = the 32 pixels of the stage, sub-lane = warp in group
        const int kg = pl >> 3, p8 = pl & 7;
        long long pix = pix0 + (long long)grp * C::PXS + pl;
        int ox = (int)(pix % p.OW), oy = (int)((pix / p.OW) % p.OH);
        int n = (int)(pix / ((long long)p.OW * p.OH));
        const uint32_t ring_u = smem_u32(ring);
        const uint32_t a_off = (uint32_t)kg * C::LBO_A + (uint32_t)sub * 128u + (uint32_t)p8 * 16u;
        const uint32_t b_off = C::A_BYTES + (uint32_t)kg * C::LBO_B + (uint32_t)(sub >> 1) * 128u +
                               (uint32_t)p8 * 16u + (uint32_t)(sub & 1) * 8u;
        // U stages per iteration: all of their byte loads are issued before the first conversion
        // (kGroups * U <= ring depth, see the producers of conv_tc_gather_kernel)
        constexpr int U = 2;
        static_assert(NG3 * U <= kMaxStages, "a producer group must stay within one ring round of the others");
        int stage = (grp * U) % S, round = (grp * U) / S;
        pix = pix0 + (long long)grp * U * C::PXS + pl;
        ox = (int)(pix % p.OW); oy = (int)((pix / p.OW) % p.OH);
        n = (int)(pix / ((long long)p.OW * p.OH));
        for (int s0 = grp * U; s0 < nst; s0 += NG3 * U) {
            float x[U][8];
            float4 z[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool valid = pix < pix_end;     // (a stage past nst has no valid pixel either)
                const uint8_t* fin = p.in + (size_t)n * p.IH * p.IW * 3;
                const int iy0 = 2 * oy - p.PT, ix0 = 2 * ox - p.PL;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int el = sub * 8 + e;             // patch element (ky, kx, ci) = (el / 9, (el % 9) / 3, el % 3)
                    const int iy = iy0 + el / 9, ix = ix0 + (el % 9) / 3;
                    const bool ok = valid && el < 27 && iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
                    x[u][e] = ok ? (float)__ldg(fin + ((size_t)iy * p.IW + ix) * 3 + el % 3) : 0.f;
                }
                z[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) z[u] = __ldg(reinterpret_cast<const float4*>(p.dz + (size_t)pix * COUT) + sub);
                pix += C::PXS;
                ox += C::PXS;
                while (ox >= p.OW) { ox -= p.OW; ++oy; }
                while (oy >= p.OH) { oy -= p.OH; ++n; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (s0 + u < nst) {
                    if (round > 0) {
                        if (lane == 0) mbar_wait(empty0 + 8 * stage, (round - 1) & 1);
                        __syncwarp();
                    }
                    const uint32_t sbase = ring_u + (uint32_t)stage * C::STAGE;
                    uint4 hi;
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.x) : "f"(x[u][1]), "f"(x[u][0]));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.y) : "f"(x[u][3]), "f"(x[u][2]));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.z) : "f"(x[u][5]), "f"(x[u][4]));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi.w) : "f"(x[u][7]), "f"(x[u][6]));
                    st_shared16(sbase + a_off, hi);
                    uint2 zh, zl;
                    split2p(z[u].x, z[u].y, zh.x, zl.x);
                    split2p(z[u].z, z[u].w, zh.y, zl.y);
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sbase + b_off), "r"(zh.x), "r"(zh.y) : "memory");
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sbase + b_off + (COUT / 8) * 128u), "r"(zl.x),
                                 "r"(zl.y) : "memory");
                    proxy_fence_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * stage);
                    if (++stage == S) { stage = 0; ++round; }
                }
            }
            stage += (NG3 - 1) * U;
            while (stage >= S) { stage -= S; ++round; }
            pix += (long long)(NG3 - 1) * U * C::PXS;
            ox += (NG3 - 1) * U * C::PXS;
            while (ox >= p.OW) { ox -= p.OW; ++oy; }
            while (oy >= p.OH) { oy -= p.OH; ++n; }
        }
    }
    if (warp == 4) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(kRows >> 4) << 24) | ((uint32_t)((2 * COUT) >> 3) << 17);
            const uint32_t ring_u = smem_u32(ring);
            int stage = 0, phase = 0;
            for (int s = 0; s < nst; ++s) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = ring_u + (uint32_t)stage * C::STAGE, sb = sa + C::A_BYTES;
#pragma unroll
                for (int ks = 0; ks < C::KS; ++ks) {
                    const uint64_t a = make_desc(sa + ks * 2 * C::LBO_A, C::LBO_A, 128);
                    const uint64_t b = make_desc(sb + ks * 2 * C::LBO_B, C::LBO_B, 128);
                    umma_bf16(tmem_base, a, b, idesc, (s > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * stage);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum);
        }
    } else if (warp < 4) {
        mbar_wait(accum, 0);
        tc_fence_after();
        float* part = p.partial + (size_t)blockIdx.x * 27 * COUT;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t a[16], b[16];
        tmem_ld16(taddr, a);
        tmem_ld16(taddr + COUT, b);
        tmem_ld_wait();
        if (tid < 27) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(part + (size_t)tid * COUT + j) = make_float4(
                    __uint_as_float(a[j]) + __uint_as_float(b[j]), __uint_as_float(a[j + 1]) + __uint_as_float(b[j + 1]),
                    __uint_as_float(a[j + 2]) + __uint_as_float(b[j + 2]), __uint_as_float(a[j + 3]) + __uint_as_float(b[j + 3]));
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TCOLS));
    }
}

// dW[i] += sum_blk partial[blk, i]   (fixed order, double accumulation)
__global__ void conv_tc_dw_reduce(const float* __restrict__ partial, int nblk, int n, float* __restrict__ dW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * n + i];
    dW[i] += (float)s;
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t img_bytes(const ConvGeo& g) { return (size_t)16 * g.CIN * g.COUT * 4; }   // 9 taps, or 4 quad taps x 4 classes
inline int fwd_tpd(const ConvGeo& g) { return cdiv((long long)g.T * g.OH * g.OW, kRows); }
inline int fwd_nchunk(const ConvGeo& g) { return (g.N / g.T / g.k) * fwd_tpd(g); }
inline size_t stat_bytes(const ConvGeo& g) { return ((size_t)g.k * fwd_nchunk(g) * g.COUT * 2 + (size_t)g.k * g.COUT) * sizeof(float); }

template <int CSRC, int NOUT>
int launch_gather(cudaStream_t st, GatherParams& p) {
    using C = GatherCfg<CSRC, NOUT>;
    const size_t fixed = (size_t)p.nw * C::WTAP + C::OST_BYTES + kRows * (sizeof(long long) + sizeof(int)) +
                         (size_t)2 * p.k * CSRC * sizeof(float);
    D2P_REQUIRE(fixed + 2 * C::STAGE <= kSmemLimit, "conv tc: shared memory (k=%d)", p.k);
    int S = (int)((kSmemLimit - fixed) / C::STAGE);
    if (S > kMaxStages) S = kMaxStages;
    if (const char* e = getenv("D2P_CONV_TC_STAGES")) {   // developer switch: ring depth of the gather kernel
        const int v = atoi(e);
        if (v >= 2 && v < S) S = v;
    }
    p.stages = S;
    p.ngroups = S / C::D < kGroups ? S / C::D : kGroups;
    D2P_REQUIRE(p.ngroups >= 1, "conv tc: ring too shallow");
    const size_t smem = fixed + (size_t)S * C::STAGE;
    D2P_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_gather_kernel<CSRC, NOUT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    if (const char* e = getenv("D2P_CONV_TC_GRID")) {   // developer switch: CTAs of the gather kernel
        const int v = atoi(e);
        if (v > 0 && v < grid) grid = v;
    }
    conv_tc_gather_kernel<CSRC, NOUT><<<grid, kThreadsTc, smem, st>>>(p);
    D2P_CHECK_LAUNCH();
    return 0;
}

int dispatch_gather(cudaStream_t st, int CSRC, int NOUT, GatherParams& p) {
#define D2P_GATHER(A_, B_) if (CSRC == A_ && NOUT == B_) return launch_gather<A_, B_>(st, p)
    D2P_GATHER(16, 16); D2P_GATHER(16, 32); D2P_GATHER(16, 48);
    D2P_GATHER(32, 16); D2P_GATHER(32, 32); D2P_GATHER(32, 48);
    D2P_GATHER(48, 16); D2P_GATHER(48, 32); D2P_GATHER(48, 48);
    D2P_GATHER(32, 64); D2P_GATHER(48, 128);       // quad input gradients of 16- and 32-channel inputs
#undef D2P_GATHER
    return fail(D2P_ERR_ARG, "conv tc: unsupported channel counts %d -> %d", CSRC, NOUT);
}

template <int CIN, int COUT>
int launch_dw(cudaStream_t st, DwParams& p, int* grid_out) {
    using C = DwCfg<CIN, COUT>;
    const size_t fixed = (size_t)2 * p.k * CIN * sizeof(float);
    int S = (int)((kSmemLimit - fixed) / C::STAGE);
    if (S > kMaxStages) S = kMaxStages;
    D2P_REQUIRE(S >= 2, "conv tc dw: shared memory");
    p.stages = S;
    p.ngroups = S < kGroups + 1 ? S : kGroups + 1;
    const long long nchunks = (p.npix + C::PXS - 1) / C::PXS;
    long long grid = nchunks < kNumSMs ? nchunks : kNumSMs;
    const long long per = (nchunks + grid - 1) / grid * C::PXS;
    grid = (p.npix + per - 1) / per;
    p.pix_per_cta = (int)per;
    *grid_out = (int)grid;
    const size_t smem = fixed + (size_t)S * C::STAGE;
    D2P_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_dw_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    conv_tc_dw_kernel<CIN, COUT><<<(int)grid, kThreadsTc, smem, st>>>(p);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace

int conv_tc_mode() { return tc_available() ? g_conv_tc_mode : 0; }

bool conv_tc_supported(const ConvGeo& g) {
    auto okc = [](int c) { return c == 16 || c == 32 || c == 48; };
    if (!okc(g.CIN) || !okc(g.COUT)) return false;
    if (g.N % g.T != 0 || (g.N / g.T) % g.k != 0) return false;
    if ((size_t)g.k * 48 * 2 * sizeof(float) > 16 * 1024) return false;
    if ((long long)g.N * g.IH * g.IW >= (1LL << 31) / 2) return false;
    return true;
}

size_t conv_tc_ws_bytes(const ConvGeo& g) {
    return 2 * al(img_bytes(g)) + al(stat_bytes(g)) +
           al((size_t)kNumSMs * 9 * g.CIN * g.COUT * sizeof(float));
}

static int pack_images(cudaStream_t st, const ConvGeo& g, const float* W, uint8_t* ws) {
    const int total = 9 * g.CIN * g.COUT;
    conv_tc_pack_w<<<cdiv(total, 256), 256, 0, st>>>(W, g.CIN, g.COUT, ws, ws + al(img_bytes(g)));
    D2P_CHECK_LAUNCH();
    return 0;
}

int conv_tc_fwd(cudaStream_t st, const ConvGeo& g, const float* in, const float* scale, const float* shift,
                const float* W, const float* bias, float* out, int training, int* nchunk, float2** partial,
                void* ws, size_t ws_bytes) {
    D2P_REQUIRE(ws_bytes >= conv_tc_ws_bytes(g), "conv tc fwd: workspace");
    uint8_t* wsb = (uint8_t*)ws;
    D2P_TRY(pack_images(st, g, W, wsb));
    GatherParams p{};
    p.src = in; p.out = out; p.wimg = wsb; p.bias = bias; p.scale = scale; p.shift = shift;
    p.partial = training ? (float2*)(wsb + 2 * al(img_bytes(g))) : nullptr;
    p.R = g.N / g.T; p.T = g.T; p.k = g.k;
    p.SH = g.IH; p.SW = g.IW; p.OH = g.OH; p.OW = g.OW;
    p.smul = 2; p.dmul = 1; p.act = 1;
    p.nw = 9;
    p.nclass = 1;
    TapClass& c = p.cls[0];
    c.tile_begin = 0; c.tpd = fwd_tpd(g); c.DH = g.OH; c.DW = g.OW; c.doff_y = c.doff_x = 0; c.ntaps = 9;
    for (int t = 0; t < 9; ++t) {
        c.offy[t] = (signed char)(t / 3 - g.PT); c.offx[t] = (signed char)(t % 3 - g.PL); c.wtap[t] = (signed char)t;
    }
    p.ntiles = p.R * c.tpd;
    p.nchunk = fwd_nchunk(g);
    *nchunk = p.nchunk;
    *partial = p.partial;
    return dispatch_gather(st, g.CIN, g.COUT, p);
}

int conv_tc_dx(cudaStream_t st, const ConvGeo& g, const float* dZ, const float* W, float* dX, void* ws,
               size_t ws_bytes) {
    D2P_REQUIRE(ws_bytes >= conv_tc_ws_bytes(g), "conv tc dx: workspace");
    uint8_t* wsb = (uint8_t*)ws;
    D2P_TRY(pack_images(st, g, W, wsb));
    GatherParams p{};
    p.src = dZ; p.out = dX; p.wimg = wsb + al(img_bytes(g));
    p.R = g.N / g.T; p.T = g.T; p.k = g.k;
    p.SH = g.OH; p.SW = g.OW; p.OH = g.IH; p.OW = g.IW;
    p.smul = 1; p.dmul = 2; p.act = 0;
    p.nw = 9;
    // developer switch: bit 0 = 16-channel inputs, bit 1 = 32-channel inputs
    int quad_mask = 3;
    if (const char* e = getenv("D2P_CONV_QUAD_MASK")) quad_mask = atoi(e);
    if (g_conv_tc_quad && ((g.COUT == 32 && g.CIN == 16 && (quad_mask & 1)) || (g.COUT == 48 && g.CIN == 32 && (quad_mask & 2)))) {
        // one stride-1 product over the dZ grid with 2x2 taps produces all four parity classes of a 2x2 input
        // block at once: every dZ pixel is gathered 4 times instead of 9, and there are 4x fewer tiles
        const int total = 4 * g.COUT * 4 * g.CIN;
        conv_tc_pack_wq<<<cdiv(total, 256), 256, 0, st>>>(W, g.CIN, g.COUT, g.PT, g.PL, wsb + al(img_bytes(g)));
        D2P_CHECK_LAUNCH();
        p.quad = g.CIN; p.nw = 4;
        p.nclass = 1;
        TapClass& c = p.cls[0];
        c.tile_begin = 0; c.DH = (g.IH + 1) / 2; c.DW = (g.IW + 1) / 2; c.doff_y = c.doff_x = 0; c.ntaps = 4;
        c.tpd = cdiv((long long)g.T * c.DH * c.DW, kRows);
        for (int ti = 0; ti < 4; ++ti) {
            c.offy[ti] = (signed char)((ti >> 1) ? (g.PT ? 1 : -1) : 0);
            c.offx[ti] = (signed char)((ti & 1) ? (g.PL ? 1 : -1) : 0);
            c.wtap[ti] = (signed char)ti;
        }
        p.ntiles = p.R * c.tpd; p.nchunk = 0;
        return dispatch_gather(st, g.COUT, 4 * g.CIN, p);
    }
    int nc = 0, tiles = 0;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const int CH = (g.IH - py + 1) / 2, CW = (g.IW - px + 1) / 2;
            if (CH <= 0 || CW <= 0) continue;
            TapClass& c = p.cls[nc];
            c.tile_begin = tiles; c.DH = CH; c.DW = CW; c.doff_y = py; c.doff_x = px; c.ntaps = 0;
            c.tpd = cdiv((long long)g.T * CH * CW, kRows);
            for (int ky = 0; ky < 3; ++ky) {
                if ((py + g.PT - ky) & 1) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    if ((px + g.PL - kx) & 1) continue;
                    c.offy[c.ntaps] = (signed char)((py + g.PT - ky) / 2);
                    c.offx[c.ntaps] = (signed char)((px + g.PL - kx) / 2);
                    c.wtap[c.ntaps] = (signed char)(ky * 3 + kx);
                    ++c.ntaps;
                }
            }
            D2P_REQUIRE(c.ntaps > 0, "conv tc dx: empty parity class");
            tiles += p.R * c.tpd;
            ++nc;
        }
    p.nclass = nc; p.ntiles = tiles; p.nchunk = 0;
    return dispatch_gather(st, g.COUT, g.CIN, p);
}

int conv_tc_dw(cudaStream_t st, const ConvGeo& g, const float* in, const float* scale, const float* shift,
               const float* dZ, float* dW, void* ws, size_t ws_bytes) {
    D2P_REQUIRE(ws_bytes >= conv_tc_ws_bytes(g), "conv tc dw: workspace");
    uint8_t* wsb = (uint8_t*)ws;
    float* partial = (float*)(wsb + 2 * al(img_bytes(g)) + al(stat_bytes(g)));
    DwParams p{};
    p.in = in; p.dz = dZ; p.scale = scale; p.shift = shift; p.partial = partial;
    p.T = g.T; p.k = g.k; p.IH = g.IH; p.IW = g.IW; p.OH = g.OH; p.OW = g.OW; p.PT = g.PT; p.PL = g.PL;
    p.npix = (long long)g.N * g.OH * g.OW;
    p.swap_ls = g_conv_tc_swap;
    int grid = 0;
    if (g.CIN == 16 && g.COUT == 32) D2P_TRY((launch_dw<16, 32>(st, p, &grid)));
    else if (g.CIN == 16 && g.COUT == 16) D2P_TRY((launch_dw<16, 16>(st, p, &grid)));
    else if (g.CIN == 32 && g.COUT == 48) D2P_TRY((launch_dw<32, 48>(st, p, &grid)));
    else if (g.CIN == 48 && g.COUT == 48) D2P_TRY((launch_dw<48, 48>(st, p, &grid)));
    else return fail(D2P_ERR_ARG, "conv tc dw: unsupported channel counts %d -> %d", g.CIN, g.COUT);
    const int n = 9 * g.CIN * g.COUT;
    conv_tc_dw_reduce<<<cdiv(n, 256), 256, 0, st>>>(partial, grid, n, dW);
    D2P_CHECK_LAUNCH();
    return 0;
}

size_t conv_tc_dw3_ws_bytes(const ConvGeo& g) { return al((size_t)kNumSMs * 27 * g.COUT * sizeof(float)); }
bool conv_tc_dw3_supported(const ConvGeo& g) {
    return g.CIN == 3 && g.COUT == 16 && (long long)g.N * g.IH * g.IW * 3 < (1LL << 31);
}
// RGB input layer (u8 frames): dW += im2col(frames)^T dZ
int conv_tc_dw3(cudaStream_t st, const ConvGeo& g, const uint8_t* in, const float* dZ, float* dW, void* ws,
                size_t ws_bytes) {
    D2P_REQUIRE(ws_bytes >= conv_tc_dw3_ws_bytes(g) && conv_tc_dw3_supported(g), "conv tc dw3: workspace / shape");
    using C = Dw3Cfg<16>;
    Dw3Params p{};
    p.in = in; p.dz = dZ; p.partial = (float*)ws;
    p.IH = g.IH; p.IW = g.IW; p.OH = g.OH; p.OW = g.OW; p.PT = g.PT; p.PL = g.PL;
    p.npix = (long long)g.N * g.OH * g.OW;
    p.stages = kMaxStages;
    const long long nchunks = (p.npix + C::PXS - 1) / C::PXS;
    long long grid = nchunks < kNumSMs ? nchunks : kNumSMs;
    const long long per = (nchunks + grid - 1) / grid * C::PXS;
    grid = (p.npix + per - 1) / per;
    p.pix_per_cta = (int)per;
    const size_t smem = (size_t)p.stages * C::STAGE;
    D2P_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_dw3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_dw3_kernel<16><<<(int)grid, kThreadsTc, smem, st>>>(p);
    D2P_CHECK_LAUNCH();
    const int nw = 27 * g.COUT;
    conv_tc_dw_reduce<<<cdiv(nw, 256), 256, 0, st>>>(p.partial, (int)grid, nw, dW);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace d2p

extern "C" int d2p_conv_set_tc(int mode) {
    const int old = d2p::g_conv_tc_mode | (d2p::g_conv_tc_swap << 8);
    d2p::g_conv_tc_mode = mode & 7;
    d2p::conv_set_rgb((mode & 16) ? 0 : 1);
    d2p::g_conv_tc_quad = (mode & 32) ? 0 : 1;
    d2p::g_conv_tc_swap = (mode >> 8) & 1;
    return old;
}
