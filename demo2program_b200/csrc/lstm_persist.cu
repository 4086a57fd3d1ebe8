// Persistent weight-stationary LSTM recurrence (tcgen05 / TMEM), forward and backward.
//
// One cooperative launch runs ALL T steps of tf.nn.dynamic_rnn(BasicLSTMCell) /
// the teacher-forced decoder cell loop (reference models/model_full.py:244-258,
// 265-277, 465-471) instead of one launch per step:
//
//  * grid = (32 column CTAs) x (row tiles of 128 sequence rows), one CTA per SM.
//    Every CTA keeps its slab of the recurrent weight resident in shared memory for
//    the whole sequence: forward 64 gate columns (i, j, f, o of 16 hidden units) x
//    K = 512 as bf16 hi/lo = 128 KB; backward a 64-unit x 512-gate-column slab of
//    Wh^T, also 128 KB.  Only the 128-row operand (h_{t-1}, resp. dZ_t) streams
//    through a 2 x 32 KB bulk-copy ring each step.
//  * the cell state c, the carried h (forward) and the dc / dh carries (backward)
//    of the thread's (row, 4 hidden units) live in REGISTERS across all steps.
//  * steps are separated by a release/acquire counter barrier among the 32 CTAs
//    of one row tile (row tiles never exchange data), not by kernel boundaries:
//    h_t is written in packed operand form with generic stores, published with
//    fence.proxy.async + a gpu-scope release, and fetched by the next step's
//    cp.async.bulk after an acquire + proxy fence.
//  * forward epilogue = BasicLSTMCell + dynamic_rnn masking, exactly as the
//    per-step kernel (lstm_tc.cu).  Backward: phase P (element-wise, all CTAs,
//    sums the 4 split-K partial slabs in fixed order, writes dZ_t as fp32 and as
//    packed operand), barrier, phase G (partial sums of dZ_t * Wh^T for one
//    K-chunk = one gate), barrier.
//
// Spin waits are bounded (about 2 s of SM clock): a barrier that never completes sets an
// error word instead of hanging the device; the host checks it where it can (eager
// calls) and the engine's parity tests would fail loudly on the garbage it leaves.
#include "tc_common.cuh"

namespace d2p {

using namespace tc;
bool tc_available();

namespace {

constexpr int PH = 512;                         // hidden size served by the persistent kernels
constexpr int PNKB = PH / BK;                   // 8 k-blocks of the resident slab
constexpr int PBN = 64, PUPT = PBN / 4;         // 64 accumulator columns = 16 hidden units x 4 gates
constexpr int PTHREADS = 512;
constexpr int PCOLS = 32;                       // column CTAs per row tile
constexpr int PTMEM = 2 * PBN;                  // accumulator columns: [A*Bhi (+Alo*Bhi) | Ahi*Blo]
constexpr int PMAXSLOTS = 8;                    // ring slots (one per k-block when the row tile is small)
constexpr uint32_t PB_BYTES = (PBN / 8) * 2048; // 16 KB: 64 n x 64 k, hi/lo
constexpr size_t P_W_BYTES = (size_t)PNKB * PB_BYTES;     // 128 KB resident slab
constexpr size_t P_RING_BYTES = 96 * 1024;                // streamed operand ring
constexpr size_t P_SMEM = P_W_BYTES + P_RING_BYTES;
constexpr long long P_SPIN_CYCLES = 4000000000LL;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Reader side of a step hand-off.  The WRITERS make their generic stores of the packed operand visible to the
// async proxy (fence.proxy.async before the release, above); the reader only has to order its own CTA's generic
// accesses to the ring (the staged outputs) before the bulk copies that overwrite it - a shared-memory-only
// proxy fence.  The all-state-space form waits for every global store the SM has in flight (measured: 3.8 k
// cycles behind the 57 KB of copy-out stores, 0.4 k without them; tools/probe_persist.py).  D2P_PERSIST_FENCE=1
// (d2p_lstm_set_persistent bit 2) restores the full fence.
__device__ __forceinline__ void proxy_fence_reader(int full) {
    if (full) asm volatile("fence.proxy.async;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Sticky device-side error word: a step barrier that timed out anywhere since the last query
// (d2p_device_error); the per-launch word in the arena is recycled by later kernels.
__device__ unsigned g_persist_sticky_error = 0;

// Wait until *ctr >= target.  err[0] becomes non-zero if any wait in the grid ran out of
// time; from then on every wait returns immediately so that the kernel still terminates.
__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned target, unsigned* err) {
    if (ld_acquire(ctr) >= target) return;
    const long long t0 = clock64();
    int it = 0;
    while (ld_acquire(ctr) < target) {
        if ((++it & 63) == 0) {
            if (ld_acquire(err) != 0u) return;
            if (clock64() - t0 > P_SPIN_CYCLES) { atomicExch(err, 1u); atomicExch(&g_persist_sticky_error, 1u); return; }
        }
    }
}

// timeline probe (developer tool): stamps of CTA (0,0) during step index P_PROBE_STEP
constexpr int P_PROBE_STEP = 5;
// The stamps go to shared memory and are flushed once at the end of the kernel: a global load / store per
// stamp queues behind whatever the SM's memory pipe is draining and distorts the very thing it measures.
__shared__ long long g_pstamps[32];
__device__ __forceinline__ void pstamp(int step, int slot) {
    if (step == P_PROBE_STEP) g_pstamps[slot & 31] = clock64();
}
// after a CTA barrier: CTA (0,0) -> slots 64..95 (0..63 belong to the per-step kernels' probe); every CTA ->
// 32 slots at 256 + 32 * (blockIdx.y * gridDim.x + blockIdx.x) (SM clocks: only differences within a CTA compare)
__device__ __forceinline__ void pstamp_flush() {
    if (threadIdx.x < 32 && g_tc_dbg != nullptr) {
        if (blockIdx.x == 0 && blockIdx.y == 0) g_tc_dbg[64 + threadIdx.x] = g_pstamps[threadIdx.x];
        g_tc_dbg[256 + 32 * (blockIdx.y * gridDim.x + blockIdx.x) + threadIdx.x] = g_pstamps[threadIdx.x];
    }
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void ld4r(const float* p, float* x) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ void st4r(float* p, const float* x) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
}
// MUFU-only forms (ex2.approx + rcp.approx: ~2 ulp, the fp32 rounding level of the cell)
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_p(float x) { return rcp_approx(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_p(float x) { return 1.0f - 2.0f * rcp_approx(1.0f + __expf(2.0f * x)); }
// staging of the per-step outputs in the (idle) operand ring: [array][row][16 units + 4 pad]
constexpr int PSROW = PUPT + 4;
constexpr size_t P_STAGE_BYTES = (size_t)6 * 128 * PSROW * sizeof(float);   // 6 arrays x 128 rows
__device__ __forceinline__ float* stage_ptr(float* stg, int arr, int row, int unit) {
    return stg + ((size_t)arr * BM + row) * PSROW + unit;
}
// 4 consecutive accumulator columns of this thread's TMEM lane (row)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* x) {
    uint32_t v0, v1, v2, v3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(taddr));
    x[0] = __uint_as_float(v0); x[1] = __uint_as_float(v1); x[2] = __uint_as_float(v2); x[3] = __uint_as_float(v3);
}

// "Stacked" operand of a 32-row recurrence (program decoder, B = 32): per 64-wide k-block 8 KB,
//     [m-group' 0..7][k-group 0..7][8 rows x 16 B],   m-group' g < 4: bf16 HI of rows 8g..8g+7,  g >= 4: LO of rows 8(g-4)..
// so that ONE M = 128 tcgen05.mma per k16 against [Bhi | Blo] (N = 128) yields rows 0-31 = Ahi*[Bhi|Blo] and rows
// 32-63 = Alo*[Bhi|Blo]: all four bf16 partial products in 32 instead of 64 MMAs per step (a tile of 32 valid rows
// costs the same ~68 cycles as a full one, M = 64 included - tools/micro/mma_rate.cu).  LBO 128, SBO 1024.
constexpr int P_STACK_ROWS = 32;
__device__ __forceinline__ void store_stacked4(uint8_t* base, int r, int k0, const float* x) {
    uint32_t h[2], l[2];
    split2(x[0], x[1], h[0], l[0]);
    split2(x[2], x[3], h[1], l[1]);
    uint8_t* p = base + (size_t)(k0 >> 6) * 8192 + (size_t)(((r >> 3) * 8 + ((k0 >> 3) & 7)) * 128) + (r & 7) * 16 +
                 (k0 & 7) * 2;
    *reinterpret_cast<uint2*>(p) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(p + 4096) = make_uint2(l[0], l[1]);
}
__global__ void pack_stacked32_kernel(const float* __restrict__ X /*[32, K]*/, int K, uint8_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (row, 4 consecutive k)
    if (idx >= P_STACK_ROWS * (K / 4)) return;
    const int r = idx / (K / 4), k0 = (idx - r * (K / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(X + (size_t)r * K + k0);
    const float x[4] = {v.x, v.y, v.z, v.w};
    store_stacked4(out, r, k0, x);
}

// barriers: full[PMAXSLOTS], empty[PMAXSLOTS], accum, wfull
struct PersistBars {
    uint32_t full0, empty0, accum, wfull;
    bool single;           // the whole streamed tile (8 k-blocks) is one contiguous run: one bulk copy
    int nslots;            // ring slots in use: min(8, ring bytes / bytes of one streamed k-block)
    uint32_t a_bytes;      // bytes of one streamed k-block (whole 8-row groups of the row tile)
};
// ring position of a role lane (producer or MMA issuer), carried across steps
// (bit s of `par` = parity of this lane's next wait on slot s, bit s of `used` = slot s has been
// filled before: per-slot phases, so that a step may start at any slot)
struct RingPos { int slot; uint32_t par; uint32_t used; };
__device__ __forceinline__ void ring_next(RingPos& r, int nslots) {
    if (++r.slot == nslots) r.slot = 0;
}

// Common setup: barriers, TMEM (64 columns), zeroed ring, resident weight slab.
__device__ __forceinline__ uint32_t persist_setup(uint8_t* smem, uint64_t* bars, uint32_t* tmem_slot,
                                                  PersistBars& pb, int rows, size_t a_kb_stride,
                                                  const uint8_t* wsrc, size_t w_kb_stride) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pb.full0 = smem_u32(&bars[0]);
    pb.empty0 = smem_u32(&bars[PMAXSLOTS]);
    pb.accum = smem_u32(&bars[2 * PMAXSLOTS]);
    pb.wfull = smem_u32(&bars[2 * PMAXSLOTS + 1]);
    pb.a_bytes = (uint32_t)((rows + 7) / 8) * 2048u;
    // the MMA reads a full 128-row image (32 KB) from every slot: the last slot must still end inside the ring
    pb.nslots = (int)((P_RING_BYTES - (BM / 8) * 2048) / pb.a_bytes) + 1;
    if (pb.nslots > PMAXSLOTS) pb.nslots = PMAXSLOTS;
    pb.single = pb.nslots == PNKB && a_kb_stride == (size_t)pb.a_bytes;
    if (tid == 0) {
        for (int s = 0; s < 2 * PMAXSLOTS + 2; ++s) mbar_init(pb.full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "r"((uint32_t)PTMEM));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // the MMA always reads 128 rows per slot: rows a partial row tile never loads must be finite
    uint4* ring = reinterpret_cast<uint4*>(smem + P_W_BYTES);
    for (int i = tid; i < (int)(P_RING_BYTES / 16); i += PTHREADS) ring[i] = make_uint4(0, 0, 0, 0);
    proxy_fence();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0 && lane == 0) {
        mbar_expect_tx(pb.wfull, (uint32_t)P_W_BYTES);
        for (int kb = 0; kb < PNKB; ++kb)
            bulk_copy(smem_u32(smem) + kb * PB_BYTES, wsrc + (size_t)kb * w_kb_stride, PB_BYTES, pb.wfull);
    }
    return *tmem_slot;
}

// One pass of the streamed operand over the resident slab.  Column CTAs walk the 8 k-blocks
// in rotated order (rot) so that the 32 CTAs of a row tile do not queue on one L2 line range.
__device__ __forceinline__ void persist_produce(const PersistBars& pb, RingPos& rp, uint32_t sbase,
                                                const uint8_t* a_src, size_t a_kb_stride, int rot,
                                                int kb_first = 0, int kb_end = PNKB) {
    if (pb.single) {
        // small row tile stored without row padding: the 8 k-blocks are one contiguous run that
        // fills the 8 slots in order - one request instead of eight (~300 issue cycles each)
        mbar_expect_tx(pb.full0, PNKB * pb.a_bytes);
        bulk_copy(sbase + (uint32_t)P_W_BYTES, a_src, PNKB * pb.a_bytes, pb.full0);
        return;
    }
    for (int kb = kb_first; kb < kb_end; ++kb) {
        const uint32_t bit = 1u << rp.slot;
        if (pb.nslots < PNKB && (rp.used & bit)) mbar_wait(pb.empty0 + 8 * rp.slot, ((rp.par >> rp.slot) & 1u) ^ 1u);
        const uint32_t bar = pb.full0 + 8 * rp.slot;
        mbar_expect_tx(bar, pb.a_bytes);
        bulk_copy(sbase + (uint32_t)P_W_BYTES + rp.slot * pb.a_bytes,
                  a_src + (size_t)((kb + rot) & (PNKB - 1)) * a_kb_stride, pb.a_bytes, bar);
        rp.par ^= bit;
        rp.used |= bit;
        ring_next(rp, pb.nslots);
    }
}
// tcgen05.mma with a compile-time accumulate flag (no predicate set-up on the issue path)
template <bool ACC>
__device__ __forceinline__ void umma_imm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    if (ACC)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

// The issuing lane cannot run ahead of the tensor pipe (measured: ALU work between two
// tcgen05.mma adds to the MMA time), so the descriptors are formed by adding constants to the
// low word of two base descriptors instead of being rebuilt per MMA.
__device__ __forceinline__ void persist_mma(const PersistBars& pb, RingPos& rp, uint32_t sbase, uint32_t tmem_d,
                                            int rot, bool first_round, bool stacked = false) {
    // streamed operand: standard packed core matrices (hi/lo interleaved, LBO 256, SBO 2048);
    // resident slab: [hi|lo][group][k-group] (LBO 128, SBO 1024) - 16 uniform row groups, so
    // Ahi * [Bhi | Blo] is ONE N = 128 MMA into columns [0,64) | [64,128), then Alo * Bhi (N = 64)
    // accumulates into [0,64).  (Measured: an M = 128 MMA costs ~68 cycles whatever N <= 128.)
    constexpr uint32_t LBO = 256, SBO = 2048, WLBO = 128, WSBO = 1024;
    constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PBN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
    constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * PBN) >> 3) << 17) |
                                  ((uint32_t)(BM >> 4) << 24);
    const uint64_t a_base = make_desc(sbase + (uint32_t)P_W_BYTES, LBO, SBO);
    const uint64_t w_base = make_desc(sbase, WLBO, WSBO);
    const uint32_t a_slot = pb.a_bytes >> 4;
    if (first_round) mbar_wait(pb.wfull, 0);
    if (stacked) {
        // single-request operand in the stacked layout: one N = 128 MMA per k16 (see store_stacked4)
        const uint64_t s_base = make_desc(sbase + (uint32_t)P_W_BYTES, 128, 1024);
        mbar_wait(pb.full0, rp.par & 1u);
        rp.par ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int kb = 0; kb < PNKB; ++kb) {
            const uint64_t ad = s_base + (uint64_t)(kb * (8192 >> 4));
            const uint64_t wd = w_base + (uint64_t)(kb * (PB_BYTES >> 4));
            if (kb == 0) umma_imm<false>(tmem_d, ad, wd, idesc128);
            else umma_imm<true>(tmem_d, ad, wd, idesc128);
#pragma unroll
            for (int kk = 1; kk < BK / 16; ++kk) umma_imm<true>(tmem_d, ad + kk * 16, wd + kk * 16, idesc128);
        }
        umma_commit(pb.accum);
        return;
    }
#pragma unroll 1
    if (pb.single) rot = 0;
    for (int kb = 0; kb < PNKB; ++kb) {
        if (!pb.single || kb == 0) {
            const int ws = pb.single ? 0 : rp.slot;
            mbar_wait(pb.full0 + 8 * ws, (rp.par >> ws) & 1u);
            rp.par ^= 1u << ws;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t ad = a_base + (uint64_t)(rp.slot * a_slot);
        const uint64_t wd = w_base + (uint64_t)(((kb + rot) & (PNKB - 1)) * (PB_BYTES >> 4));
        if (kb == 0) umma_imm<false>(tmem_d, ad, wd, idesc128);
        else umma_imm<true>(tmem_d, ad, wd, idesc128);
        umma_imm<true>(tmem_d, ad + 8, wd, idesc64);
#pragma unroll
        for (int kk = 1; kk < BK / 16; ++kk) {
            umma_imm<true>(tmem_d, ad + kk * 32, wd + kk * 16, idesc128);
            umma_imm<true>(tmem_d, ad + kk * 32 + 8, wd + kk * 16, idesc64);
        }
        // slots are recycled within a step only when the ring is shorter than the 8 k-blocks
        if (pb.nslots < PNKB) umma_commit(pb.empty0 + 8 * rp.slot);
        if (kb == PNKB - 1) umma_commit(pb.accum);
        ring_next(rp, pb.nslots);
    }
}

struct FwdArgs {
    const uint8_t* whpk; int mgp_w;     // Op_B[n = permuted gate column, k = hidden unit], gate tile 64
    uint8_t* hpk0; uint8_t* hpk1; int mgp_h;   // packed h ping-pong (hpk0 holds h0 on entry)
    float* gates; float* cells; float* Y;
    const float* h0; const float* c0; float* hT; float* cT;
    const int* len; int R, T; float forget_bias;
    unsigned* sync;                     // [row tiles] arrival counters, sync[63] = error word
    int rt;                             // rows per row tile (multiple of 8, <= 128): row tile mt = rows [mt*rt, mt*rt + rt)
    int full_fence;                     // reader-side proxy fence over all state spaces (see proxy_fence_reader)
    int stacked;                        // R == 32: operand in the stacked hi/lo layout (store_stacked4)
};

__global__ void __launch_bounds__(PTHREADS, 1) lstm_persist_fwd_kernel(const FwdArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * PMAXSLOTS + 2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.y, m0 = mt * a.rt, n0 = blockIdx.x * PBN, u0 = blockIdx.x * PUPT;
    const int H = PH, G4 = 4 * PH, R = a.R, rot = blockIdx.x & (PNKB - 1);
    unsigned* ctr = a.sync + mt;
    unsigned* err = a.sync + 63;
    int rows = R - m0;
    if (rows > a.rt) rows = a.rt;
    PersistBars pb;
    const size_t a_kb_stride = (size_t)a.mgp_h * 2048;
    const uint32_t tmem_d = persist_setup(smem, bars, &tmem_slot, pb, rows, a_kb_stride,
                                          a.whpk + (size_t)(n0 / PBN) * PB_BYTES, (size_t)a.mgp_w * 2048);
    const uint32_t sbase = smem_u32(smem);
    RingPos rp{0, 0, 0};   // used by the producer lane and (separately) by the MMA lane
    // Head of a step's operand stream: the first `nhead` k-blocks of step t+1 are issued at the END of step t,
    // right after the staged outputs have been read back into registers (the whole ring is free then) and
    // BEFORE the copy-out stores are issued.  The order matters: the producer lane's barrier polls
    // (ld.acquire) share the SM's memory pipe with those 57 KB of stores and used to sit behind them for
    // ~4 k cycles per step (tools/probe_persist.py).
    const int nhead = pb.single ? PNKB : (pb.nslots < PNKB ? pb.nslots : PNKB);
    int issued = 0;              // (producer lane) k-blocks of the coming step already in flight

    // This thread's item: accumulator row (= its TMEM lane) 32 (warp % 4) + lane, hidden units
    // u0 + 4 (warp / 4) .. + 4.  The cell state c and the carried h stay in registers.
    const int row = (warp & 3) * 32 + lane, jq = (warp >> 2) * 4, r = m0 + row;
    const bool valid = row < rows;      // (rows past the tile belong to the next row tile's CTAs)
    const size_t su = (size_t)(valid ? r : 0) * H + u0 + jq;
    const uint32_t tacc = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)jq;
    float c[4] = {0.f, 0.f, 0.f, 0.f}, h[4] = {0.f, 0.f, 0.f, 0.f};
    int mylen = 0;
    if (valid) {
        mylen = a.len[r];
        if (a.c0) ld4r(a.c0 + su, c);
        if (a.h0) ld4r(a.h0 + su, h);
    }

    // copy-out item of this thread (coalesced: 4 lanes per row, 64 B per row and array)
    const int crow = tid >> 2, cpart = (tid & 3) * 4, cr = m0 + crow;
    const bool cvalid = crow < rows;
    const int clen = cvalid ? a.len[cr] : 0;
    float* stg = reinterpret_cast<float*>(smem + P_W_BYTES);
    // copy-out registers (staged outputs of the step just finished)
    float4 o_c = make_float4(0.f, 0.f, 0.f, 0.f), o_y = o_c, o_g[4] = {o_c, o_c, o_c, o_c};
    auto copy_out = [&](int ts) {
        if (cvalid) {
            const size_t gu = ((size_t)ts * R + cr) * H + u0 + cpart;
            *reinterpret_cast<float4*>(a.cells + gu) = o_c;
            if (ts < clen) {
                float* gout = a.gates + ((size_t)ts * R + cr) * G4 + u0 + cpart;
#pragma unroll
                for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(gout + g * H) = o_g[g];
            }
            // ts >= len: output row is zero, (c, h) are carried through (dynamic_rnn, A.5)
            *reinterpret_cast<float4*>(a.Y + gu) = o_y;
        }
    };

    for (int t = 0; t < a.T; ++t) {
        const float* grow = a.gates + ((size_t)t * R + (valid ? r : 0)) * G4 + u0 + jq;
        const bool live = valid && t < mylen;
        float zi[4], zj[4], zf[4], zo[4];
        // (the two role lanes issue their own loads after their latency-critical work)
        if (tid == 0) { pstamp(t, 0); pstamp(t - 1, 11); }
        if (warp == 0 && lane == 0) {
            if (issued == 0) {
                if (t > 0) grid_wait(ctr, (unsigned)(PCOLS * t), err);
                pstamp(t, 1);
                proxy_fence_reader(a.full_fence);   // orders the previous step's generic staging accesses before the bulk writes
            }
            if (!(pb.single && issued > 0))
                persist_produce(pb, rp, sbase, ((t & 1) ? a.hpk1 : a.hpk0) + (size_t)(m0 / 8) * 2048, a_kb_stride, rot,
                                issued);
            pstamp(t, 2);
            if (live) { ld4r(grow, zi); ld4r(grow + H, zj); ld4r(grow + 2 * H, zf); ld4r(grow + 3 * H, zo); }
        } else if (warp == 1 && lane == 0) {
            persist_mma(pb, rp, sbase, tmem_d, rot, t == 0, a.stacked != 0);
            pstamp(t, 3);
            if (live) { ld4r(grow, zi); ld4r(grow + H, zj); ld4r(grow + 2 * H, zf); ld4r(grow + 3 * H, zo); }
            // only this lane polls the accumulator barrier; everybody else parks on the hardware
            // CTA barrier below (hundreds of threads spinning on try_wait steal shared-memory
            // bandwidth from the tensor core's operand reads)
            mbar_wait(pb.accum, t & 1);
        } else {
            // hoisted x*Wx + b part of this thread's pre-activations: in flight during the main loop
            if (live) { ld4r(grow, zi); ld4r(grow + H, zj); ld4r(grow + 2 * H, zf); ld4r(grow + 3 * H, zo); }
        }
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) pstamp(t, 5);
        float ai[4], aj[4], af[4], ao[4];
        float bi[4], bj[4], bf[4], bo[4];
        tmem_ld4(tacc, ai); tmem_ld4(tacc + PUPT, aj); tmem_ld4(tacc + 2 * PUPT, af); tmem_ld4(tacc + 3 * PUPT, ao);
        tmem_ld4(tacc + PBN, bi); tmem_ld4(tacc + PBN + PUPT, bj); tmem_ld4(tacc + PBN + 2 * PUPT, bf);
        tmem_ld4(tacc + PBN + 3 * PUPT, bo);
        tmem_ld_wait();
        if (a.stacked) {
            // accumulator rows 32-63 (TMEM lanes of the warps with warp % 4 == 1) hold h_lo * [Wh_hi | Wh_lo] of
            // rows 0-31: hand them to the owners of those rows through the ring (above the operand, idle now)
            float* xch = reinterpret_cast<float*>(smem + P_W_BYTES + 64 * 1024) + ((warp >> 2) * 32 + lane) * 16;
            if ((warp & 3) == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    xch[e] = ai[e] + bi[e]; xch[4 + e] = aj[e] + bj[e];
                    xch[8 + e] = af[e] + bf[e]; xch[12 + e] = ao[e] + bo[e];
                }
            }
            __syncthreads();
            if ((warp & 3) == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ai[e] += xch[e]; aj[e] += xch[4 + e]; af[e] += xch[8 + e]; ao[e] += xch[12 + e];
                }
            }
        }
        if (live) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                zi[e] = sigmoid_p(zi[e] + (ai[e] + bi[e]));
                zj[e] = tanh_p(zj[e] + (aj[e] + bj[e]));
                zf[e] = sigmoid_p(zf[e] + (af[e] + bf[e]) + a.forget_bias);
                zo[e] = sigmoid_p(zo[e] + (ao[e] + bo[e]));
                c[e] = c[e] * zf[e] + zi[e] * zj[e];
                h[e] = tanh_p(c[e]) * zo[e];
            }
        }
        if (tid == 64) pstamp(t, 6);
        // publish h_t (or the carried h) as next step's packed operand FIRST; everything the
        // backward pass needs is written out after the arrive, off the step-to-step critical path
        if (t + 1 < a.T && valid) {
            if (a.stacked) store_stacked4((t & 1) ? a.hpk0 : a.hpk1, r, u0 + jq, h);
            else store_packed4((t & 1) ? a.hpk0 : a.hpk1, a.mgp_h, r, u0 + jq, h);
        }
        if (tid == 64) pstamp(t, 7);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // all MMAs of this step retired (ring idle), accumulator read, h_t stores issued
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // one gpu-scope fence by the arriving thread publishes the whole CTA's stores (they
        // happen-before it through the CTA barrier; fences are cumulative)
        if (tid == 0 && t + 1 < a.T) { __threadfence(); proxy_fence(); red_relaxed(ctr, 1u); pstamp(t, 8); }
        // The thread that owns an accumulator row holds 16 B of each output row; writing those
        // straight to HBM costs 32 cache lines per warp store.  Park the outputs in the idle
        // ring and write them back with 4 lanes per row (64 B segments).
        if (valid) {
            if (live) {
                st4r(stage_ptr(stg, 0, row, jq), zi); st4r(stage_ptr(stg, 1, row, jq), zj);
                st4r(stage_ptr(stg, 2, row, jq), zf); st4r(stage_ptr(stg, 3, row, jq), zo);
            }
            st4r(stage_ptr(stg, 4, row, jq), c);
            st4r(stage_ptr(stg, 5, row, jq), h);
        }
        __syncthreads();
        if (tid == 0) pstamp(t, 14);
        // staged outputs -> registers (4 lanes per row, 64 B per row and array), so that the ring is free again
        o_y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cvalid) {
            o_c = *reinterpret_cast<const float4*>(stage_ptr(stg, 4, crow, cpart));
            if (t < clen) {
#pragma unroll
                for (int g = 0; g < 4; ++g) o_g[g] = *reinterpret_cast<const float4*>(stage_ptr(stg, g, crow, cpart));
                o_y = *reinterpret_cast<const float4*>(stage_ptr(stg, 5, crow, cpart));
            }
        }
        __syncthreads();
        if (tid == 0) {
            issued = 0;
            pstamp(t, 15);
            if (t + 1 < a.T) {
                // head of step t+1's operand stream (all MMAs of step t have retired: every slot is free)
                grid_wait(ctr, (unsigned)(PCOLS * (t + 1)), err);
                pstamp(t, 17);
                proxy_fence_reader(a.full_fence);
                persist_produce(pb, rp, sbase, (((t + 1) & 1) ? a.hpk1 : a.hpk0) + (size_t)(m0 / 8) * 2048,
                                a_kb_stride, rot, 0, nhead);
                issued = nhead;
                pstamp(t, 10);
            }
        }
        // (measured, profiles/r02u_*: holding the other warps' stores back behind one more CTA barrier, or
        // deferring the two role warps' stores to after their next role work, is no faster than this)
        copy_out(t);
        if (tid == 64) pstamp(t, 9);
        if (tid == PTHREADS - 1) pstamp(t, 13);
        if (tid == 0) pstamp(t, 12);
    }
    if (valid) {
        st4r(a.hT + su, h);
        st4r(a.cT + su, c);
    }
    __syncthreads();
    pstamp_flush();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)PTMEM));
}

// ---------------------------------------------------------------------------------------------
// Compact forward recurrence: ONE column of 32 CTAs walks all row tiles of the sequence batch
// (R <= 512 rows in up to 4 tiles) instead of one CTA per (column, row tile).  Row tiles are
// independent recurrences, so a CTA time-multiplexes them: while tile j's cell math / publish /
// copy-out runs on the epilogue warps, the producer lane is already streaming tile j+1's operand
// and the MMA lane is issuing its products into tile j+1's own TMEM accumulator (NT x 128
// columns).  The step-to-step dependency chain of one tile (barrier, bulk-copy latency, MMA,
// MUFU, publish) hides behind the other tiles' streams, so 32 SMs sustain about the step rate
// 96 SMs reach with one tile per CTA - and three such recurrences (the action, perception and
// program decoders of models/model_full.py:440-595) fit on the GPU side by side.
//   warps 0..15  epilogue: thread = (accumulator row, 4 hidden units); c and h of every tile in registers
//   warp 16      producer lane: per (step, tile) waits for the tile's 32 publishers, streams the
//                8 k-blocks of packed h_{t-1} through the ring
//   warp 17      MMA lane: Ahi*[Bhi|Blo] (N=128) + Alo*Bhi (N=64) per k16 into the tile's accumulator
constexpr int MT_MAX_TILES = 4;
constexpr int MT_EPI_THREADS = 512;
constexpr int MT_THREADS = MT_EPI_THREADS + 64;
constexpr long long MT_MBAR_SPIN = 2000000000LL;

struct FwdMtArgs {
    FwdArgs f;
    int nt;                        // row tiles
    int g0[MT_MAX_TILES + 1];      // first 8-row group of tile j (g0[nt] = groups)
};

// mbarrier wait with a time-out: a protocol error sets the sticky word instead of hanging the GPU
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity, unsigned* err) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    int it = 0;
    while (!mbar_try(bar, parity)) {
        if ((++it & 255) == 0) {
            if (ld_acquire(err) != 0u) return;
            if (clock64() - t0 > MT_MBAR_SPIN) { atomicExch(err, 1u); atomicExch(&g_persist_sticky_error, 1u); return; }
        }
    }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// timeline probe of the compact kernels: CTA 0, step P_PROBE_STEP, tile j -> slots 64 + 16 j + k
__device__ __forceinline__ void mstamp(int step, int j, int k) {
    if (g_tc_dbg != nullptr && step == P_PROBE_STEP && blockIdx.x == 0) g_tc_dbg[64 + 16 * j + k] = clock64();
}

template <int NT>
__global__ void __launch_bounds__(MT_THREADS, 1) lstm_persist_fwd_mt_kernel(const FwdMtArgs am) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * PMAXSLOTS + 2 * MT_MAX_TILES + 1];
    __shared__ uint32_t tmem_slot;
    const FwdArgs& a = am.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * PBN, u0 = blockIdx.x * PUPT;
    const int H = PH, G4 = 4 * PH, R = a.R, rot = blockIdx.x & (PNKB - 1);
    unsigned* err = a.sync + 63;
    constexpr uint32_t TCOLS = NT <= 2 ? 256u : 512u;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[PMAXSLOTS]);
    const uint32_t accf0 = smem_u32(&bars[2 * PMAXSLOTS]), acce0 = smem_u32(&bars[2 * PMAXSLOTS + MT_MAX_TILES]);
    const uint32_t wfull = smem_u32(&bars[2 * PMAXSLOTS + 2 * MT_MAX_TILES]);
    // ring geometry: slots sized for the largest tile; the MMA reads a full 128-row image from a
    // slot, so the last slot must start at least 32 KB before the end of the ring
    int gmax = 0;
#pragma unroll
    for (int j = 0; j < NT; ++j) gmax = max(gmax, am.g0[j + 1] - am.g0[j]);
    const uint32_t slot_bytes = (uint32_t)gmax * 2048u;
    int nslots = (int)((P_RING_BYTES - (BM / 8) * 2048) / slot_bytes) + 1;
    if (nslots > PMAXSLOTS) nslots = PMAXSLOTS;
    const size_t a_kb_stride = (size_t)a.mgp_h * 2048;

    if (tid == 0) {
        for (int s2 = 0; s2 < 2 * PMAXSLOTS; ++s2) mbar_init(full0 + 8 * s2, 1);
        for (int j = 0; j < MT_MAX_TILES; ++j) { mbar_init(accf0 + 8 * j, 1); mbar_init(acce0 + 8 * j, MT_EPI_THREADS / 32); }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // rows a partial tile never loads must hold finite values
        uint4* ring = reinterpret_cast<uint4*>(smem + P_W_BYTES);
        for (int i = tid; i < (int)(P_RING_BYTES / 16); i += MT_THREADS) ring[i] = make_uint4(0, 0, 0, 0);
    }
    proxy_fence();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;

    if (warp == MT_EPI_THREADS / 32) {
        // ===================== producer lane =====================
        if (lane == 0) {
            mbar_expect_tx(wfull, (uint32_t)P_W_BYTES);
            const uint8_t* wsrc = a.whpk + (size_t)(n0 / PBN) * PB_BYTES;
            for (int kb = 0; kb < PNKB; ++kb)
                bulk_copy(sbase + kb * PB_BYTES, wsrc + (size_t)kb * a.mgp_w * 2048, PB_BYTES, wfull);
            int slot = 0;
            uint32_t phase = 0;
            bool wrapped = false;
            for (int t = 0; t < a.T; ++t) {
                const uint8_t* hsrc = (t & 1) ? a.hpk1 : a.hpk0;
#pragma unroll 1
                for (int j = 0; j < NT; ++j) {
                    mstamp(t, j, 0);
                    if (t > 0) grid_wait(a.sync + j, (unsigned)(PCOLS * t), err);
                    mstamp(t, j, 1);
                    proxy_fence_reader(a.full_fence);
                    const uint32_t bytes = (uint32_t)(am.g0[j + 1] - am.g0[j]) * 2048u;
                    const uint8_t* src = hsrc + (size_t)am.g0[j] * 2048;
                    for (int kb = 0; kb < PNKB; ++kb) {
                        if (wrapped) mbar_wait_bounded(empty0 + 8 * slot, phase ^ 1u, err);
                        const uint32_t bar = full0 + 8 * slot;
                        mbar_expect_tx(bar, bytes);
                        bulk_copy(sbase + (uint32_t)P_W_BYTES + slot * slot_bytes,
                                  src + (size_t)((kb + rot) & (PNKB - 1)) * a_kb_stride, bytes, bar);
                        if (++slot == nslots) { slot = 0; phase ^= 1u; wrapped = true; }
                    }
                    mstamp(t, j, 2);
                }
            }
        }
    } else if (warp == MT_EPI_THREADS / 32 + 1) {
        // ===================== MMA lane =====================
        if (lane == 0) {
            constexpr uint32_t LBO = 256, SBO = 2048, WLBO = 128, WSBO = 1024;
            constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PBN >> 3) << 17) |
                                         ((uint32_t)(BM >> 4) << 24);
            constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * PBN) >> 3) << 17) |
                                          ((uint32_t)(BM >> 4) << 24);
            const uint64_t a_base = make_desc(sbase + (uint32_t)P_W_BYTES, LBO, SBO);
            const uint64_t w_base = make_desc(sbase, WLBO, WSBO);
            const uint32_t a_slot = slot_bytes >> 4;
            int slot = 0;
            uint32_t phase = 0;
            mbar_wait_bounded(wfull, 0, err);
            for (int t = 0; t < a.T; ++t) {
#pragma unroll 1
                for (int j = 0; j < NT; ++j) {
                    mstamp(t, j, 3);
                    if (t > 0) mbar_wait_bounded(acce0 + 8 * j, (uint32_t)((t - 1) & 1), err);
                    mstamp(t, j, 4);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t td = tmem_d + (uint32_t)(j * PTMEM);
#pragma unroll 1
                    for (int kb = 0; kb < PNKB; ++kb) {
                        mbar_wait_bounded(full0 + 8 * slot, phase, err);
                        if (kb == 0) mstamp(t, j, 5);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t ad = a_base + (uint64_t)(slot * a_slot);
                        const uint64_t wd = w_base + (uint64_t)(((kb + rot) & (PNKB - 1)) * (PB_BYTES >> 4));
                        if (kb == 0) umma_imm<false>(td, ad, wd, idesc128);
                        else umma_imm<true>(td, ad, wd, idesc128);
                        umma_imm<true>(td, ad + 8, wd, idesc64);
#pragma unroll
                        for (int kk = 1; kk < BK / 16; ++kk) {
                            umma_imm<true>(td, ad + kk * 32, wd + kk * 16, idesc128);
                            umma_imm<true>(td, ad + kk * 32 + 8, wd + kk * 16, idesc64);
                        }
                        umma_commit(empty0 + 8 * slot);
                        if (kb == PNKB - 1) umma_commit(accf0 + 8 * j);
                        if (++slot == nslots) { slot = 0; phase ^= 1u; }
                    }
                    mstamp(t, j, 6);
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int row = (warp & 3) * 32 + lane, jq = (warp >> 2) * 4;
        const uint32_t tacc = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)jq;
        float c[NT][4], h[NT][4];
        int mylen[NT], rr[NT];
        bool valid[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int nr = min((am.g0[j + 1] - am.g0[j]) * 8, R - am.g0[j] * 8);
            rr[j] = am.g0[j] * 8 + row;
            valid[j] = row < nr;
            mylen[j] = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) c[j][e] = h[j][e] = 0.f;
            if (valid[j]) {
                const size_t su = (size_t)rr[j] * H + u0 + jq;
                mylen[j] = a.len[rr[j]];
                if (a.c0) ld4r(a.c0 + su, c[j]);
                if (a.h0) ld4r(a.h0 + su, h[j]);
            }
        }
        for (int t = 0; t < a.T; ++t) {
            uint8_t* hnext = (t & 1) ? a.hpk0 : a.hpk1;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int r = valid[j] ? rr[j] : 0;
                const bool live = valid[j] && t < mylen[j];
                float* grow = a.gates + ((size_t)t * R + r) * G4 + u0 + jq;
                float zi[4], zj[4], zf[4], zo[4];
                if (live) { ld4r(grow, zi); ld4r(grow + H, zj); ld4r(grow + 2 * H, zf); ld4r(grow + 3 * H, zo); }
                if (tid == 0) mstamp(t, j, 7);
                if (lane == 0) mbar_wait_bounded(accf0 + 8 * j, (uint32_t)(t & 1), err);
                __syncwarp();
                if (tid == 0) mstamp(t, j, 8);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    const uint32_t tj = tacc + (uint32_t)(j * PTMEM);
                    float ai[4], aj[4], af[4], ao[4];
                    float bi[4], bj[4], bf[4], bo[4];
                    tmem_ld4(tj, ai); tmem_ld4(tj + PUPT, aj); tmem_ld4(tj + 2 * PUPT, af); tmem_ld4(tj + 3 * PUPT, ao);
                    tmem_ld4(tj + PBN, bi); tmem_ld4(tj + PBN + PUPT, bj); tmem_ld4(tj + PBN + 2 * PUPT, bf);
                    tmem_ld4(tj + PBN + 3 * PUPT, bo);
                    tmem_ld_wait();
                    // the accumulator of tile j may be overwritten by the next step's products
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acce0 + 8 * j);
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            zi[e] = sigmoid_p(zi[e] + (ai[e] + bi[e]));
                            zj[e] = tanh_p(zj[e] + (aj[e] + bj[e]));
                            zf[e] = sigmoid_p(zf[e] + (af[e] + bf[e]) + a.forget_bias);
                            zo[e] = sigmoid_p(zo[e] + (ao[e] + bo[e]));
                            c[j][e] = c[j][e] * zf[e] + zi[e] * zj[e];
                            h[j][e] = tanh_p(c[j][e]) * zo[e];
                        }
                    }
                }
                // Everything this step leaves in global memory is stored BEFORE the publish: a gpu-scope
                // fence waits for every store the SM has in flight, so output stores issued by the other
                // warps while the publishing thread sits in its fence would stretch the fence (measured:
                // 10 k cycles).  The tile's chain has slack here - the other tiles' streams hide it.
                if (tid == 0) mstamp(t, j, 9);
                if (valid[j]) {
                    if (t + 1 < a.T) store_packed4(hnext, a.mgp_h, rr[j], u0 + jq, h[j]);
                    const size_t gu = ((size_t)t * R + rr[j]) * H + u0 + jq;
                    st4r(a.cells + gu, c[j]);
                    if (live) {
                        st4r(grow, zi); st4r(grow + H, zj); st4r(grow + 2 * H, zf); st4r(grow + 3 * H, zo);
                        st4r(a.Y + gu, h[j]);
                    } else {
                        *reinterpret_cast<float4*>(a.Y + gu) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                epi_bar_sync();
                if (tid == 0) mstamp(t, j, 10);
                if (tid == 0 && t + 1 < a.T) { __threadfence(); proxy_fence(); red_relaxed(a.sync + j, 1u); }
                if (tid == 0) mstamp(t, j, 11);
                if (tid == 0) mstamp(t, j, 12);
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j)
            if (valid[j]) {
                const size_t su = (size_t)rr[j] * H + u0 + jq;
                st4r(a.hT + su, h[j]);
                st4r(a.cT + su, c[j]);
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TCOLS));
}

struct BwdArgs {
    const uint8_t* wtpk; int mgp_w;     // Op_B[n = hidden unit, k = gate column] = Wh[n, k]
    uint8_t* dzpk; int mgp_z;           // packed dZ: one step [R, 4H], or (full != 0) all steps [T*R, 4H]
    int full;                           // dZ_t lives at rows t*R.. of the full operand (reused by the dX product)
    float* gates;                       // [T,R,4H] in: activated gates, out: dZ
    const float* cells; const float* c0; const float* dY;
    const float* dhT; const float* dcT; float* dh0; float* dc0;
    const int* len; int R, T, has_h0;
    int rt;                             // rows per row tile (see FwdArgs)
    int full_fence;
    int stacked;                        // R == 32: packed dZ in the stacked hi/lo layout
    float* dbpart;                      // [row tiles][4H] column sums of dZ over this tile's rows and all steps
    unsigned* sync;                     // [row tiles] dZ_t published counters, [63] error word
};

// cluster helpers (the 4 K-chunk CTAs of one 64-unit output tile form a cluster)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem4(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(cta_rank));
    float4 v;
    // no "memory" clobber: the four loads of a reduction stay in flight together; they are
    // ordered against the cluster barrier (which has one)
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(raddr));
    return v;
}
constexpr int PGROW = PBN + 4;          // staged partial-sum tile row (floats): conflict-free float4 columns

// Backward recurrence.  CTA q of a row tile: output tile nt = q / 4 (64 hidden units), K-chunk
// ks = q % 4 (= gate ks, 512 gate columns); the 4 K-chunk CTAs of a tile are one CLUSTER.
// Step t:  phase P - this CTA's 16 hidden units (64 nt + 16 ks ..): dh = carry + the four
//          partial tiles of the cluster (own + 3 over distributed shared memory, fixed order),
//          cell backward, dZ_t published as packed operand;
//          row-tile barrier;
//          phase G - partial[ks] = dZ_t[:, gate ks] * Wh[tile nt, gate ks]^T into the
//          accumulator, parked in the (idle) ring for the cluster;
//          cluster barrier.
// The partial sums never leave the SMs.  A peer's reads of this CTA's parked tile are over before
// that peer publishes its dZ, i.e. before this CTA's next bulk copies (which wait for all 32
// publishers of the row tile) can overwrite the ring.
__global__ void __launch_bounds__(PTHREADS, 1) lstm_persist_bwd_kernel(const BwdArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * PMAXSLOTS + 2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.y, m0 = mt * a.rt;
    const int q = blockIdx.x, nt = q >> 2, ks = q & 3;     // 64-unit output tile, K-chunk (= gate)
    const int kb0 = ks * PNKB, rot = nt & (PNKB - 1);
    const int H = PH, G4 = 4 * PH, R = a.R;
    const size_t RH = (size_t)R * H;
    unsigned* ctrP = a.sync + mt;
    unsigned* err = a.sync + 63;
    int rows = R - m0;
    if (rows > a.rt) rows = a.rt;
    PersistBars pb;
    const size_t a_kb_stride = (size_t)a.mgp_z * 2048;
    const uint32_t tmem_d = persist_setup(smem, bars, &tmem_slot, pb, rows, a_kb_stride,
                                          a.wtpk + (size_t)kb0 * a.mgp_w * 2048 + (size_t)nt * PB_BYTES,
                                          (size_t)a.mgp_w * 2048);
    const uint32_t sbase = smem_u32(smem);
    float* stage = reinterpret_cast<float*>(smem + P_W_BYTES);          // [128][PGROW] partial tile
    const uint32_t stage_u32 = sbase + (uint32_t)P_W_BYTES;
    RingPos rp{0, 0, 0};

    // phase-P item of this thread: (row, 4 hidden units) of units [16 q, 16 q + 16)
    const int row = tid >> 2, up = (tid & 3) * 4, u = q * PUPT + up, r = m0 + row;
    const bool valid = row < rows;
    const size_t su = (size_t)(valid ? r : 0) * H + u;
    // my 16 B of each partial tile: row `row`, tile columns 16 ks + up .. + 4
    const uint32_t my_part = stage_u32 + (uint32_t)((row * PGROW + ks * PUPT + up) * sizeof(float));
    float dhc[4] = {0.f, 0.f, 0.f, 0.f}, dcc[4] = {0.f, 0.f, 0.f, 0.f};
    float sdb[16];           // bias gradient: this thread's dZ summed over the steps (i | j | f | o)
#pragma unroll
    for (int e = 0; e < 16; ++e) sdb[e] = 0.f;
    int mylen = 0;
    if (valid) {
        mylen = a.len[r];
        if (a.dhT) ld4r(a.dhT + su, dhc);
        if (a.dcT) ld4r(a.dcT + su, dcc);
    }
    // Saved activations / incoming gradient of a step, fetched one step ahead (during the previous step's
    // phase G, when the epilogue threads are idle): the loads used to sit at the top of the step, ~2 k
    // cycles of exposed latency in front of phase P.
    float gi[4], gj[4], gf[4], go[4], cc[4], cp[4], dy[4];
    float ngi[4], ngj[4], ngf[4], ngo[4], ncc[4], ncp[4], ndy[4];
    auto fetch = [&](int ts, float* xi, float* xj, float* xf, float* xo, float* xcc, float* xcp, float* xdy) {
        if (valid && ts >= 0 && ts < mylen) {
            const float* gs = a.gates + ((size_t)ts * R + r) * G4 + u;
            ld4r(gs, xi); ld4r(gs + H, xj); ld4r(gs + 2 * H, xf); ld4r(gs + 3 * H, xo);
            ld4r(a.cells + (size_t)ts * RH + su, xcc);
            if (ts > 0) ld4r(a.cells + (size_t)(ts - 1) * RH + su, xcp);
            else if (a.c0) ld4r(a.c0 + su, xcp);
            else { xcp[0] = xcp[1] = xcp[2] = xcp[3] = 0.f; }
            if (a.dY) ld4r(a.dY + (size_t)ts * RH + su, xdy);
            else { xdy[0] = xdy[1] = xdy[2] = xdy[3] = 0.f; }
        }
    };
    fetch(a.T - 1, gi, gj, gf, go, cc, cp, dy);
    cluster_sync_all();      // every CTA of the cluster is resident and set up before any remote access
    int ground = 0;          // GEMM rounds completed so far
    bool have_partials = false;
    for (int t = a.T - 1; t >= 0; --t) {
        const int step = a.T - 1 - t;
        const int zrow0 = a.full ? t * R : 0;      // first packed row of this step's dZ
        if (tid == 0) { pstamp(step, 16); pstamp(step - 1, 16 + 9); }
        // ---- phase P: dZ_t for this CTA's 16 hidden units ----
        float* g = a.gates + ((size_t)t * R + (valid ? r : 0)) * G4 + u;
        const bool live = valid && t < mylen;
        if (have_partials && valid) {
            const float4 p0 = ld_dsmem4(my_part, 0u), p1 = ld_dsmem4(my_part, 1u);
            const float4 p2 = ld_dsmem4(my_part, 2u), p3 = ld_dsmem4(my_part, 3u);
            // fixed order: K-chunks 0..3
            dhc[0] += p0.x; dhc[1] += p0.y; dhc[2] += p0.z; dhc[3] += p0.w;
            dhc[0] += p1.x; dhc[1] += p1.y; dhc[2] += p1.z; dhc[3] += p1.w;
            dhc[0] += p2.x; dhc[1] += p2.y; dhc[2] += p2.z; dhc[3] += p2.w;
            dhc[0] += p3.x; dhc[1] += p3.y; dhc[2] += p3.z; dhc[3] += p3.w;
            if (a.stacked) {    // rows 32-63 of the parked tiles: the dZ_lo part of the same rows
#pragma unroll
                for (int sc = 0; sc < 4; ++sc) {
                    const float4 pl = ld_dsmem4(my_part + (uint32_t)(P_STACK_ROWS * PGROW * sizeof(float)), (uint32_t)sc);
                    dhc[0] += pl.x; dhc[1] += pl.y; dhc[2] += pl.z; dhc[3] += pl.w;
                }
            }
        }
        if (tid == 64) pstamp(step, 18);
        float di[4], dj[4], df[4], dq[4];
        if (live) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float dht = dhc[e] + dy[e];
                const float tcn = tanh_p(cc[e]);
                dq[e] = dht * tcn * go[e] * (1.f - go[e]);
                const float dct = dcc[e] + dht * go[e] * (1.f - tcn * tcn);
                di[e] = dct * gj[e] * gi[e] * (1.f - gi[e]);
                dj[e] = dct * gi[e] * (1.f - gj[e] * gj[e]);
                df[e] = dct * cp[e] * gf[e] * (1.f - gf[e]);
                dcc[e] = dct * gf[e];
                dhc[e] = 0.f;   // consumed: the recurrent product carries the new dh
                sdb[e] += di[e]; sdb[4 + e] += dj[e]; sdb[8 + e] += df[e]; sdb[12 + e] += dq[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) di[e] = dj[e] = df[e] = dq[e] = 0.f;   // state copied through
        }
        const bool do_gemm = t > 0 || a.has_h0;
        if (valid && (do_gemm || a.full)) {
            if (a.stacked) {
                store_stacked4(a.dzpk, r, u, di); store_stacked4(a.dzpk, r, H + u, dj);
                store_stacked4(a.dzpk, r, 2 * H + u, df); store_stacked4(a.dzpk, r, 3 * H + u, dq);
            } else {
                store_packed4(a.dzpk, a.mgp_z, zrow0 + r, u, di);
                store_packed4(a.dzpk, a.mgp_z, zrow0 + r, H + u, dj);
                store_packed4(a.dzpk, a.mgp_z, zrow0 + r, 2 * H + u, df);
                store_packed4(a.dzpk, a.mgp_z, zrow0 + r, 3 * H + u, dq);
            }
        }
        if (do_gemm) {
            // packed operand first, then publish; the fp32 dZ (for the dW products after the
            // kernel) is stored behind the arrive
            if (tid == 64) pstamp(step, 19);
            __syncthreads();
            if (tid == 0) { __threadfence(); proxy_fence(); red_relaxed(ctrP, 1u); pstamp(step, 20); }
            __syncthreads();
        }
        if (!do_gemm) {
            if (valid) { st4r(g, di); st4r(g + H, dj); st4r(g + 2 * H, df); st4r(g + 3 * H, dq); }
            have_partials = false;
            break;
        }
        // ---- phase G: partial[ks] = dZ_t[:, gate ks] * Wh[n-tile, gate ks]^T ----
        // The fp32 dZ stores (32 KB per CTA, for the dW products after the kernel) are issued only once the
        // head of the operand stream is out: the producer lane's barrier polls share the SM's memory pipe
        // with them and sat behind them for ~1.5 k cycles per step (tools/probe_persist.py).
        const uint8_t* zsrc = a.dzpk + ((size_t)kb0 * a.mgp_z + (zrow0 + m0) / 8) * 2048;
        const int nhead = pb.single ? PNKB : (pb.nslots < PNKB ? pb.nslots : PNKB);
        if (warp == 0 && lane == 0) {
            grid_wait(ctrP, (unsigned)(PCOLS * (ground + 1)), err);
            pstamp(step, 21);
            proxy_fence_reader(a.full_fence);
            persist_produce(pb, rp, sbase, zsrc, a_kb_stride, rot, 0, nhead);
        }
        __syncthreads();
        if (warp == 0 && lane == 0) {
            if (!pb.single) persist_produce(pb, rp, sbase, zsrc, a_kb_stride, rot, nhead, PNKB);
        } else if (warp == 1 && lane == 0) {
            persist_mma(pb, rp, sbase, tmem_d, rot, ground == 0, a.stacked != 0);
        }
        fetch(t - 1, ngi, ngj, ngf, ngo, ncc, ncp, ndy);     // next step's inputs
        if (valid) { st4r(g, di); st4r(g + H, dj); st4r(g + 2 * H, df); st4r(g + 3 * H, dq); }
        if (warp == 1 && lane == 0) mbar_wait(pb.accum, ground & 1);
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) pstamp(step, 22);
        {   // accumulator ([A*Bhi + Alo*Bhi | Ahi*Blo]) -> the idle ring, one row per thread
            const int lq = warp & 3, cg = warp >> 2;
            uint32_t v[PUPT], w[PUPT];
            tmem_ld16(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(cg * PUPT), v);
            tmem_ld16(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(PBN + cg * PUPT), w);
            tmem_ld_wait();
            float* dst = stage + (size_t)(lq * 32 + lane) * PGROW + cg * PUPT;
#pragma unroll
            for (int j = 0; j < PUPT; j += 4)
                *reinterpret_cast<float4*>(dst + j) =
                    make_float4(__uint_as_float(v[j]) + __uint_as_float(w[j]),
                                __uint_as_float(v[j + 1]) + __uint_as_float(w[j + 1]),
                                __uint_as_float(v[j + 2]) + __uint_as_float(w[j + 2]),
                                __uint_as_float(v[j + 3]) + __uint_as_float(w[j + 3]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (tid == 64) pstamp(step, 23);
        cluster_sync_all();      // the four partial tiles of the cluster are parked and visible
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) pstamp(step, 24);
        ++ground;
        have_partials = true;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            cc[e] = ncc[e];
            gi[e] = ngi[e]; gj[e] = ngj[e]; gf[e] = ngf[e]; go[e] = ngo[e]; cp[e] = ncp[e]; dy[e] = ndy[e];
        }
    }
    if (have_partials && valid) {   // dh0 = carry + last partial sums
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
            const float4 p = ld_dsmem4(my_part, (uint32_t)sc);
            dhc[0] += p.x; dhc[1] += p.y; dhc[2] += p.z; dhc[3] += p.w;
        }
        if (a.stacked) {
#pragma unroll
            for (int sc = 0; sc < 4; ++sc) {
                const float4 p = ld_dsmem4(my_part + (uint32_t)(P_STACK_ROWS * PGROW * sizeof(float)), (uint32_t)sc);
                dhc[0] += p.x; dhc[1] += p.y; dhc[2] += p.z; dhc[3] += p.w;
            }
        }
    }
    if (valid) {
        st4r(a.dh0 + su, dhc);
        st4r(a.dc0 + su, dcc);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();      // no CTA leaves (or recycles its ring) while a peer may still read its shared memory
    // bias gradient: column sums of dZ over this tile's rows, fixed order (rows 0..127)
    {
        float* srow = stage + (size_t)row * PGROW + up;
#pragma unroll
        for (int gte = 0; gte < 4; ++gte)
            *reinterpret_cast<float4*>(srow + gte * PUPT) =
                valid ? make_float4(sdb[gte * 4], sdb[gte * 4 + 1], sdb[gte * 4 + 2], sdb[gte * 4 + 3])
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        if (tid < PBN) {
            float s = 0.f;
            for (int rr = 0; rr < BM; ++rr) s += stage[(size_t)rr * PGROW + tid];
            a.dbpart[(size_t)mt * G4 + (tid >> 4) * H + q * PUPT + (tid & 15)] = s;
        }
    }
    __syncthreads();
    pstamp_flush();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)PTMEM));
}

int g_persist_mode = 1;   // 0 = per-step launches, 1 = persistent kernels where supported
int g_persist_full_fence = 0;   // reader-side proxy fence over all state spaces (A/B switch)
int g_persist_stacked = 1;      // 32-row recurrences use the stacked hi/lo operand (d2p_lstm_set_persistent bit 3 clears it)

template <class Args>
int launch_coop(void (*kern)(const Args), dim3 grid, size_t smem, cudaStream_t st, const Args& args,
                int cluster_x = 1, int threads = PTHREADS) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (g_persist_mode != 2) {   // mode 2 (experiment): plain launch
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    D2P_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
    count_launch();
    return 0;
}

}  // namespace

int lstm_persist_error(unsigned* out, bool clear) {
    D2P_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_persist_sticky_error, sizeof(unsigned)));
    if (clear && *out) {
        const unsigned zero = 0;
        D2P_CHECK_CUDA(cudaMemcpyToSymbol(g_persist_sticky_error, &zero, sizeof(unsigned)));
    }
    return 0;
}

// device address of the sticky error word (read by the optimizer kernel and the async error copy)
int lstm_persist_error_addr(unsigned** out) {
    D2P_CHECK_CUDA(cudaGetSymbolAddress((void**)out, g_persist_sticky_error));
    return 0;
}

int lstm_persist_set_probe(long long* buf) {
    D2P_CHECK_CUDA(cudaMemcpyToSymbol(tc::g_tc_dbg, &buf, sizeof(buf)));
    return 0;
}

// rows per row tile: 128, or - for a recurrence that has the GPU to itself (D2P_LSTM_WIDE) - the rows split
// evenly over as many row tiles as co-resident CTAs allow (4 x 32): R = 320 -> 4 tiles of 80 rows, so the
// per-step operand stream of a CTA is 160 KB instead of 256 KB
static int persist_row_tile(int R, bool wide) {
    if (!wide || R <= BM) return BM;
    const int nt = kNumSMs / PCOLS, groups = cdiv(R, 8);
    const int rt = 8 * cdiv(groups, nt);
    return rt < BM ? rt : BM;
}

bool lstm_persist_supported(int R, int H) {
    return g_persist_mode != 0 && tc_available() && H == PH && R >= 1 && cdiv(R, BM) * PCOLS <= kNumSMs;
}

// Compact grid (32 CTAs walking all row tiles): more than one row tile and at most MT_MAX_TILES.
bool lstm_persist_compact_supported(int R, int H) {
    return lstm_persist_supported(R, H) && R > BM && cdiv(cdiv(R, 8), BM / 8) <= MT_MAX_TILES;
}

namespace {
// balanced split of the 8-row groups over the fewest 128-row tiles
int mt_partition(int R, int* g0) {
    const int G = cdiv(R, 8), nt = cdiv(G, BM / 8);
    const int base = G / nt, rem = G % nt;
    g0[0] = 0;
    for (int j = 0; j < nt; ++j) g0[j + 1] = g0[j] + base + (j < rem ? 1 : 0);
    return nt;
}
template <int NT>
int launch_fwd_mt(cudaStream_t st, const FwdMtArgs& am) {
    static bool attr_set = false;
    if (!attr_set) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(lstm_persist_fwd_mt_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)P_SMEM));
        attr_set = true;
    }
    return launch_coop<FwdMtArgs>(lstm_persist_fwd_mt_kernel<NT>, dim3(PCOLS, 1), P_SMEM, st, am, 1, MT_THREADS);
}
}  // namespace

// Recurrence phase of lstm_seq_fwd: gates already hold X*Wx + b.
int lstm_persist_fwd(cudaStream_t st, int T, int R, int H, const int* len, const float* h0, const float* c0,
                     const float* Wh, float forget_bias, float* Y, float* hT, float* cT, float* gates,
                     float* cells, bool compact, bool wide) {
    const int G4 = 4 * H;
    size_t off = 0;
    // R <= 32: no padding to 128-row tiles, so a row tile's 8 k-blocks are contiguous (one bulk copy)
    const int mgp_h = R <= 32 ? cdiv(R, 8) : mgp_of(R);
    const size_t hbytes = (size_t)kgp_of(H) * mgp_h * 256;
    FwdArgs a;
    a.hpk0 = (uint8_t*)tc_scratch_alloc(st, &off, hbytes);
    a.hpk1 = (uint8_t*)tc_scratch_alloc(st, &off, hbytes);
    a.sync = (unsigned*)tc_scratch_alloc(st, &off, 256);
    D2P_REQUIRE(a.hpk0 && a.hpk1 && a.sync, "lstm persist fwd: tensor-core scratch arena too small");
    const void* whpk;
    D2P_TRY(get_packed(st, Wh, G4, H, G4, false, true, &off, &whpk, 1000 + PBN, H));
    const bool stacked = R == P_STACK_ROWS && g_persist_stacked;
    if (h0 && stacked) {
        pack_stacked32_kernel<<<cdiv(P_STACK_ROWS * (H / 4), 256), 256, 0, st>>>(h0, H, a.hpk0);
        D2P_CHECK_LAUNCH();
    } else if (h0) D2P_TRY(pack_bf16(st, h0, R, H, H, true, a.hpk0, 0, 0, mgp_h));
    else D2P_CHECK_CUDA(cudaMemsetAsync(a.hpk0, 0, hbytes, st));
    D2P_CHECK_CUDA(cudaMemsetAsync(a.hpk1, 0, hbytes, st));
    D2P_CHECK_CUDA(cudaMemsetAsync(a.sync, 0, 256, st));
    a.whpk = (const uint8_t*)whpk; a.mgp_w = mgp_of(G4); a.mgp_h = mgp_h;
    a.gates = gates; a.cells = cells; a.Y = Y; a.h0 = h0; a.c0 = c0; a.hT = hT; a.cT = cT;
    a.len = len; a.R = R; a.T = T; a.forget_bias = forget_bias;
    a.rt = persist_row_tile(R, wide && !compact);
    a.full_fence = g_persist_full_fence;
    a.stacked = (R == P_STACK_ROWS && g_persist_stacked) ? 1 : 0;
    static bool attr_set = false;
    if (!attr_set) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(lstm_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)P_SMEM));
        attr_set = true;
    }
    if (compact && lstm_persist_compact_supported(R, H)) {
        FwdMtArgs am;
        am.f = a;
        am.nt = mt_partition(R, am.g0);
        switch (am.nt) {
            case 2: return launch_fwd_mt<2>(st, am);
            case 3: return launch_fwd_mt<3>(st, am);
            case 4: return launch_fwd_mt<4>(st, am);
        }
    }
    return launch_coop<FwdArgs>(lstm_persist_fwd_kernel, dim3(PCOLS, cdiv(R, a.rt)), P_SMEM, st, a);
}

// Recurrence phase of lstm_seq_bwd (without dX): gates -> dZ, dh0, dc0.
namespace {
__global__ void add_db_partials(const float* __restrict__ part, int nt, int n, float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int t = 0; t < nt; ++t) s += part[(size_t)t * n + i];
    db[i] += s;
}
}  // namespace

// The kernel also accumulates the bias gradient (column sums of dZ): db += sum over row tiles.
int lstm_persist_bwd(cudaStream_t st, int T, int R, int H, const int* len, const float* h0, const float* c0,
                     const float* Wh, float* gates, const float* cells, const float* dY, const float* dhT,
                     const float* dcT, float* dh0, float* dc0, float* db, const void** dz_full, size_t* off_out,
                     bool wide) {
    const int G4 = 4 * H;
    size_t off = 0;
    // R > 32 and whole 8-row groups per step: the kernel writes dZ_t straight into the packed
    // [T*R, 4H] operand the dX product consumes (no pack pass over dZ afterwards)
    const bool full = dz_full != nullptr && R > 32 && R % 8 == 0 &&
                      al256((size_t)kgp_of(G4) * mgp_of(T * R) * 256) + (16u << 20) <= tc_scratch_capacity(st);
    const int mgp_z = full ? mgp_of(T * R) : (R <= 32 ? cdiv(R, 8) : mgp_of(R));
    const size_t zbytes = (size_t)kgp_of(G4) * mgp_z * 256;
    BwdArgs a;
    a.rt = persist_row_tile(R, wide);
    a.full_fence = g_persist_full_fence;
    a.stacked = (R == P_STACK_ROWS && g_persist_stacked) ? 1 : 0;
    const int ntiles = cdiv(R, a.rt);
    a.full = full ? 1 : 0;
    a.dzpk = (uint8_t*)tc_scratch_alloc(st, &off, zbytes);
    a.sync = (unsigned*)tc_scratch_alloc(st, &off, 256);
    a.dbpart = (float*)tc_scratch_alloc(st, &off, (size_t)ntiles * G4 * sizeof(float));
    D2P_REQUIRE(a.dzpk && a.sync && a.dbpart, "lstm persist bwd: tensor-core scratch arena too small");
    const void* wtpk;
    D2P_TRY(get_packed(st, Wh, H, G4, G4, true, true, &off, &wtpk, 1000, 0));
    if (!full || mgp_of(T * R) * 8 != T * R) D2P_CHECK_CUDA(cudaMemsetAsync(a.dzpk, 0, zbytes, st));
    D2P_CHECK_CUDA(cudaMemsetAsync(a.sync, 0, 256, st));
    a.wtpk = (const uint8_t*)wtpk; a.mgp_w = mgp_of(H); a.mgp_z = mgp_z;
    a.gates = gates; a.cells = cells; a.c0 = c0; a.dY = dY; a.dhT = dhT; a.dcT = dcT;
    a.dh0 = dh0; a.dc0 = dc0; a.len = len; a.R = R; a.T = T; a.has_h0 = h0 != nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(lstm_persist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)P_SMEM));
        attr_set = true;
    }
    D2P_TRY(launch_coop<BwdArgs>(lstm_persist_bwd_kernel, dim3(PCOLS, ntiles), P_SMEM, st, a, 4));
    add_db_partials<<<cdiv(G4, 256), 256, 0, st>>>(a.dbpart, ntiles, G4, db);
    D2P_CHECK_LAUNCH();
    if (dz_full) *dz_full = full ? a.dzpk : nullptr;
    if (off_out) *off_out = off;
    return 0;
}

}  // namespace d2p

// 0: one launch per recurrent step; 1 (default): persistent kernels where supported.
extern "C" int d2p_lstm_set_persistent(int mode) {
    d2p::g_persist_full_fence = (mode >> 2) & 1;
    d2p::g_persist_stacked = ((mode >> 3) & 1) ? 0 : 1;
    mode &= 3;
    d2p::g_persist_mode = mode;
    return 0;
}
