// Shared device code of the tcgen05 kernels: PTX wrappers, the packed operand
// format and the warp-specialised bulk-copy -> tcgen05.mma main loop.
//
// Packed operand Op[mn, k] (bf16 hi/lo split, K-major UMMA core matrices of
// 8 mn x 8 k = 128 B), ordered [k/64][mn/8][(k/8)%8][hi|lo]:
//     byte(mn, k, hl) = ((((k/64)*MGp + mn/8)*8 + (k/8)%8)*2 + hl)*128 + (mn%8)*16 + (k%8)*2
// MGp = mn-groups, padded to whole 128-row tiles; K padded to whole 64-wide
// k-blocks (zero filled).  A ROWS x 64 operand tile is ONE contiguous ROWS*256 B
// run in HBM that is already the no-swizzle canonical smem image (SBO = 2048 B
// between mn-groups, LBO = 256 B between k-groups), so each pipeline stage is
// fetched with two plain cp.async.bulk copies completing on an mbarrier.
// (Measured on B200: a cp.async.bulk request costs ~250 SM cycles of issue
// bandwidth whatever its size - 2 KB -> 8 B/clk, 24 KB -> 97 B/clk per SM - so
// stages are made of few, large requests: 32 KB of A + 16/32 KB of B.)
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace d2p {
namespace tc {

constexpr int BM = 128, BK = 64, KG_PER_BLOCK = BK / 8;
typedef __nv_bfloat16 bf16;

struct Packed {
    const uint8_t* p;   // packed core-matrix array
    int mgp;            // mn-groups per k-group
};

inline int mgp_of(int MN) { return (MN + 127) / 128 * 16; }
inline int kgp_of(int K) { return (K + BK - 1) / BK * KG_PER_BLOCK; }
inline size_t packed_bytes(int MN, int K) { return (size_t)kgp_of(K) * mgp_of(MN) * 256; }
inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// UMMA shared-memory descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// 16 consecutive accumulator columns of this thread's TMEM lane (row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp32 -> (hi, lo) bf16 pair packed two elements per 32-bit word
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    bf16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    bf16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    bf16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}
// write 8 consecutive k-elements of row `mn` (k0 multiple of 8) into a packed operand
__device__ __forceinline__ void store_packed8(uint8_t* base, int mgp, int mn, int k0, const float* x) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(x[2 * j], x[2 * j + 1], h[j], l[j]);
    uint8_t* p = base + (((((size_t)(k0 >> 6) * mgp + (mn >> 3)) * 8 + ((k0 >> 3) & 7)) * 2) * 128) +
                 (mn & 7) * 16;
    *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(p + 128) = make_uint4(l[0], l[1], l[2], l[3]);
}
// same for 4 consecutive k-elements (k0 multiple of 4): two 8-byte stores
__device__ __forceinline__ void store_packed4(uint8_t* base, int mgp, int mn, int k0, const float* x) {
    uint32_t h[2], l[2];
    split2(x[0], x[1], h[0], l[0]);
    split2(x[2], x[3], h[1], l[1]);
    uint8_t* p = base + (((((size_t)(k0 >> 6) * mgp + (mn >> 3)) * 8 + ((k0 >> 3) & 7)) * 2) * 128) +
                 (mn & 7) * 16 + (k0 & 7) * 2;
    *reinterpret_cast<uint2*>(p) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(p + 128) = make_uint2(l[0], l[1]);
}

// optional timeline probe (developer tool): slots of SM-clock stamps for CTA (0,0,0)
static __device__ long long* g_tc_dbg = nullptr;   // one copy per translation unit (no -rdc)
__device__ __forceinline__ void dbg_stamp(int slot) {
    if (g_tc_dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        g_tc_dbg[slot] = clock64();
}

template <int BN, int STAGES>
constexpr size_t tc_smem_bytes() {
    return (size_t)STAGES * ((BM / 8) * 2048 + (BN / 8) * 2048);
}

struct NoPrefetch {
    __device__ __forceinline__ void operator()(int, int) const {}
};

// Setup + producer + MMA issue for one 128 x BN accumulator tile over k-blocks
// [kb0, kb0 + nk).  Warp 0 / lane 0 produces, warp 1 / lane 0 issues the MMAs; the
// other warps first run `prefetch(worker_index, n_workers)` (epilogue inputs ->
// smem, overlapping the main loop).  Returns the TMEM base address once the
// accumulator is complete (all threads).  Call tc_teardown() after the epilogue.
template <int BN, int STAGES, class Prefetch = NoPrefetch>
__device__ __forceinline__ uint32_t tc_mainloop(const Packed& A, const Packed& B, int m0, int n0,
                                                int kb0, int nk, uint8_t* smem,
                                                Prefetch prefetch = Prefetch()) {
    constexpr uint32_t A_BYTES = (BM / 8) * 2048, B_BYTES = (BN / 8) * 2048;   // one k-block
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t LBO = 256, SBO = 2048;
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];   // full[S], empty[S], accum
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) dbg_stamp(0);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]),
                   accum = smem_u32(&bars[2 * STAGES]);

    if (tid == 0) {
        for (int s = 0; s < 2 * STAGES + 1; ++s) mbar_init(full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)), "r"((uint32_t)BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) dbg_stamp(1);

    if (warp == 0 && lane == 0) {
        // ===== producer =====
        const uint8_t* a_src = A.p + (size_t)(m0 / 8) * 2048;
        const uint8_t* b_src = B.p + (size_t)(n0 / 8) * 2048;
        const size_t a_kb = (size_t)A.mgp * 2048, b_kb = (size_t)B.mgp * 2048;
        const int rot = nk > 1 ? (int)((blockIdx.x * 3u + blockIdx.y * 7u) % (unsigned)nk) : 0;
        for (int i = 0; i < nk; ++i) {
            const int slot = i % STAGES;
            if (i >= STAGES) mbar_wait(empty0 + 8 * slot, ((i / STAGES) - 1) & 1);
            const uint32_t bar = full0 + 8 * slot;
            mbar_expect_tx(bar, STAGE_BYTES);
            const uint32_t sa = sbase + slot * STAGE_BYTES, sb = sa + A_BYTES;
            // CTAs that share an operand tile walk K in rotated order so that, at any
            // instant, they hit different lines/L2 slices instead of hot-spotting one
            int kr = i + rot;
            if (kr >= nk) kr -= nk;
            bulk_copy(sa, a_src + (size_t)(kb0 + kr) * a_kb, A_BYTES, bar);
            bulk_copy(sb, b_src + (size_t)(kb0 + kr) * b_kb, B_BYTES, bar);
            if (i < 20) dbg_stamp(8 + i);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): f32 accum, bf16 x bf16, K-major A/B
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
        for (int i = 0; i < nk; ++i) {
            const int slot = i % STAGES;
            mbar_wait(full0 + 8 * slot, (i / STAGES) & 1);
            if (i < 20) dbg_stamp(32 + i);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = sbase + slot * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                uint64_t ahi = make_desc(sa + kk * 2 * LBO, LBO, SBO);
                uint64_t alo = make_desc(sa + kk * 2 * LBO + 128, LBO, SBO);
                uint64_t bhi = make_desc(sb + kk * 2 * LBO, LBO, SBO);
                uint64_t blo = make_desc(sb + kk * 2 * LBO + 128, LBO, SBO);
                umma_bf16(tmem_d, ahi, bhi, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                umma_bf16(tmem_d, ahi, blo, idesc, 1u);
                umma_bf16(tmem_d, alo, bhi, idesc, 1u);
            }
            umma_commit(empty0 + 8 * slot);          // stage reusable once these MMAs retire
            if (i == nk - 1) umma_commit(accum);      // accumulator complete
        }
    } else if (warp >= 2) {
        prefetch((warp - 2) * 32 + lane, ((int)blockDim.x / 32 - 2) * 32);
    }
    __syncwarp();
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 64) dbg_stamp(2);
    return tmem_d;
}

template <int BN>
__device__ __forceinline__ void tc_teardown(uint32_t tmem_d) {
    if (threadIdx.x == 64) dbg_stamp(3);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) dbg_stamp(4);
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                     "r"((uint32_t)BN));
    }
}

}  // namespace tc

// host-side API of gemm_tc.cu
int pack_bf16(cudaStream_t st, const float* S, int MN, int K, int ld, bool k_contig, void* out,
              int gate_tile = 0, int gate_H = 0, int mgp_override = 0);
int gemm_tc_packed(cudaStream_t st, const void* Apk, const void* Bpk, int M, int N, int K, float alpha,
                   float beta, float* C, int ldc, const float* bias, int ksplit = 1,
                   float* partials = nullptr);
int get_packed(cudaStream_t st, const float* S, int MN, int K, int ld, bool k_contig, bool is_const,
               size_t* scratch_off, const void** out, int gate_tile = 0, int gate_H = 0);
void* tc_scratch_alloc(cudaStream_t st, size_t* scratch_off, size_t bytes);
size_t tc_scratch_capacity(cudaStream_t st);
int gemm_tc_packed_auto(cudaStream_t st, const void* Apk, const void* Bpk, int M, int N, int K, float alpha,
                        float beta, float* C, int ldc, size_t* scratch_off);

}  // namespace d2p
