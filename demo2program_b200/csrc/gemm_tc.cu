// Tensor-core GEMM engine (tcgen05 / TMEM), fp32-equivalent via bf16x3.
//
// C[M,N] = alpha * A*B + beta*C (+bias), fp32 in HBM.  tcgen05 has no fp32
// MMA, and the parity bar of this path is "training loss within 1e-4 fp32", so
// every operand x is split once into two bf16 terms, x = hi + lo (+ O(2^-17)),
// and each k-step issues three kind::f16 MMAs into the same TMEM accumulator:
//     A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi          (dropped terms ~2^-16 relative)
//
// Operand format ("packed"): the split pass writes each operand Op[mn, k]
// directly as an array of UMMA core matrices (8 mn-rows x 8 k, 128 bytes, hi
// and lo adjacent), ordered [k/8][mn/8][hi|lo]:
//     byte(mn, k, hl) = ((k/8 * MGp + mn/8) * 2 + hl) * 128 + (mn%8)*16 + (k%8)*2
// Whatever the memory order of the fp32 source (any transpose combination of
// row-major arrays), the packed operand is K-major, and a 128-row x 8-k slab of
// it is ONE contiguous 4 KB run in HBM that is already in the UMMA no-swizzle
// canonical shared-memory layout (SBO = 256 B between mn-groups, LBO = slab
// size between k-groups).  The main loop therefore needs no TMA tensor maps:
// a producer thread issues plain bulk async copies (cp.async.bulk ->
// UBLKCP) that complete on an mbarrier.
//
// Kernel: one CTA per 128 x BN tile (BN = 64/128), 4 warps, warp-specialised:
//   warp 0 / lane 0 : producer - waits "empty", arms "full" with expect_tx,
//                     issues the bulk copies of one k-block (BK = 32)
//   warp 1 / lane 0 : MMA issuer - waits "full", issues 6 tcgen05.mma
//                     (2 k-steps x {hi*hi, hi*lo, lo*hi}), tcgen05.commit ->
//                     "empty"; the last commit also signals the epilogue
//   warps 0-3       : epilogue - tcgen05.ld of the 32 TMEM lanes each warp
//                     owns, alpha/beta/bias, vectorised stores.
// 3 stages x 32 KB = 96 KB of shared memory, so two CTAs share an SM and one
// CTA's epilogue overlaps the other's main loop.
#include "common.cuh"
#include <cuda_bf16.h>

namespace d2p {

namespace {

constexpr int BM = 128, BK = 32, KG_PER_BLOCK = BK / 8;
typedef __nv_bfloat16 bf16;

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

struct Packed {
    const uint8_t* p;   // packed core-matrix array
    int mgp;            // mn-groups per k-group (padded so every tile run is in range)
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(128)
gemm_tc_kernel(Packed A, Packed B, int M, int N, int K, float alpha, float beta,
               float* __restrict__ C, int ldc, const float* __restrict__ bias) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t A_SLAB = (BM / 8) * 256, B_SLAB = (BN / 8) * 256;     // one k-group
    constexpr uint32_t A_BYTES = KG_PER_BLOCK * A_SLAB, B_BYTES = KG_PER_BLOCK * B_SLAB;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];   // full[S], empty[S], accum
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
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
    const int nk = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        // ===== producer =====
        const uint8_t* a_src = A.p + (size_t)(m0 / 8) * 256;
        const uint8_t* b_src = B.p + (size_t)(n0 / 8) * 256;
        const size_t a_kg = (size_t)A.mgp * 256, b_kg = (size_t)B.mgp * 256;
        for (int kb = 0; kb < nk; ++kb) {
            const int slot = kb % STAGES;
            if (kb >= STAGES) mbar_wait(empty0 + 8 * slot, ((kb / STAGES) - 1) & 1);
            const uint32_t bar = full0 + 8 * slot;
            mbar_expect_tx(bar, STAGE_BYTES);
            const uint32_t sa = sbase + slot * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int g = 0; g < KG_PER_BLOCK; ++g) {
                const size_t kg = (size_t)kb * KG_PER_BLOCK + g;
                bulk_copy(sa + g * A_SLAB, a_src + kg * a_kg, A_SLAB, bar);
                bulk_copy(sb + g * B_SLAB, b_src + kg * b_kg, B_SLAB, bar);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): f32 accum, bf16 x bf16, K-major A/B
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
        for (int kb = 0; kb < nk; ++kb) {
            const int slot = kb % STAGES;
            mbar_wait(full0 + 8 * slot, (kb / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = sbase + slot * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                uint64_t ahi = make_desc(sa + kk * 2 * A_SLAB, A_SLAB, 256);
                uint64_t alo = make_desc(sa + kk * 2 * A_SLAB + 128, A_SLAB, 256);
                uint64_t bhi = make_desc(sb + kk * 2 * B_SLAB, B_SLAB, 256);
                uint64_t blo = make_desc(sb + kk * 2 * B_SLAB + 128, B_SLAB, 256);
                umma_bf16(tmem_d, ahi, bhi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
                umma_bf16(tmem_d, ahi, blo, idesc, 1u);
                umma_bf16(tmem_d, alo, bhi, idesc, 1u);
            }
            umma_commit(empty0 + 8 * slot);          // stage reusable once these MMAs retire
            if (kb == nk - 1) umma_commit(accum);     // accumulator complete
        }
    }
    // ===== epilogue: TMEM -> registers -> HBM =====
    __syncwarp();
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + warp * 32 + lane;
    const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < M) {
            float* crow = C + (size_t)m * ldc;
            const int nb = n0 + c0;
            if (vec_ok && nb + 32 <= N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r;
                    r.x = alpha * __uint_as_float(v[j]);     r.y = alpha * __uint_as_float(v[j + 1]);
                    r.z = alpha * __uint_as_float(v[j + 2]); r.w = alpha * __uint_as_float(v[j + 3]);
                    if (bias) {
                        float4 b4 = *reinterpret_cast<const float4*>(bias + nb + j);
                        r.x += b4.x; r.y += b4.y; r.z += b4.z; r.w += b4.w;
                    }
                    float4* dst = reinterpret_cast<float4*>(crow + nb + j);
                    if (beta != 0.f) {
                        float4 o = *dst;
                        r.x += beta * o.x; r.y += beta * o.y; r.z += beta * o.z; r.w += beta * o.w;
                    }
                    *dst = r;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    int n = nb + j;
                    if (n < N) {
                        float r = alpha * __uint_as_float(v[j]);
                        if (bias) r += bias[n];
                        if (beta != 0.f) r += beta * crow[n];
                        crow[n] = r;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                     "r"((uint32_t)BN));
    }
}

// fp32 source -> packed bf16 hi/lo core matrices.  One thread per (kg, mg, row-in-group).
// K_CONTIG: Op[mn,k] = S[mn*ld + k]; else Op[mn,k] = S[k*ld + mn].
template <bool K_CONTIG>
__global__ void pack_bf16_kernel(const float* __restrict__ S, int MN, int K, int ld, int mgp, int kgp,
                                 uint8_t* __restrict__ out) {
    const size_t total = (size_t)kgp * mgp * 8;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx & 7);
        const size_t g = idx >> 3;
        const int mg = (int)(g % mgp), kg = (int)(g / mgp);
        const int mn = mg * 8 + r, k0 = kg * 8;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            float val = 0.f;
            if (mn < MN && k < K) val = K_CONTIG ? S[(size_t)mn * ld + k] : S[(size_t)k * ld + mn];
            x[j] = val;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bf16 h0 = __float2bfloat16_rn(x[2 * j]), h1 = __float2bfloat16_rn(x[2 * j + 1]);
            bf16 l0 = __float2bfloat16_rn(x[2 * j] - __bfloat162float(h0));
            bf16 l1 = __float2bfloat16_rn(x[2 * j + 1] - __bfloat162float(h1));
            h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        uint8_t* base = out + (g * 2) * 128 + r * 16;
        *reinterpret_cast<uint4*>(base) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(base + 128) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

template <int BN, int STAGES>
int launch_tc(cudaStream_t st, Packed A, Packed B, int M, int N, int K, float alpha, float beta,
              float* C, int ldc, const float* bias) {
    constexpr size_t smem = (size_t)STAGES * KG_PER_BLOCK * ((BM / 8) * 256 + (BN / 8) * 256);
    static bool attr_set = false;
    if (!attr_set) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    gemm_tc_kernel<BN, STAGES><<<grid, 128, smem, st>>>(A, B, M, N, K, alpha, beta, C, ldc, bias);
    D2P_CHECK_LAUNCH();
    return 0;
}

inline int mgp_of(int MN) { return (MN + 127) / 128 * 16; }          // whole 128-row tiles
inline int kgp_of(int K) { return (K + BK - 1) / BK * KG_PER_BLOCK; }  // whole k-blocks

}  // namespace

size_t packed_bytes(int MN, int K) { return (size_t)kgp_of(K) * mgp_of(MN) * 256; }

// Pack Op[mn,k] (MN x K) from an fp32 array; k_contig selects the source memory order.
int pack_bf16(cudaStream_t st, const float* S, int MN, int K, int ld, bool k_contig, void* out) {
    const int mgp = mgp_of(MN), kgp = kgp_of(K);
    size_t total = (size_t)kgp * mgp * 8;
    size_t b = (total + 255) / 256, cap = 16 * (size_t)kNumSMs;
    int blocks = (int)(b < cap ? (b < 1 ? 1 : b) : cap);
    if (k_contig)
        pack_bf16_kernel<true><<<blocks, 256, 0, st>>>(S, MN, K, ld, mgp, kgp, (uint8_t*)out);
    else
        pack_bf16_kernel<false><<<blocks, 256, 0, st>>>(S, MN, K, ld, mgp, kgp, (uint8_t*)out);
    D2P_CHECK_LAUNCH();
    return 0;
}

// C = alpha * A*B + beta*C (+bias) from packed operands (A: M x K, B: N x K).
int gemm_tc_packed(cudaStream_t st, const void* Apk, const void* Bpk, int M, int N, int K, float alpha,
                   float beta, float* C, int ldc, const float* bias) {
    Packed A{(const uint8_t*)Apk, mgp_of(M)};
    Packed B{(const uint8_t*)Bpk, mgp_of(N)};
    // Large grids: 128-wide tiles, 3 stages (96 KB) so two CTAs share an SM and one CTA's
    // epilogue overlaps the other's main loop.  Small grids (recurrent steps, dW with
    // few tiles): 64-wide tiles for more CTAs and an 8-deep ring (192 KB) because a lone
    // CTA must cover the HBM/L2 latency by itself (bytes in flight = bandwidth x latency).
    bool narrow = (long long)cdiv(N, 128) * cdiv(M, BM) < kNumSMs;
    return narrow ? launch_tc<64, 8>(st, A, B, M, N, K, alpha, beta, C, ldc, bias)
                  : launch_tc<128, 3>(st, A, B, M, N, K, alpha, beta, C, ldc, bias);
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t gemm_tc_ws_bytes(int M, int N, int K) {
    return al256(packed_bytes(M, K)) + al256(packed_bytes(N, K));
}

// fp32 operands: pack both into the workspace, then run the tensor-core kernel.
int gemm_tc(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha, const float* A,
            int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
            void* ws, size_t ws_bytes) {
    D2P_REQUIRE(ws && ws_bytes >= gemm_tc_ws_bytes(M, N, K), "gemm_tc: workspace too small");
    char* w = (char*)ws;
    void* apk = w;
    void* bpk = w + al256(packed_bytes(M, K));
    // A stored [M,K] (ta=0, k contiguous) or [K,M] (ta=1); B stored [K,N] (tb=0, n contiguous) or [N,K]
    D2P_TRY(pack_bf16(st, A, M, K, lda, !ta, apk));
    D2P_TRY(pack_bf16(st, B, N, K, ldb, tb, bpk));
    return gemm_tc_packed(st, apk, bpk, M, N, K, alpha, beta, C, ldc, bias);
}

}  // namespace d2p

extern "C" size_t d2p_gemm_tc_ws_bytes(int M, int N, int K) { return d2p::gemm_tc_ws_bytes(M, N, K); }

extern "C" int d2p_gemm_tc(int transA, int transB, int M, int N, int K, float alpha, const float* A,
                           int lda, const float* B, int ldb, float beta, float* C, int ldc,
                           const float* bias, void* ws, size_t ws_bytes, void* stream) {
    D2P_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm_tc: bad arguments");
    return d2p::gemm_tc((cudaStream_t)stream, transA != 0, transB != 0, M, N, K, alpha, A, lda, B, ldb,
                        beta, C, ldc, bias, ws, ws_bytes);
}

// ---- arena + per-step cache of packed constant operands (weights) ------------
namespace d2p {
namespace {
struct CacheEntry { const float* src; int MN, K, ld; bool k_contig; size_t off; };
struct TcState {
    char* scratch = nullptr; size_t scratch_bytes = 0;
    char* cache = nullptr; size_t cache_bytes = 0; size_t cache_used = 0;
    CacheEntry entries[256]; int n_entries = 0;
    int enabled = 1;
};
TcState g_tc;
}  // namespace

bool tc_available() { return g_tc.enabled && g_tc.scratch != nullptr; }

// Returns a packed copy of Op (MN x K); constant operands are packed once per step.
static int get_packed(cudaStream_t st, const float* S, int MN, int K, int ld, bool k_contig,
                      bool is_const, size_t* scratch_off, const void** out) {
    size_t bytes = al256(packed_bytes(MN, K));
    if (is_const && g_tc.cache) {
        for (int i = 0; i < g_tc.n_entries; ++i) {
            const CacheEntry& e = g_tc.entries[i];
            if (e.src == S && e.MN == MN && e.K == K && e.ld == ld && e.k_contig == k_contig) {
                *out = g_tc.cache + e.off;
                return 0;
            }
        }
        if (g_tc.n_entries < 256 && g_tc.cache_used + bytes <= g_tc.cache_bytes) {
            CacheEntry& e = g_tc.entries[g_tc.n_entries++];
            e = CacheEntry{S, MN, K, ld, k_contig, g_tc.cache_used};
            g_tc.cache_used += bytes;
            D2P_TRY(pack_bf16(st, S, MN, K, ld, k_contig, g_tc.cache + e.off));
            *out = g_tc.cache + e.off;
            return 0;
        }
    }
    D2P_REQUIRE(*scratch_off + bytes <= g_tc.scratch_bytes, "tensor-core scratch arena too small");
    void* dst = g_tc.scratch + *scratch_off;
    *scratch_off += bytes;
    D2P_TRY(pack_bf16(st, S, MN, K, ld, k_contig, dst));
    *out = dst;
    return 0;
}

bool tc_eligible(int M, int N, int K) {
    if (!tc_available()) return false;
    if ((double)M * N * K < (double)(1 << 18)) return false;   // tiny: SIMT engine
    return al256(packed_bytes(M, K)) + al256(packed_bytes(N, K)) <= g_tc.scratch_bytes;
}

int gemm_tc_auto(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha, const float* A,
                 int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
                 int flags) {
    size_t off = 0;
    const void *apk, *bpk;
    D2P_TRY(get_packed(st, A, M, K, lda, !ta, (flags & GEMM_CONST_A) != 0, &off, &apk));
    D2P_TRY(get_packed(st, B, N, K, ldb, tb, (flags & GEMM_CONST_B) != 0, &off, &bpk));
    return gemm_tc_packed(st, apk, bpk, M, N, K, alpha, beta, C, ldc, bias);
}
}  // namespace d2p

// scratch: packed activations (reused by every GEMM on the stream);
// cache: packed weights, valid until d2p_tc_new_step().
extern "C" int d2p_tc_configure(void* scratch, size_t scratch_bytes, void* cache, size_t cache_bytes,
                                int enabled) {
    d2p::g_tc.scratch = (char*)scratch; d2p::g_tc.scratch_bytes = scratch_bytes;
    d2p::g_tc.cache = (char*)cache; d2p::g_tc.cache_bytes = cache_bytes;
    d2p::g_tc.cache_used = 0; d2p::g_tc.n_entries = 0;
    d2p::g_tc.enabled = enabled;
    return 0;
}

extern "C" int d2p_tc_new_step(void) {
    d2p::g_tc.cache_used = 0; d2p::g_tc.n_entries = 0;
    return 0;
}
