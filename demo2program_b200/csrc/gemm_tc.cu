// Tensor-core GEMM engine (tcgen05 / TMEM), fp32-equivalent via bf16x3.
//
// C[M,N] = alpha * A*B + beta*C (+bias), fp32 in HBM.  tcgen05 has no fp32
// MMA, and the parity bar of this path is "training loss within 1e-4 fp32", so
// every operand x is split once into two bf16 terms, x = hi + lo (+ O(2^-17)),
// and each k-step issues three kind::f16 MMAs into the same TMEM accumulator:
//     A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi          (dropped terms ~2^-16 relative)
//
// Kernel shape: one CTA per 128 x BN output tile (BN = 64 or 128), 128 threads.
// All four warps stream bf16 operand tiles with 16-byte cp.async into a
// 3-stage ring of shared-memory buffers laid out in the UMMA *no-swizzle
// canonical* layout (8 x 16-byte core matrices; LBO = stride between core
// matrices along K, SBO = stride along M/N), so both K-major and MN-major
// operands (i.e. all four transpose combinations of row-major arrays) are fed
// without any transposition pass.  One elected thread issues the
// tcgen05.mma's; tcgen05.commit on an mbarrier releases each stage back to the
// loaders and finally signals the epilogue, where each warp pulls its 32 TMEM
// lanes with tcgen05.ld and applies alpha/beta/bias.
#include "common.cuh"
#include <cuda_bf16.h>
#include <map>
#include <mutex>

namespace d2p {

namespace {

constexpr int BM = 128, BK = 32, STAGES = 3;
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
        "}" ::"r"(bar), "r"(parity));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar));
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}

struct Operand {
    const bf16* hi;
    const bf16* lo;
    int ld;        // elements; multiple of 8
    int mn_total;  // extent along M (A) or N (B)
};

// Stage one ROWS x BK operand tile (hi and lo) into canonical no-swizzle layout:
//   smem byte offset of the 16-byte chunk = kgroup*LBO + mngroup*128 + inner*16
// MN_MAJOR = false: source is [mn, k] with k contiguous; chunk = 8 k of one mn row.
// MN_MAJOR = true : source is [k, mn] with mn contiguous; chunk = 8 mn of one k row.
template <int ROWS, bool MN_MAJOR>
__device__ __forceinline__ void load_tile(const Operand& op, int mn0, int k0, int K,
                                          uint32_t s_hi, uint32_t s_lo, int tid) {
    constexpr uint32_t LBO = ROWS * 16;
    constexpr int CHUNKS = ROWS * (BK / 8);
#pragma unroll
    for (int it = 0; it < CHUNKS / 128; ++it) {
        int q = it * 128 + tid;
        int inner = q & 7, c4 = (q >> 3) & 3, rest = q >> 5;
        uint32_t soff;
        size_t goff;
        int valid;   // number of valid elements in this chunk (0..8)
        if (!MN_MAJOR) {
            int mn = rest * 8 + inner, kc = c4;      // rest in [0, ROWS/8)
            int gmn = mn0 + mn, gk = k0 + kc * 8;
            soff = kc * LBO + rest * 128 + inner * 16;
            valid = (gmn < op.mn_total) ? min(8, max(0, K - gk)) : 0;
            goff = (size_t)gmn * op.ld + gk;
        } else {
            constexpr int MG = ROWS / 32;            // groups of 4 mn-chunks
            int mc = (rest % MG) * 4 + c4, kg = rest / MG;
            int k = kg * 8 + inner;
            int gk = k0 + k, gmn = mn0 + mc * 8;
            soff = kg * LBO + mc * 128 + inner * 16;
            valid = (gk < K) ? min(8, max(0, op.mn_total - gmn)) : 0;
            goff = (size_t)gk * op.ld + gmn;
        }
        if (valid == 0) goff = 0;
        cp_async16(s_hi + soff, op.hi + goff, valid * 2);
        cp_async16(s_lo + soff, op.lo + goff, valid * 2);
    }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128)
gemm_tc_kernel(Operand A, Operand B, int M, int N, int K, float alpha, float beta,
               float* __restrict__ C, int ldc, const float* __restrict__ bias) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr uint32_t LBO_A = BM * 16, LBO_B = BN * 16, SBO = 128;
    __shared__ __align__(8) uint64_t bars[STAGES + 1];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = smem_u32(&bars[0]);

    if (tid == 0) {
        for (int s = 0; s <= STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
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

    // instruction descriptor (cute::UMMA::InstrDescriptor): f32 accum, bf16 x bf16
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                               ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);

    const int nk = (K + BK - 1) / BK;
    auto stage_ptr = [&](int slot, int which) -> uint32_t {   // which: 0 Ahi 1 Alo 2 Bhi 3 Blo
        uint32_t off = slot * STAGE_BYTES;
        if (which == 1) off += A_BYTES;
        else if (which == 2) off += 2 * A_BYTES;
        else if (which == 3) off += 2 * A_BYTES + B_BYTES;
        return sbase + off;
    };
    auto load_stage = [&](int slot, int kb) {
        load_tile<BM, A_MN>(A, m0, kb * BK, K, stage_ptr(slot, 0), stage_ptr(slot, 1), tid);
        load_tile<BN, B_MN>(B, n0, kb * BK, K, stage_ptr(slot, 2), stage_ptr(slot, 3), tid);
    };

    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < nk; ++kb) {
        // k-block kb has landed once at most STAGES-2 younger groups are pending
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int slot = kb % STAGES;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                uint64_t ahi = make_desc(stage_ptr(slot, 0) + kk * 2 * LBO_A, LBO_A, SBO);
                uint64_t alo = make_desc(stage_ptr(slot, 1) + kk * 2 * LBO_A, LBO_A, SBO);
                uint64_t bhi = make_desc(stage_ptr(slot, 2) + kk * 2 * LBO_B, LBO_B, SBO);
                uint64_t blo = make_desc(stage_ptr(slot, 3) + kk * 2 * LBO_B, LBO_B, SBO);
                umma_bf16(tmem_d, ahi, bhi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
                umma_bf16(tmem_d, ahi, blo, idesc, 1u);
                umma_bf16(tmem_d, alo, bhi, idesc, 1u);
            }
            umma_commit(bar0 + 8 * slot);               // stage free once these MMAs retire
            if (kb == nk - 1) umma_commit(bar0 + 8 * STAGES);   // accumulator complete
        }
        // refill the slot k-block kb-1 used (its MMAs were issued one iteration ago, so
        // the tensor pipe keeps working on k-block kb while we wait and reload)
        const int nxt = kb + STAGES - 1;
        if (nxt < nk) {
            const int slot = nxt % STAGES;
            if (kb >= 1) mbar_wait(bar0 + 8 * slot, ((nxt / STAGES) - 1) & 1);
            load_stage(slot, nxt);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- epilogue: TMEM -> registers -> HBM ----
    mbar_wait(bar0 + 8 * STAGES, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + warp * 32 + lane;
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
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int n = n0 + c0 + j;
                if (n < N) {
                    float r = alpha * __uint_as_float(v[j]);
                    if (bias) r += bias[n];
                    if (beta != 0.f) r += beta * crow[n];
                    crow[n] = r;
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

// x (fp32 [rows, cols], leading dim ld) -> hi/lo bf16 [rows, ld_out], zero padded
__global__ void split_bf16_kernel(const float* __restrict__ X, int rows, int cols, int ld,
                                  bf16* __restrict__ hi, bf16* __restrict__ lo, int ld_out) {
    size_t total = (size_t)rows * ld_out;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(idx % ld_out);
        size_t r = idx / ld_out;
        float x = c < cols ? X[r * ld + c] : 0.f;
        bf16 h = __float2bfloat16_rn(x);
        hi[idx] = h;
        lo[idx] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}

template <int BN, bool A_MN, bool B_MN>
int launch_tc(cudaStream_t st, Operand A, Operand B, int M, int N, int K, float alpha, float beta,
              float* C, int ldc, const float* bias) {
    constexpr size_t smem = (size_t)STAGES * (2 * BM * BK * 2 + 2 * BN * BK * 2);
    static bool attr_set = false;
    if (!attr_set) {
        D2P_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    gemm_tc_kernel<BN, A_MN, B_MN><<<grid, 128, smem, st>>>(A, B, M, N, K, alpha, beta, C, ldc, bias);
    D2P_CHECK_LAUNCH();
    return 0;
}

}  // namespace

int split_bf16(cudaStream_t st, const float* X, int rows, int cols, int ld, void* hi, void* lo,
               int ld_out) {
    size_t total = (size_t)rows * ld_out;
    size_t b = (total + 255) / 256, cap = 8 * (size_t)kNumSMs;
    split_bf16_kernel<<<(int)(b < cap ? (b < 1 ? 1 : b) : cap), 256, 0, st>>>(X, rows, cols, ld, (bf16*)hi,
                                                                            (bf16*)lo, ld_out);
    D2P_CHECK_LAUNCH();
    return 0;
}

// A_mn: op(A)[m,k] is stored with m contiguous (i.e. the row-major array was transposed);
// B_mn: op(B)[k,n] is stored with n contiguous (the plain row-major [K,N] case).
int gemm_tc_presplit(cudaStream_t st, const void* Ahi, const void* Alo, int lda, bool A_mn,
                     const void* Bhi, const void* Blo, int ldb, bool B_mn, int M, int N, int K,
                     float alpha, float beta, float* C, int ldc, const float* bias) {
    D2P_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm_tc: bf16 leading dims must be multiples of 8");
    Operand A{(const bf16*)Ahi, (const bf16*)Alo, lda, M};
    Operand B{(const bf16*)Bhi, (const bf16*)Blo, ldb, N};
    // narrow tiles when the grid would otherwise leave most SMs idle
    bool narrow = (long long)cdiv(N, 128) * cdiv(M, BM) < kNumSMs;
#define D2P_TC(BN_)                                                                              \
    (A_mn ? (B_mn ? launch_tc<BN_, true, true>(st, A, B, M, N, K, alpha, beta, C, ldc, bias)      \
                  : launch_tc<BN_, true, false>(st, A, B, M, N, K, alpha, beta, C, ldc, bias))    \
          : (B_mn ? launch_tc<BN_, false, true>(st, A, B, M, N, K, alpha, beta, C, ldc, bias)     \
                  : launch_tc<BN_, false, false>(st, A, B, M, N, K, alpha, beta, C, ldc, bias)))
    return narrow ? D2P_TC(64) : D2P_TC(128);
#undef D2P_TC
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t gemm_tc_ws_bytes(int M, int N, int K) {
    size_t a = al256((size_t)(M > K ? M : K) * (size_t)(((M > K ? K : M) + 7) / 8 * 8) * 2);
    size_t b = al256((size_t)(N > K ? N : K) * (size_t)(((N > K ? K : N) + 7) / 8 * 8) * 2);
    // generous: rows x padded cols for either orientation
    a = al256((size_t)(M + 8) * (K + 8) * 2);
    b = al256((size_t)(N + 8) * (K + 8) * 2);
    return 2 * a + 2 * b;
}

// fp32 operands: split both into the workspace, then run the tensor-core kernel.
int gemm_tc(cudaStream_t st, bool ta, bool tb, int M, int N, int K, float alpha, const float* A,
            int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
            void* ws, size_t ws_bytes) {
    D2P_REQUIRE(ws && ws_bytes >= gemm_tc_ws_bytes(M, N, K), "gemm_tc: workspace too small");
    // stored shapes: A is [M,K] (ta=0) or [K,M] (ta=1); B is [K,N] (tb=0) or [N,K] (tb=1)
    int ar = ta ? K : M, ac = ta ? M : K, br = tb ? N : K, bc = tb ? K : N;
    int ald = (ac + 7) / 8 * 8, bld = (bc + 7) / 8 * 8;
    char* w = (char*)ws;
    size_t asz = al256((size_t)ar * ald * 2), bsz = al256((size_t)br * bld * 2);
    void *ahi = w, *alo = w + asz, *bhi = w + 2 * asz, *blo = w + 2 * asz + bsz;
    D2P_TRY(split_bf16(st, A, ar, ac, lda, ahi, alo, ald));
    D2P_TRY(split_bf16(st, B, br, bc, ldb, bhi, blo, bld));
    return gemm_tc_presplit(st, ahi, alo, ald, ta, bhi, blo, bld, !tb, M, N, K, alpha, beta, C, ldc,
                            bias);
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
