// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, operands in shared memory)
// as a function of N and of the shared-memory layout the descriptors name:
//   mode 0: SWIZZLE_NONE, core matrices hi/lo interleaved (LBO 256 B, SBO 2048 B)  - libd2p's format
//   mode 1: SWIZZLE_NONE, dense core matrices (LBO 128 B, SBO 1024 B)
//   mode 2: SWIZZLE_128B K-major atoms of 8 rows x 128 B, atoms 2048 B apart (hi/lo interleaved)
//   mode 3: SWIZZLE_128B, atoms 1024 B apart (dense)
// Values are irrelevant (smem zeroed); one thread issues `n` MMAs back to back, commits, waits.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n .reg .pred P1;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra.uni WAIT_DONE;\n bra.uni WAIT_LOOP;\n WAIT_DONE:\n }" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n }"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int N, int M = 128>
__global__ void __launch_bounds__(128) rate_kernel(int mode, int n_mma, int distinct, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
        uint32_t lbo, sbo, layout, kstep;
        if (mode == 0) { lbo = 256; sbo = 2048; layout = 0; kstep = 512; }
        else if (mode == 1) { lbo = 128; sbo = 1024; layout = 0; kstep = 256; }
        else if (mode == 2) { lbo = 16; sbo = 2048; layout = 2; kstep = 32; }
        else { lbo = 16; sbo = 1024; layout = 2; kstep = 32; }
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t kk = distinct ? (uint32_t)(i & 3) : 0u;   // walk the 4 k16 steps of a 64-wide k-block
            const uint32_t blk = distinct ? (uint32_t)((i >> 2) & 1) : 0u;
            umma(tm, make_desc(sa + blk * 32768 + kk * kstep, lbo, sbo, layout),
                 make_desc(sb + blk * 65536 + kk * kstep, lbo, sbo, layout), idesc, i > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        cycles[2 * blockIdx.x] = t1 - t0;
        cycles[2 * blockIdx.x + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}
// mode 4: the persistent LSTM kernel's inner loop: per k16 one N=128 MMA (A: mode-0 packed image,
// B: slab [hi|lo] image, LBO 128 / SBO 1024) and one N=64 MMA, commit every 8 MMAs.
__global__ void __launch_bounds__(512) loop_kernel(int n_kb, int commit_every, int nthreads_wait, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 32) {
        const uint32_t i64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t i128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sw = smem_u32(smem), sa = sw + 128 * 1024;
        const long long t0 = clock64();
        int n = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
            const uint32_t a = sa + (kb & 7) * 4096, b = sw + (kb & 7) * 16384;
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ahi = make_desc(a + kk * 512, 256, 2048, 0), alo = make_desc(a + kk * 512 + 128, 256, 2048, 0);
                const uint64_t bw = make_desc(b + kk * 256, 128, 1024, 0);
                umma(tm, ahi, bw, i128, (kb | kk) != 0);
                umma(tm, alo, bw, i64, 1);
                n += 2;
                if (commit_every && n % commit_every == 0)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        mbar_wait(smem_u32(&bar[0]), 0);
        cycles[blockIdx.x] = clock64() - t0;
    } else if ((int)threadIdx.x < nthreads_wait && threadIdx.x != 32) {
        mbar_wait(smem_u32(&bar[0]), 0);     // other threads spinning on the accumulator barrier
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u));
}
void run_loop(int n_kb, int commit_every, int nwait, long long* d) {
    const int smem = 192 * 1024;
    cudaFuncSetAttribute(loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
        loop_kernel<<<96, 512, smem>>>(n_kb, commit_every, nwait, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d, sizeof(long long) * 96, cudaMemcpyDeviceToHost);
    long long tot = 0;
    for (int i = 0; i < 96; ++i) tot += h[i];
    printf("persist loop: %d k-blocks (%d MMAs), commit every %d, %d threads polling: %.0f cycles = %.1f cyc/mma\n", n_kb,
           n_kb * 8, commit_every, nwait, (double)tot / 96, (double)tot / 96 / (n_kb * 8));
}
template <int N, int M = 128>
void run(int mode, int n_mma, int grid, long long* d) {
    const int smem = 192 * 1024;
    cudaFuncSetAttribute(rate_kernel<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[2 * 148];
    for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<N, M><<<grid, 128, smem>>>(mode, n_mma, 1, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    long long issue = 0, done = 0;
    for (int i = 0; i < grid; ++i) { issue += h[2 * i]; done += h[2 * i + 1]; }
    printf("M=%3d N=%3d mode=%d grid=%3d n=%4d: issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)\n", M, N, mode, grid, n_mma,
           (double)issue / grid / n_mma, (double)done / grid / n_mma, 128 * N / 256);
}
int main() {
    long long* d;
    cudaMalloc(&d, sizeof(long long) * 2 * 148);
    run_loop(8, 0, 0, d);
    run_loop(80, 0, 0, d);
    for (int ce = 2; ce <= 64; ce *= 2) run_loop(80, ce, 0, d);
    run_loop(8, 8, 512, d);
    run<64, 64>(0, 960, 148, d);       // M = 64 tiles (half the rows): is the 64-cycle floor per instruction?
    run<128, 64>(0, 960, 148, d);
    run<256, 64>(0, 960, 148, d);
    for (int mode = 0; mode < 4; mode += 3) {
        run<64>(mode, 96, 1, d);
        run<64>(mode, 960, 148, d);
        run<128>(mode, 960, 148, d);
        run<256>(mode, 960, 148, d);
    }
    return 0;
}
