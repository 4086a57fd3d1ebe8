// Microbenchmark: per-SM throughput of cp.async.bulk (UBLKCP) global->shared vs
// request size and number of requests in flight; and LDGSTS (cp.async 16B).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n .reg .pred P1;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra.uni WAIT_DONE;\n bra.uni WAIT_LOOP;\n WAIT_DONE:\n }" ::"r"(bar), "r"(parity) : "memory");
}
__global__ void bulk_kernel(const uint8_t* src, size_t src_bytes, int req_bytes, int n_req, int stages, long long* cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[16];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        size_t off = (size_t)blockIdx.x * 1048576 % src_bytes;
        for (int i = 0; i < n_req; ++i) {
            int slot = i % stages;
            uint32_t bar = smem_u32(&bars[slot]);
            if (i >= stages) mbar_wait(bar, ((i / stages) - 1) & 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(req_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + (size_t)slot * req_bytes)), "l"(src + off), "r"(req_bytes), "r"(bar) : "memory");
            off += req_bytes; if (off + req_bytes > src_bytes) off = 0;
        }
        for (int i = (n_req > stages ? n_req - stages : 0); i < n_req; ++i)
            mbar_wait(smem_u32(&bars[i % stages]), (i / stages) & 1);
        cycles[blockIdx.x] = clock64() - t0;
    }
}
__global__ void ldgsts_kernel(const uint8_t* src, size_t src_bytes, int req_bytes, int n_req, int stages, long long* cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    long long t0 = clock64();
    size_t off = (size_t)blockIdx.x * 1048576 % src_bytes;
    for (int i = 0; i < n_req; ++i) {
        int slot = i % stages;
        for (int b = threadIdx.x * 16; b < req_bytes; b += blockDim.x * 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + (size_t)slot * req_bytes + b)), "l"(src + off + b) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (stages == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else if (stages == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (stages == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else asm volatile("cp.async.wait_group 7;" ::: "memory");
        off += req_bytes; if (off + req_bytes > src_bytes) off = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
    size_t src_bytes = 64u << 20;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    long long* cyc; cudaMallocManaged(&cyc, 148 * sizeof(long long));
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int grids[] = {1, 32, 148};
    for (int mode = 0; mode < 2; ++mode)
    for (int g : grids)
    for (int req : {2048, 4096, 8192, 16384, 24576}) {
        for (int st : {1, 2, 4, 8}) {
            if ((size_t)req * st > 196608) continue;
            int n_req = (4 << 20) / req;   // 4 MB per CTA
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) bulk_kernel<<<g, 128, req * st>>>(src, src_bytes, req, n_req, st, cyc);
                else ldgsts_kernel<<<g, 128, req * st>>>(src, src_bytes, req, n_req, st, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long mx = 0; for (int i = 0; i < g; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
            printf("%s grid %3d req %5d B stages %d : %.1f B/clk/SM\n", mode ? "ldgsts" : "bulk  ", g, req, st, (double)req * n_req / mx);
        }
    }
    return 0;
}
