// mma_rate.cu — developer microbenchmark: cycles per tcgen05.mma (bf16, K = 16) as a function of N, of where the A
// operand lives (shared memory "SS" vs tensor memory "TS") and of M (64 / 128), one CTA per SM, operands = zeros.
// Answers: is the per-instruction cost of the narrow layers the shared-memory read of A?
#include <stdio.h>
#include <stdlib.h>
#include "../ptx.cuh"
using namespace fd;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

template <bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int m, int n, int iters, int two_acc, long long* out, int noswz_sbo = 0) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::make_idesc_bf16_f32(m, n);
        uint64_t adesc = ptx::make_kmajor_desc(ptx::smem_u32(smem), 1024, 2);
        uint64_t bdesc = ptx::make_kmajor_desc(ptx::smem_u32(smem) + 32768, 1024, 2);
        if (noswz_sbo) {  // un-swizzled K-major: LBO = 16 B (next K chunk), SBO = noswz_sbo (next 8-row group); B: LBO 512, SBO 128
            adesc = static_cast<uint64_t>((ptx::smem_u32(smem) >> 4) & 0x3FFF) | (1ull << 16) | (static_cast<uint64_t>(noswz_sbo >> 4) << 32) | (1ull << 46);
            bdesc = static_cast<uint64_t>(((ptx::smem_u32(smem) + 32768) >> 4) & 0x3FFF) | (32ull << 16) | (8ull << 32) | (1ull << 46);
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = tmem + ((two_acc && (u & 1)) ? 256 : 0);
                const uint32_t ks = (u & 3) * 2;  // walk the 4 k-steps of a 64-wide K block like the real mainloop
                if (TS) umma_ts(d, tmem + 480, bdesc + ks, idesc, 1u);
                else ptx::umma_bf16(d, adesc + ks, bdesc + ks, idesc, 1u);
            }
        }
        ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 4096;
    printf("%-4s %-4s %-5s %-8s %s\n", "mode", "M", "N", "two_acc", "cycles/MMA (floor N/2 at M=128)");
    for (int ts = 0; ts < 2; ++ts)
        for (int m : {128, 64})
            for (int n : {32, 64, 128, 256})
                for (int two : {0, 1}) {
                    if (two && n > 128 + 128) continue;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (ts) rate_kernel<true><<<148, 128, 100 * 1024>>>(m, n, iters, two, d);
                        else rate_kernel<false><<<148, 128, 100 * 1024>>>(m, n, iters, two, d);
                    }
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%s M=%d N=%d: %s\n", ts ? "TS" : "SS", m, n, cudaGetErrorString(e)); return 1; }
                    long long h[148];
                    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                    double s = 0;
                    for (int i = 0; i < 148; ++i) s += double(h[i]);
                    printf("%-4s %-4d %-5d %-8d %.1f\n", ts ? "TS" : "SS", m, n, two, s / 148 / iters);
                }
    for (int sbo : {160, 256, 128, 1024})
        for (int n : {32, 64}) {
            for (int rep = 0; rep < 2; ++rep) rate_kernel<false><<<148, 128, 100 * 1024>>>(128, n, iters, 0, d, sbo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("noswz sbo=%d: %s\n", sbo, cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < 148; ++i) s += double(h[i]);
            printf("un-swizzled A (LBO 16, SBO %4d), N=%d: %.1f cycles/MMA\n", sbo, n, s / 148 / iters);
        }
    return 0;
}
