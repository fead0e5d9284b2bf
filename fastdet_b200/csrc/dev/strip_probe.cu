// dev probe for the round-2 conv design ("flat halo strip"): can one im2col TMA load, with a bounding box one pixel
// larger than the image on every side, deliver the zero-padded, flattened pixel strip of a tile, so that all nine taps
// of a 3x3 convolution are UMMA descriptor row offsets into that ONE strip — and can the result leave through an
// im2col_no_offs TMA store that drops the pad positions?  Three hardware questions in one kernel:
//   1. cuTensorMapEncodeIm2col with pixelBoxUpperCorner = +1 (traversal width W + 2, height H + 2, OOB = zeros)
//   2. SWIZZLE_128B K-major descriptors whose start is not 1024-byte aligned (row offsets 0..2(W+2)+2, base-offset field)
//   3. cp.async.bulk.tensor.4d.global.shared::cta.im2col_no_offs stores over the same bounding box
//   build/strip_probe            (prints PASS / FAIL per question it could check)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../ptx.cuh"

using namespace fd;

constexpr int NB = 3, H = 6, W = 10, C = 64, COUT = 64;
constexpr int WP = W + 2, HP = H + 2;
constexpr int STRIP = 128 + 2 * WP + 2;  // rows of the strip one 128-position tile needs
constexpr int Q_FIRST = WP + 1;          // first padded-flat position that can be a real pixel

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, int use_base_offset) {
    uint64_t d = ptx::make_kmajor_desc(addr, 1024, 2);
    if (use_base_offset) d |= static_cast<uint64_t>((addr >> 7) & 7) << 49;
    return d;
}

__device__ __forceinline__ void tma_store_im2col_4d(const CUtensorMap* m, const void* src, int c, int w, int h, int n) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.im2col_no_offs.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(ptx::smem_u32(src)), "r"(c), "r"(w), "r"(h), "r"(n)
                 : "memory");
}

__global__ void __launch_bounds__(128)
strip_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
             int use_base_offset, float* dbg, int do_store) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* strip = smem;                               // STRIP rows x 128 B (SW128 by the TMA)
    uint8_t* bsm = smem + ((STRIP * 128 + 1023) / 1024) * 1024;  // 9 taps x 64 filters x 128 B
    uint8_t* stage = bsm + 9 * 8192;                     // 128 rows x 128 B
    __shared__ uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        ptx::mbar_init(&full_bar, 1);
        ptx::mbar_init(&done_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 64); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int q0 = Q_FIRST + 128 * blockIdx.x;  // first padded-flat output position of this tile
    if (tid == 0) {
        const int qs = q0 - WP - 1;  // strip start (>= 0 by construction)
        const int n = qs / (HP * WP), rem = qs - n * HP * WP, yp = rem / WP, xp = rem - yp * WP;
        ptx::mbar_arrive_expect_tx(&full_bar, STRIP * 128 + 9 * 8192);
        ptx::tma_load_im2col_4d(strip, &tmA, &full_bar, 0, xp - 1, yp - 1, n, 0, 0);
        for (int tap = 0; tap < 9; ++tap) ptx::tma_load_2d(bsm + tap * 8192, &tmB, &full_bar, tap * 64, 0);
        ptx::mbar_wait(&full_bar, 0);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, COUT);
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - 3 * ky;
            const uint32_t a_addr = ptx::smem_u32(strip) + (ky * WP + kx) * 128;
            const uint64_t adesc = desc_sw128(a_addr, use_base_offset);
            const uint64_t bdesc = desc_sw128(ptx::smem_u32(bsm) + tap * 8192, 0);
            for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | k) ? 1u : 0u);
        }
        ptx::umma_commit(&done_bar);
    }
    __syncthreads();
    ptx::mbar_wait(&done_bar, 0);
    ptx::tc_fence_after();
    uint32_t acc[32];
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < COUT; c0 += 32) {
        ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, acc);
        ptx::tmem_ld_wait();
        if (dbg) for (int j = 0; j < 32; ++j) dbg[(size_t(blockIdx.x) * 128 + row) * COUT + c0 + j] = __uint_as_float(acc[j]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 8 bf16 = 16 B per chunk; chunk index within the 128-byte row: c0/8 + c
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 2 * j]), __uint_as_float(acc[8 * c + 2 * j + 1]));
                pk[j] = *reinterpret_cast<const uint32_t*>(&v);
            }
            const int chunk = c0 / 8 + c;
            *reinterpret_cast<uint4*>(stage + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0 && do_store) {
        const int n = q0 / (HP * WP), rem = q0 - n * HP * WP, yp = rem / WP, xp = rem - yp * WP;
        tma_store_im2col_4d(&tmOut, stage, 0, xp - 1, yp - 1, n);
        ptx::tma_store_commit();
        ptx::tma_store_wait<0>();
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(uint16_t v) { uint32_t u = uint32_t(v) << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t to_bf(float f) { uint32_t u; memcpy(&u, &f, 4); return uint16_t(u >> 16); }

int main() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    PFN_encodeTiled encTiled = reinterpret_cast<PFN_encodeTiled>(f);
    cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q);
    PFN_encodeIm2col encIm2col = reinterpret_cast<PFN_encodeIm2col>(f);
    if (!encTiled || !encIm2col) { printf("no tensor-map encoders\n"); return 1; }

    const size_t nx = size_t(NB) * H * W * C, nw = size_t(9) * COUT * C, ny = size_t(NB) * H * W * COUT, guard = 4096;
    std::vector<uint16_t> hx(nx), hw(nw), hy(ny + guard);
    srand(7);
    for (auto& v : hx) v = to_bf(float(rand() % 5 - 2));
    for (auto& v : hw) v = to_bf(float(rand() % 3 - 1));
    // weights as the B operand: [cout][K = tap*64 + c]
    std::vector<uint16_t> hb(nw);
    for (int t = 0; t < 9; ++t)
        for (int o = 0; o < COUT; ++o)
            for (int c = 0; c < C; ++c) hb[size_t(o) * 9 * C + t * C + c] = hw[(size_t(t) * COUT + o) * C + c];
    uint16_t *dx, *db, *dy;
    // big allocations (> 128 KiB) keep clear of the small-tensor im2col driver quirk conv_tc.cu works around
    cudaMalloc(&dx, 1 << 20); cudaMalloc(&db, 1 << 20); cudaMalloc(&dy, 1 << 20);
    cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), nw * 2, cudaMemcpyHostToDevice);

    CUtensorMap tmA, tmB, tmOut;
    cuuint64_t dims[4] = {C, W, H, NB};
    cuuint64_t strides[3] = {C * 2, cuuint64_t(C) * 2 * W, cuuint64_t(C) * 2 * W * H};
    int lower[2] = {-1, -1}, upper[2] = {1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encIm2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, C, STRIP, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("Q1 encode im2col load map with upper corner +1, %d pixels per column: %s (CUresult %d)\n", STRIP, r == CUDA_SUCCESS ? "ok" : "REFUSED", int(r));
    if (r != CUDA_SUCCESS) return 2;
    cuuint64_t odims[4] = {COUT, W, H, NB};
    cuuint64_t ostrides[3] = {COUT * 2, cuuint64_t(COUT) * 2 * W, cuuint64_t(COUT) * 2 * W * H};
    r = encIm2col(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, odims, ostrides, lower, upper, COUT, 128, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("Q3 encode im2col store map: %s (CUresult %d)\n", r == CUDA_SUCCESS ? "ok" : "REFUSED", int(r));
    if (r != CUDA_SUCCESS) return 2;
    cuuint64_t bdims[2] = {9 * C, COUT};
    cuuint64_t bstrides[1] = {9 * C * 2};
    cuuint32_t bbox[2] = {64, 64}, bes[2] = {1, 1};
    r = encTiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, bdims, bstrides, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("B map encode failed %d\n", int(r)); return 2; }

    // CPU reference
    std::vector<float> ref(ny, 0.f);
    for (int n = 0; n < NB; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int o = 0; o < COUT; ++o) {
                    float s = 0.f;
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx) {
                            const int yy = y + ky - 1, xx = x + kx - 1;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            for (int c = 0; c < C; ++c)
                                s += bf(hx[((size_t(n) * H + yy) * W + xx) * C + c]) * bf(hw[(size_t(ky * 3 + kx) * COUT + o) * C + c]);
                        }
                    ref[((size_t(n) * H + y) * W + x) * COUT + o] = s;
                }
    const int total_q = NB * HP * WP;
    const int tiles = (total_q - Q_FIRST + 127) / 128;
    const size_t smem_bytes = 1024 + ((STRIP * 128 + 1023) / 1024) * 1024 + 9 * 8192 + 128 * 128;
    cudaFuncSetAttribute(strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
    float* ddbg;
    cudaMalloc(&ddbg, size_t(tiles) * 128 * COUT * 4);
    for (int use_bo = 1; use_bo >= 0; --use_bo) {
        for (auto& v : hy) v = 0x7fc0;  // NaN sentinel
        cudaMemcpy(dy, hy.data(), hy.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(ddbg, 0, size_t(tiles) * 128 * COUT * 4);
        strip_kernel<<<tiles, 128, smem_bytes>>>(tmA, tmB, tmOut, use_bo, ddbg, getenv("NO_STORE") ? 0 : 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed (base offset %d): %s\n", use_bo, cudaGetErrorString(e)); return 3; }
        std::vector<float> hd(size_t(tiles) * 128 * COUT);
        cudaMemcpy(hd.data(), ddbg, hd.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost);
        // accumulators (questions 1 + 2): position q of tile t, row r = q - q0
        long bad_acc = 0, checked = 0;
        for (int t = 0; t < tiles; ++t)
            for (int rr = 0; rr < 128; ++rr) {
                const int qq = Q_FIRST + 128 * t + rr;
                if (qq >= total_q) continue;
                const int n = qq / (HP * WP), rem = qq - n * HP * WP, yp = rem / WP, xp = rem - yp * WP;
                if (yp < 1 || yp > H || xp < 1 || xp > W) continue;
                for (int o = 0; o < COUT; ++o) {
                    ++checked;
                    if (hd[(size_t(t) * 128 + rr) * COUT + o] != ref[((size_t(n) * H + yp - 1) * W + xp - 1) * COUT + o]) ++bad_acc;
                }
            }
        printf("Q1+Q2 strip load + tap offsets by descriptor (base-offset field %s): %ld of %ld accumulators wrong -> %s\n",
               use_bo ? "set" : "zero", bad_acc, checked, bad_acc ? "FAIL" : "PASS");
        long bad_out = 0, untouched = 0, guard_hit = 0;
        for (size_t i = 0; i < ny; ++i) {
            if (hy[i] == 0x7fc0) { ++untouched; continue; }
            if (bf(hy[i]) != bf(to_bf(ref[i]))) ++bad_out;
        }
        for (size_t i = ny; i < ny + guard; ++i) guard_hit += hy[i] != 0x7fc0;
        printf("Q3 im2col_no_offs store: %ld wrong, %ld never written, %ld guard elements overwritten -> %s\n", bad_out, untouched,
               guard_hit, (bad_out || untouched || guard_hit) ? "FAIL" : "PASS");
    }
    return 0;
}
