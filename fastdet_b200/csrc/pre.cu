// pre.cu — frame ingest kernels (HBM-bound; coalesced / vectorised byte traffic, no tensor cores).
//
//   normalise_f32_nchw : reference server/detector.py:133-134  (u8 HWC -> x/255 in f64 -> f32 -> NCHW)
//   letterbox_u8       : extension (the reference rejects non-net-sized frames, detector.py:131-132)
//   conv0_u8           : the same normalisation fused into the first convolution (Cin = 3), so the f32
//                        NCHW tensor the reference materialises never exists on the serving path.
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace fd {

// float32(k / 255.0): the reference divides in float64 and rounds once to float32.
__device__ __forceinline__ float norm_u8(int k) { return static_cast<float>(static_cast<double>(k) / 255.0); }

__global__ void __launch_bounds__(256)
normalise_f32_nchw_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int n, int hw) {
    __shared__ float lut[256];
    lut[threadIdx.x] = norm_u8(threadIdx.x);
    __syncthreads();
    const long long quads_per_frame = hw / 4;  // host guarantees hw % 4 == 0 on this path
    const long long total = quads_per_frame * n;
    for (long long q = blockIdx.x * 256LL + threadIdx.x; q < total; q += 256LL * gridDim.x) {
        const long long f = q / quads_per_frame;
        const long long p = (q - f * quads_per_frame) * 4;  // first pixel of the quad inside the frame
        const uint32_t* src = reinterpret_cast<const uint32_t*>(frames + (f * hw + p) * 3);
        const uint32_t a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);  // 12 bytes = 4 RGB pixels
        // bytes: a = r0 g0 b0 r1 | b = g1 b1 r2 g2 | c = b2 r3 g3 b3
        const float4 r = make_float4(lut[a & 255], lut[a >> 24], lut[(b >> 16) & 255], lut[(c >> 8) & 255]);
        const float4 g = make_float4(lut[(a >> 8) & 255], lut[b & 255], lut[b >> 24], lut[(c >> 16) & 255]);
        const float4 bl = make_float4(lut[(a >> 16) & 255], lut[(b >> 8) & 255], lut[c & 255], lut[c >> 24]);
        float* o = out + f * 3 * hw + p;
        *reinterpret_cast<float4*>(o) = r;
        *reinterpret_cast<float4*>(o + hw) = g;
        *reinterpret_cast<float4*>(o + 2LL * hw) = bl;
    }
}

__global__ void __launch_bounds__(256)
normalise_f32_nchw_scalar_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int n, int hw) {
    const long long total = 1LL * n * hw;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long f = i / hw, p = i - f * hw;
        const uint8_t* s = frames + i * 3;
        float* o = out + f * 3 * hw + p;
        o[0] = norm_u8(s[0]);
        o[hw] = norm_u8(s[1]);
        o[2LL * hw] = norm_u8(s[2]);
    }
}

int launch_normalise_f32_nchw(const uint8_t* frames, float* out, int n, int h, int w, cudaStream_t s) {
    const int hw = h * w;
    const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const long long work = vec ? 1LL * n * hw / 4 : 1LL * n * hw;
    const int blocks = static_cast<int>(work / 256 + 1 < 148 * 16 ? work / 256 + 1 : 148 * 16);
    if (vec) normalise_f32_nchw_kernel<<<blocks, 256, 0, s>>>(frames, out, n, hw);
    else normalise_f32_nchw_scalar_kernel<<<blocks, 256, 0, s>>>(frames, out, n, hw);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------ letterbox
__device__ __forceinline__ void lb_axis(int d, int n_dst, int n_src, int* i0, int* i1, int* fr) {
    long long pos = ((2LL * d + 1) * n_src * 65536LL) / (2LL * n_dst) - 32768LL;
    const long long hi = (n_src - 1) * 65536LL;
    pos = pos < 0 ? 0 : (pos > hi ? hi : pos);
    *i0 = static_cast<int>(pos >> 16);
    *fr = static_cast<int>(pos & 0xFFFF);
    *i1 = *i0 + 1 < n_src ? *i0 + 1 : n_src - 1;
}

__global__ void __launch_bounds__(256)
letterbox_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int sh, int sw, int h, int w,
                    int new_w, int new_h, int off_x, int off_y, int fill) {
    const long long total = 1LL * n * h * w;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int f = static_cast<int>(i / (1LL * h * w));
        const int rem = static_cast<int>(i - 1LL * f * h * w);
        const int y = rem / w, x = rem - y * w;
        uint8_t* o = dst + i * 3;
        const int dx = x - off_x, dy = y - off_y;
        if (dx < 0 || dx >= new_w || dy < 0 || dy >= new_h) {
            o[0] = o[1] = o[2] = static_cast<uint8_t>(fill);
            continue;
        }
        int x0, x1, fx, y0, y1, fy;
        lb_axis(dx, new_w, sw, &x0, &x1, &fx);
        lb_axis(dy, new_h, sh, &y0, &y1, &fy);
        const uint8_t* base = src + 1LL * f * sh * sw * 3;
        const uint8_t* p00 = base + (1LL * y0 * sw + x0) * 3;
        const uint8_t* p01 = base + (1LL * y0 * sw + x1) * 3;
        const uint8_t* p10 = base + (1LL * y1 * sw + x0) * 3;
        const uint8_t* p11 = base + (1LL * y1 * sw + x1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long top = 1LL * p00[c] * (65536 - fx) + 1LL * p01[c] * fx;
            const long long bot = 1LL * p10[c] * (65536 - fx) + 1LL * p11[c] * fx;
            o[c] = static_cast<uint8_t>((top * (65536 - fy) + bot * fy + (1LL << 31)) >> 32);
        }
    }
}

int launch_letterbox_u8(const uint8_t* src, uint8_t* dst, int n, int sh, int sw, int h, int w, int fill,
                        cudaStream_t s) {
    int new_w, new_h;
    if (1LL * sw * h >= 1LL * sh * w) {
        new_w = w;
        new_h = static_cast<int>((1LL * sh * w + sw / 2) / sw);
        if (new_h < 1) new_h = 1;
    } else {
        new_h = h;
        new_w = static_cast<int>((1LL * sw * h + sh / 2) / sh);
        if (new_w < 1) new_w = 1;
    }
    const int off_x = (w - new_w) / 2, off_y = (h - new_h) / 2;
    const long long total = 1LL * n * h * w;
    const int blocks = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    letterbox_u8_kernel<<<blocks, 256, 0, s>>>(src, dst, n, sh, sw, h, w, new_w, new_h, off_x, off_y, fill);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------ first conv
// Tile: 32 x 16 output pixels per CTA of 128 threads; each thread owns a 4-pixel horizontal strip and
// walks the output channels 8 at a time, so every 2 broadcast LDS.128 of weights feed 32 FMAs.
static constexpr int C0_TW = 32, C0_TH = 16, C0_THREADS = 128;

__device__ __forceinline__ uint32_t c0_pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(C0_THREADS)
conv0_u8_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ wgt, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ out, int h, int wd, int cout, int out_pitch, int act, float alpha) {
    extern __shared__ __align__(16) uint8_t c0_smem[];
    float* lut = reinterpret_cast<float*>(c0_smem);             // 256
    float* in_s = lut + 256;                                     // (TH+2) x (TW+2) x 3
    float* w_s = in_s + (C0_TH + 2) * (C0_TW + 2) * 3 + 4;       // 27 x cout   (+4 keeps 16-byte alignment)
    float* b_s = w_s + 27 * cout;                                // cout
    __nv_bfloat16* o_s = reinterpret_cast<__nv_bfloat16*>(b_s + cout);  // TH*TW x cout

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * C0_TW, y0 = blockIdx.y * C0_TH, f = blockIdx.z;
    for (int i = tid; i < 256; i += C0_THREADS) lut[i] = norm_u8(i);
    for (int i = tid; i < 27 * cout; i += C0_THREADS) w_s[i] = __ldg(wgt + i);
    for (int i = tid; i < cout; i += C0_THREADS) b_s[i] = __ldg(bias + i);
    __syncthreads();
    const uint8_t* fr = frames + 1LL * f * h * wd * 3;
    constexpr int ROW = (C0_TW + 2) * 3;
    for (int i = tid; i < (C0_TH + 2) * ROW; i += C0_THREADS) {
        const int yy = i / ROW, rem = i - yy * ROW;
        const int xx = rem / 3;
        const int gy = y0 - 1 + yy, gx = x0 - 1 + xx;
        float v = 0.f;  // zero padding of the *normalised* input, as Conv pads=1 does
        if (gy >= 0 && gy < h && gx >= 0 && gx < wd) v = lut[__ldg(fr + (1LL * gy * wd + x0 - 1) * 3 + rem)];
        in_s[i] = v;
    }
    __syncthreads();

    const int tx = tid & 7, ty = tid >> 3;
    float xin[3][6][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) xin[r][c][ci] = in_s[((ty + r) * (C0_TW + 2) + 4 * tx + c) * 3 + ci];

    for (int co0 = 0; co0 < cout; co0 += 8) {
        float acc[4][8];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = b_s[co0 + j];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w_s + ((r * 3 + s) * 3 + ci) * cout + co0);
                    const float4 w1 = *reinterpret_cast<const float4*>(w_s + ((r * 3 + s) * 3 + ci) * cout + co0 + 4);
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float x = xin[r][p + s][ci];
                        acc[p][0] = fmaf(x, w0.x, acc[p][0]); acc[p][1] = fmaf(x, w0.y, acc[p][1]);
                        acc[p][2] = fmaf(x, w0.z, acc[p][2]); acc[p][3] = fmaf(x, w0.w, acc[p][3]);
                        acc[p][4] = fmaf(x, w1.x, acc[p][4]); acc[p][5] = fmaf(x, w1.y, acc[p][5]);
                        acc[p][6] = fmaf(x, w1.z, acc[p][6]); acc[p][7] = fmaf(x, w1.w, acc[p][7]);
                    }
                }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (act) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[p][j] = acc[p][j] > 0.f ? acc[p][j] : acc[p][j] * alpha;
            }
            uint4 q;
            q.x = c0_pack(acc[p][0], acc[p][1]); q.y = c0_pack(acc[p][2], acc[p][3]);
            q.z = c0_pack(acc[p][4], acc[p][5]); q.w = c0_pack(acc[p][6], acc[p][7]);
            *reinterpret_cast<uint4*>(o_s + (ty * C0_TW + 4 * tx + p) * cout + co0) = q;
        }
    }
    __syncthreads();
    // coalesced write-out: consecutive threads -> consecutive 16-byte chunks of consecutive pixels
    const int chunks = cout / 8;
    for (int i = tid; i < C0_TH * C0_TW * chunks; i += C0_THREADS) {
        const int pix = i / chunks, ch = i - pix * chunks;
        const int py = pix / C0_TW, px = pix - py * C0_TW;
        const int gy = y0 + py, gx = x0 + px;
        if (gy < h && gx < wd) {
            const uint4 v = *reinterpret_cast<const uint4*>(o_s + pix * cout + ch * 8);
            *reinterpret_cast<uint4*>(out + ((1LL * f * h + gy) * wd + gx) * out_pitch + ch * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------ first conv on tensor cores
// The same layer as conv0_u8_kernel as an implicit GEMM on tcgen05: M = 128 output pixels (a 32 x 4 patch),
// K = 27 taps*channels padded to 32, N = Cout (16 or 32).  Each thread builds the bf16 im2col row of its pixel
// from a u8 halo patch in shared memory through a 257-entry LUT (bf16(k/255); entry 256 = the zero padding),
// writes it K-major with the 64-byte swizzle the UMMA descriptor expects, one thread issues two 128xNx16 MMAs,
// and every warp drains its TMEM lane quarter: bias + LeakyReLU + bf16, 64 contiguous bytes per pixel, 2 KB per
// warp.  HBM-bound by the 64 B/pixel it writes; several small CTAs per SM overlap build / MMA / drain.
static constexpr int C0T_TW = 32, C0T_TH = 4, C0T_THREADS = 128;

__global__ void __launch_bounds__(C0T_THREADS)
conv0_tc_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ wgt, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ out, int n, int h, int wd, int cout, int out_pitch, int act, float alpha) {
    __shared__ __align__(1024) uint8_t sA[128 * 64];   // 128 pixel rows x 32 bf16 (SWIZZLE_64B, K-major)
    __shared__ __align__(1024) uint8_t sB[32 * 64];    // up to 32 filter rows x 32 bf16
    __shared__ uint16_t lut[257];
    __shared__ uint16_t patch[(C0T_TH + 2) * (C0T_TW + 2) * 3];
    __shared__ float b_s[32];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 257; i += C0T_THREADS) {
        const __nv_bfloat16 v = __float2bfloat16(i < 256 ? norm_u8(i) : 0.f);
        lut[i] = *reinterpret_cast<const uint16_t*>(&v);
    }
    if (tid < 32) b_s[tid] = tid < cout ? __ldg(bias + tid) : 0.f;
    // B: row = output channel, 32 K values (27 real: (r*3+s)*3+ci), chunk c of row m at slot c ^ ((m >> 1) & 3)
    for (int i = tid; i < 32 * 4; i += C0T_THREADS) {
        const int m = i >> 2, c = i & 3;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k0 = c * 8 + 2 * e, k1 = k0 + 1;
            const float f0 = (m < cout && k0 < 27) ? __ldg(wgt + k0 * cout + m) : 0.f;
            const float f1 = (m < cout && k1 < 27) ? __ldg(wgt + k1 * cout + m) : 0.f;
            w[e] = c0_pack(f0, f1);
        }
        *reinterpret_cast<uint4*>(sB + m * 64 + ((c ^ ((m >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid == 0) {
        ptx::mbar_init(&mma_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&tmem_slot, 32);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    ptx::grid_dep_launch();  // the next layer may start its set-up on SMs as they drain (it waits before reading)
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, cout);
    const uint64_t adesc = ptx::make_kmajor_desc(ptx::smem_u32(sA), 512, 4);
    const uint64_t bdesc = ptx::make_kmajor_desc(ptx::smem_u32(sB), 512, 4);

    const int tiles_x = (wd + C0T_TW - 1) / C0T_TW, tiles_y = (h + C0T_TH - 1) / C0T_TH;
    const long long total = 1LL * n * tiles_x * tiles_y;
    constexpr int PROW = (C0T_TW + 2) * 3;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int f = static_cast<int>(tile / (tiles_x * tiles_y));
        const int rem = static_cast<int>(tile - 1LL * f * tiles_x * tiles_y);
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int x0 = tx * C0T_TW, y0 = ty * C0T_TH;
        const uint8_t* fr = frames + 1LL * f * h * wd * 3;
        for (int i = tid; i < (C0T_TH + 2) * PROW; i += C0T_THREADS) {
            const int yy = i / PROW, off = i - yy * PROW;
            const int gy = y0 - 1 + yy, gx = x0 - 1 + off / 3;
            uint16_t v = 256;  // zero padding of the normalised input
            if (gy >= 0 && gy < h && gx >= 0 && gx < wd) v = __ldg(fr + (1LL * gy * wd + x0 - 1) * 3 + off);
            patch[i] = v;
        }
        __syncthreads();
        {   // im2col row of pixel (px = lane, py = warp)
            uint32_t kk[16];
            uint16_t e[32];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 9; ++q) e[r * 9 + q] = lut[patch[(warp + r) * PROW + lane * 3 + q]];
#pragma unroll
            for (int q = 27; q < 32; ++q) e[q] = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) kk[q] = static_cast<uint32_t>(e[2 * q]) | (static_cast<uint32_t>(e[2 * q + 1]) << 16);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(sA + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) =
                    make_uint4(kk[4 * c], kk[4 * c + 1], kk[4 * c + 2], kk[4 * c + 3]);
        }
        ptx::fence_proxy_async();  // generic-proxy writes of sA -> visible to the tensor core (async proxy)
        ptx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            ptx::tc_fence_after();
            ptx::umma_bf16(tmem, adesc, bdesc, idesc, 0u);
            ptx::umma_bf16(tmem, adesc + 2, bdesc + 2, idesc, 1u);
            ptx::umma_commit(&mma_bar);
        }
        ptx::mbar_wait(&mma_bar, phase);
        phase ^= 1;
        ptx::tc_fence_after();
        uint32_t acc[32];
        ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), acc);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        const int gy = y0 + warp, gx = x0 + lane;
        if (gy < h && gx < wd) {
            __nv_bfloat16* op = out + ((1LL * f * h + gy) * wd + gx) * out_pitch;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
                if (c >= cout) break;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float x = __uint_as_float(acc[c + q]) + b_s[c + q];
                    v[q] = (act && x < 0.f) ? x * alpha : x;
                }
                *reinterpret_cast<uint4*>(op + c) =
                    make_uint4(c0_pack(v[0], v[1]), c0_pack(v[2], v[3]), c0_pack(v[4], v[5]), c0_pack(v[6], v[7]));
            }
        }
        // the next tile's __syncthreads (after the patch load) orders these TMEM reads and the patch/sA reuse
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 32);
    }
}

int kernels_init() {
    return cudaFuncSetAttribute(conv0_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) ==
                   cudaSuccess
               ? 0
               : -1;
}

int launch_conv0_u8(const uint8_t* frames, const float* w, const float* bias, __nv_bfloat16* out, int n, int h,
                    int wd, int cout, int out_pitch, int act, float alpha, cudaStream_t s) {
    static const bool force_cuda_core = getenv("FASTDET_CONV0_FMA") != nullptr;
    if (!force_cuda_core && (cout == 16 || cout == 32)) {
        const long long tiles = 1LL * n * ((wd + C0T_TW - 1) / C0T_TW) * ((h + C0T_TH - 1) / C0T_TH);
        static const int per_sm = getenv("FASTDET_C0_CTAS") ? atoi(getenv("FASTDET_C0_CTAS")) : 16;
        const int blocks = static_cast<int>(tiles < 148LL * per_sm ? tiles : 148LL * per_sm);
        conv0_tc_kernel<<<blocks, C0T_THREADS, 0, s>>>(frames, w, bias, out, n, h, wd, cout, out_pitch, act, alpha);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    const size_t smem = (256 + (C0_TH + 2) * (C0_TW + 2) * 3 + 4 + 27 * cout + cout) * sizeof(float) +
                        static_cast<size_t>(C0_TH) * C0_TW * cout * 2;
    dim3 grid((wd + C0_TW - 1) / C0_TW, (h + C0_TH - 1) / C0_TH, n);
    conv0_u8_kernel<<<grid, C0_THREADS, smem, s>>>(frames, w, bias, out, h, wd, cout, out_pitch, act, alpha);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace fd
