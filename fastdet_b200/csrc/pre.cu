// pre.cu — frame ingest kernels (HBM-bound; coalesced / vectorised byte traffic, no tensor cores).
//
//   normalise_f32_nchw : reference server/detector.py:133-134  (u8 HWC -> x/255 in f64 -> f32 -> NCHW)
//   letterbox_u8       : extension (the reference rejects non-net-sized frames, detector.py:131-132)
//   conv0_u8           : the same normalisation fused into the first convolution (Cin = 3), so the f32
//                        NCHW tensor the reference materialises never exists on the serving path.
#include <stdlib.h>

#include "conv_tc.h"
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

// ------------------------------------------------------------------------------------ first conv, im2col by descriptor
// Same layer again, with the im2col done by the tensor core's operand addressing instead of by threads.
//   * The CTA keeps a bf16 halo patch of the input in shared memory, one 16-byte unit per pixel
//     (R, G, B, 0, 0, 0, 0, 0): patch[18 rows][10 pixels] for a tile of 16 rows x 8 columns of output pixels.
//   * In the un-swizzled K-major operand layout a "core matrix" is 8 rows x 16 bytes with the rows 16 bytes apart,
//     the next 8-row group SBO bytes further, the next 16-byte K chunk LBO bytes further.  With the patch above,
//     8 consecutive pixels ARE a core matrix; SBO = one patch row (160 B) walks down the tile's 16 image rows, and
//     LBO = 16 B steps to the next pixel, i.e. the next filter tap of the same filter row.  So the A operand of the
//     MMA for filter row r, taps (s, s+1) is just the patch viewed from pixel (r, s): no per-pixel gather at all
//     (the thread-built version spent ~54 shared-memory loads per output pixel on it).
//   * K per filter row = 4 taps x 8 channels (tap 3 and channels 3..7 meet zero weights): 6 MMAs of 128 x Cout x 16
//     per tile, ~45 cycles each (dev/mma_rate.cu), against the 64 B/pixel the layer has to write to HBM.
//   * Epilogue: each warp drains its TMEM lane quarter (lane = pixel), bias + LeakyReLU in registers, bf16 rows staged
//     in shared memory and written by one TMA store per warp (box = Cout x 8 pixels x 4 rows).
static constexpr int C0D_TW = 8, C0D_TH = 16, C0D_THREADS = 128;
static constexpr int C0D_PW = C0D_TW + 2, C0D_PH = C0D_TH + 2;   // halo patch, pixels
static constexpr int C0D_PATCH_BYTES = (C0D_PH * C0D_PW + 8) * 16;  // + slack: tap 3 of the last pixels reads past the end

__global__ void __launch_bounds__(C0D_THREADS)
conv0_desc_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ wgt, const float* __restrict__ bias,
                  const __grid_constant__ CUtensorMap tm_out, int n, int h, int wd, int cout, int act, float alpha) {
    __shared__ __align__(1024) uint8_t s_patch[2][C0D_PATCH_BYTES];  // double-buffered halo patch
    __shared__ __align__(1024) uint8_t s_w[12 * 32 * 16];            // B: [K chunk 0..11][32 filters][16 B]
    __shared__ __align__(1024) uint8_t s_out[4][2048];               // per warp: 32 pixels x 64 B (SWIZZLE_64B layout)
    __shared__ uint16_t lut[256];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += C0D_THREADS) {
        const __nv_bfloat16 v = __float2bfloat16(norm_u8(i));
        lut[i] = *reinterpret_cast<const uint16_t*>(&v);
    }
    // weights: chunk c = filter row r * 4 + tap s; 8 bf16 = channels (R, G, B, 0...); tap 3 and filters >= cout are zero
    for (int i = tid; i < 12 * 32; i += C0D_THREADS) {
        const int c = i >> 5, f = i & 31, r = c >> 2, sx = c & 3;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f;
        if (sx < 3 && f < cout) {
            const float* wp = wgt + ((r * 3 + sx) * 3) * cout + f;
            w0 = __ldg(wp); w1 = __ldg(wp + cout); w2 = __ldg(wp + 2 * cout);
        }
        *reinterpret_cast<uint4*>(s_w + i * 16) = make_uint4(c0_pack(w0, w1), c0_pack(w2, 0.f), 0u, 0u);
    }
    for (int i = tid; i < 2 * C0D_PATCH_BYTES / 16; i += C0D_THREADS) reinterpret_cast<uint4*>(&s_patch[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        ptx::mbar_init(&mma_bar, 1);
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tm_out);
    }
    if (warp == 0) {
        ptx::tmem_alloc(&tmem_slot, 64);  // two accumulator stages of up to 32 columns
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    ptx::grid_dep_launch();  // the next layer may start its set-up on SMs as they drain (it waits before reading)

    float bv[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bv[c] = (c < cout) ? __ldg(bias + c) : 0.f;
    const float alpha_eff = act ? alpha : 1.0f;
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, cout);
    // un-swizzled K-major descriptors: LBO (K chunk stride) and SBO (8-row group stride), both in 16-byte units
    auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo) -> uint64_t {
        return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
               (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (static_cast<uint64_t>(1) << 46);
    };
    const uint32_t w_addr = ptx::smem_u32(s_w);

    const int tiles_x = (wd + C0D_TW - 1) / C0D_TW, tiles_y = (h + C0D_TH - 1) / C0D_TH;
    const int total = n * tiles_x * tiles_y;  // < 2^31 by far (host checks)
    // raw bytes of the next tile's patch are fetched into registers while the current tile is in the tensor core
    // (the three bytes stay in separate registers until build(): any arithmetic on them here would wait for the loads)
    auto fetch = [&](int tile, uint32_t (&raw)[6], uint32_t& inside) {
        const int f = tile / (tiles_x * tiles_y);
        const int rem = tile - f * tiles_x * tiles_y;
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const uint8_t* fr = frames + 1LL * f * h * wd * 3;
        inside = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = tid + k * C0D_THREADS;  // patch pixel
            if (q < C0D_PH * C0D_PW) {
                const int py = q / C0D_PW, px = q - py * C0D_PW;
                const int gy = ty * C0D_TH - 1 + py, gx = tx * C0D_TW - 1 + px;
                if (gy >= 0 && gy < h && gx >= 0 && gx < wd) {  // outside: zero padding of the normalised input
                    const uint8_t* sp = fr + (1LL * gy * wd + gx) * 3;
                    inside |= 1u << k;
                    raw[3 * k] = __ldg(sp);
                    raw[3 * k + 1] = __ldg(sp + 1);
                    raw[3 * k + 2] = __ldg(sp + 2);
                }
            }
        }
    };
    auto build = [&](int buf, const uint32_t (&raw)[6], uint32_t inside) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = tid + k * C0D_THREADS;
            if (q < C0D_PH * C0D_PW) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (inside & (1u << k)) {
                    v.x = static_cast<uint32_t>(lut[raw[3 * k]]) | (static_cast<uint32_t>(lut[raw[3 * k + 1]]) << 16);
                    v.y = static_cast<uint32_t>(lut[raw[3 * k + 2]]);
                }
                *reinterpret_cast<uint4*>(&s_patch[buf][q * 16]) = v;
            }
        }
    };

    uint32_t raw[6] = {0, 0, 0, 0, 0, 0}, inside = 0;
    uint32_t phase = 0;
    int buf = 0;
    if (blockIdx.x < total) fetch(blockIdx.x, raw, inside);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
        build(buf, raw, inside);
        ptx::fence_proxy_async();  // generic-proxy writes of the patch -> visible to the tensor core
        ptx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            ptx::tc_fence_after();
            const uint32_t p_addr = ptx::smem_u32(&s_patch[buf][0]);
            const uint32_t d = tmem + buf * 32;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int r = j >> 1, s0 = (j & 1) * 2;  // filter row, first of the two taps of this K step
                ptx::umma_bf16(d, desc(p_addr + (r * C0D_PW + s0) * 16, 16, C0D_PW * 16),
                               desc(w_addr + (r * 4 + s0) * 512, 512, 128), idesc, j ? 1u : 0u);
            }
            ptx::umma_commit(&mma_bar);
        }
        const int next = tile + gridDim.x;
        if (next < total) fetch(next, raw, inside);  // global loads in flight while the MMAs run
        const int f = tile / (tiles_x * tiles_y);
        const int rem = tile - f * tiles_x * tiles_y;
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        ptx::mbar_wait(&mma_bar, phase);
        phase ^= 1;
        ptx::tc_fence_after();
        uint32_t acc[32];
        ptx::tmem_ld_32x32(tmem + buf * 32 + (static_cast<uint32_t>(warp * 32) << 16), acc);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        // lane = pixel (warp w: tile rows 4w .. 4w+3, 8 pixels each); bias + LeakyReLU as max(x, alpha x)
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float x0 = __uint_as_float(acc[2 * c]) + bv[2 * c], x1 = __uint_as_float(acc[2 * c + 1]) + bv[2 * c + 1];
            if (act != 2) { x0 = fmaxf(x0, x0 * alpha_eff); x1 = fmaxf(x1, x1 * alpha_eff); }
            else { x0 = x0 > 0.f ? x0 : x0 * alpha; x1 = x1 > 0.f ? x1 : x1 * alpha; }
            pk[c] = c0_pack(x0, x1);
        }
        uint8_t* so = s_out[warp];
        if (lane == 0) ptx::tma_store_wait_read<0>();  // the previous tile's store has finished reading the staging rows
        __syncwarp();
        if (cout == 32) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(so + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        } else {  // cout == 16: 32-byte rows, SWIZZLE_32B (16-byte chunk index ^ bit 2 of the row)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                *reinterpret_cast<uint4*>(so + lane * 32 + ((c ^ ((lane >> 2) & 1)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            ptx::tma_store_4d(&tm_out, so, 0, tx * C0D_TW, ty * C0D_TH + 4 * warp, f);  // clipped at the frame edges
            ptx::tma_store_commit();
        }
        // the next tile's __syncthreads orders these TMEM reads before the accumulator stage is reused two tiles on
    }
    if (lane == 0) ptx::tma_store_wait<0>();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 64);
    }
}

// Warp-specialised form of conv0_desc_kernel (same arithmetic, same operand layouts): the lock-step version spends
// most of a tile's ~5000 cycles waiting at its barriers (ncu: barrier + load-use stalls), so here the three stages run
// decoupled over rings:
//   warps 9..12  builders: warp b owns patch slot b and every 4th tile; it gathers the 180 halo pixels of a tile (6 per
//                lane, the next tile's bytes already in registers), converts through the LUT and publishes the patch
//   warp 8       issues the 6 MMAs of a tile into a ring of 4 TMEM accumulators
//   warps 0..7   two epilogue groups (even / odd tiles; TMEM lane quarter = warp % 4): bias + LeakyReLU, bf16 rows
//                staged in shared memory, one TMA store per warp and tile
static constexpr int C0W_GROUPS = 3;                       // epilogue groups of 4 warps
static constexpr int C0W_EPI_WARPS = 4 * C0W_GROUPS;       // warps 0 .. 11
static constexpr int C0W_MMA_WARP = C0W_EPI_WARPS;         // warp 12
static constexpr int C0W_BUILD_WARPS = 8;                  // one per patch slot
static constexpr int C0W_THREADS = (C0W_EPI_WARPS + 1 + C0W_BUILD_WARPS) * 32, C0W_PATCHES = 8, C0W_ACCS = 8;  // ring depth: the
// publish -> MMA -> drain -> release round trip is ~3000 cycles, so 4 slots capped the kernel at ~800 cycles per tile
// q = x / d for 0 <= x < 2^24 and d < 2^12 as one 64-bit multiply: m = ceil(2^40 / d)
__device__ __forceinline__ int div_magic(int x, unsigned long long m) { return static_cast<int>((static_cast<unsigned long long>(x) * m) >> 40); }
static constexpr int C0W_PPL = (C0D_PH * C0D_PW + 31) / 32;  // patch pixels per builder lane

__global__ void __launch_bounds__(C0W_THREADS)
conv0_ws_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ wgt, const float* __restrict__ bias,
                const __grid_constant__ CUtensorMap tm_out, int n, int h, int wd, int cout, int act, float alpha,
                unsigned long long m_per_frame, unsigned long long m_tiles_x, __nv_bfloat16* __restrict__ out, int out_pitch,
                int direct_store) {
    extern __shared__ __align__(1024) uint8_t c0w_dyn[];        // patch ring
    uint8_t (*s_patch)[C0D_PATCH_BYTES] = reinterpret_cast<uint8_t (*)[C0D_PATCH_BYTES]>(c0w_dyn);
    __shared__ __align__(1024) uint8_t s_w[12 * 32 * 16];       // B: [K chunk 0..11][32 filters][16 B]
    __shared__ __align__(1024) uint8_t s_out[C0W_EPI_WARPS][2048];  // per epilogue warp: one staging tile of 32 pixels x 64 B
    __shared__ uint16_t lut[256];
    __shared__ __align__(8) uint64_t patch_full[C0W_PATCHES], patch_empty[C0W_PATCHES], acc_full[C0W_ACCS], acc_empty[C0W_ACCS];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += C0W_THREADS) {
        const __nv_bfloat16 v = __float2bfloat16(norm_u8(i));
        lut[i] = *reinterpret_cast<const uint16_t*>(&v);
    }
    for (int i = tid; i < 12 * 32; i += C0W_THREADS) {
        const int c = i >> 5, f = i & 31, r = c >> 2, sx = c & 3;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f;
        if (sx < 3 && f < cout) {
            const float* wp = wgt + ((r * 3 + sx) * 3) * cout + f;
            w0 = __ldg(wp); w1 = __ldg(wp + cout); w2 = __ldg(wp + 2 * cout);
        }
        *reinterpret_cast<uint4*>(s_w + i * 16) = make_uint4(c0_pack(w0, w1), c0_pack(w2, 0.f), 0u, 0u);
    }
    for (int i = tid; i < C0W_PATCHES * C0D_PATCH_BYTES / 16; i += C0W_THREADS) reinterpret_cast<uint4*>(c0w_dyn)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int i = 0; i < C0W_PATCHES; ++i) { ptx::mbar_init(&patch_full[i], 1); ptx::mbar_init(&patch_empty[i], 1); }
        for (int i = 0; i < C0W_ACCS; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tm_out);
    }
    if (warp == C0W_MMA_WARP) {
        ptx::tmem_alloc(&tmem_slot, 32 * C0W_ACCS);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async();  // the weight tile and the zeroed patch slack are read by the tensor core
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    ptx::grid_dep_launch();  // the next layer may start its set-up on SMs as they drain (it waits before reading)

    const int tiles_x = (wd + C0D_TW - 1) / C0D_TW, tiles_y = (h + C0D_TH - 1) / C0D_TH;
    const int per_frame = tiles_x * tiles_y;
    const int total = n * per_frame;  // < 2^31 by far
    const int my_tiles = (total > static_cast<int>(blockIdx.x)) ? (total - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

    if (warp > C0W_MMA_WARP) {
        // ---------------------------------------------------------------- builders: u8 halo patch -> bf16, 16 B per pixel
        // warp s of this role owns patch slot s and every 8th tile; a lane handles 6 fixed patch pixels, whose
        // offsets inside a tile are computed once (one warp's dependent-instruction latency per tile is what bounds
        // this role, so the per-tile work is kept to the loads, the LUT and the stores)
        const int slot = warp - C0W_MMA_WARP - 1;
        int pq[C0W_PPL], ppy[C0W_PPL], ppx[C0W_PPL], poff[C0W_PPL];
#pragma unroll
        for (int k = 0; k < C0W_PPL; ++k) {
            const int q = lane + 32 * k;  // patch pixel
            pq[k] = q < C0D_PH * C0D_PW ? q : -1;
            ppy[k] = q / C0D_PW;
            ppx[k] = q - ppy[k] * C0D_PW;
            poff[k] = (ppy[k] * wd + ppx[k]) * 3;
        }
        // two register sets: the bytes of this warp's next TWO tiles are in flight while it converts the current one
        // (the loads come from DRAM: one tile of look-ahead left the warp waiting on them)
        uint32_t raw_a[3 * C0W_PPL], raw_b[3 * C0W_PPL], inside_a = 0, inside_b = 0;
#pragma unroll
        for (int i = 0; i < 3 * C0W_PPL; ++i) raw_a[i] = raw_b[i] = 0;
        auto fetch = [&](int it, uint32_t (&raw)[3 * C0W_PPL], uint32_t& inside) {  // no arithmetic on the bytes here: it would wait for the loads
            const int tile = blockIdx.x + it * gridDim.x;
            const int f = div_magic(tile, m_per_frame);
            const int rem = tile - f * per_frame;
            const int ty = div_magic(rem, m_tiles_x), tx = rem - ty * tiles_x;
            const int y0 = ty * C0D_TH - 1, x0 = tx * C0D_TW - 1;
            const uint8_t* origin = frames + (1LL * f * h * wd + 1LL * y0 * wd + x0) * 3;  // may point before the frame: only used when inside
            inside = 0;
#pragma unroll
            for (int k = 0; k < C0W_PPL; ++k) {
                const int gy = y0 + ppy[k], gx = x0 + ppx[k];
                if (pq[k] >= 0 && gy >= 0 && gy < h && gx >= 0 && gx < wd) {  // outside: zero padding of the normalised input
                    const uint8_t* sp = origin + poff[k];
                    inside |= 1u << k;
                    raw[3 * k] = __ldg(sp);
                    raw[3 * k + 1] = __ldg(sp + 1);
                    raw[3 * k + 2] = __ldg(sp + 2);
                }
            }
        };
        auto publish = [&](const uint32_t (&raw)[3 * C0W_PPL], uint32_t inside, uint32_t phase) {
            ptx::mbar_wait(&patch_empty[slot], phase ^ 1);
#pragma unroll
            for (int k = 0; k < C0W_PPL; ++k) {
                if (pq[k] >= 0) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (inside & (1u << k)) {
                        v.x = static_cast<uint32_t>(lut[raw[3 * k]]) | (static_cast<uint32_t>(lut[raw[3 * k + 1]]) << 16);
                        v.y = static_cast<uint32_t>(lut[raw[3 * k + 2]]);
                    }
                    *reinterpret_cast<uint4*>(&s_patch[slot][pq[k] * 16]) = v;
                }
            }
        };
        auto signal = [&]() {
            ptx::fence_proxy_async();  // generic-proxy writes of the patch -> visible to the tensor core
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&patch_full[slot]);
        };
        uint32_t phase = 0;
        if (slot < my_tiles) fetch(slot, raw_a, inside_a);
        if (slot + C0W_PATCHES < my_tiles) fetch(slot + C0W_PATCHES, raw_b, inside_b);
        for (int it = slot; it < my_tiles; it += 2 * C0W_PATCHES) {
            publish(raw_a, inside_a, phase);
            if (it + 2 * C0W_PATCHES < my_tiles) fetch(it + 2 * C0W_PATCHES, raw_a, inside_a);
            signal();
            phase ^= 1;
            if (it + C0W_PATCHES >= my_tiles) break;
            publish(raw_b, inside_b, phase);
            if (it + 3 * C0W_PATCHES < my_tiles) fetch(it + 3 * C0W_PATCHES, raw_b, inside_b);
            signal();
            phase ^= 1;
        }
    } else if (warp == C0W_MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        // This warp's issue rate bounds the kernel (6 MMAs + 2 commits per 128 pixels), so everything it needs is a
        // warp-uniform value prepared outside the loop and the elected lane only issues (an `if (lane == 0)` body made
        // the compiler wrap every tcgen05.mma in an ELECT loop: ~950 cycles per tile).
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, cout);
        // un-swizzled K-major descriptors: LBO (K chunk stride) and SBO (8-row group stride), in 16-byte units
        auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo) -> uint64_t {
            return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
                   (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (static_cast<uint64_t>(1) << 46);
        };
        const uint32_t w_addr = __shfl_sync(0xffffffffu, ptx::smem_u32(s_w), 0);
        const uint32_t p_addr0 = __shfl_sync(0xffffffffu, ptx::smem_u32(&s_patch[0][0]), 0);
        const uint32_t bar_pe = __shfl_sync(0xffffffffu, ptx::smem_u32(&patch_empty[0]), 0);
        const uint32_t bar_pf = __shfl_sync(0xffffffffu, ptx::smem_u32(&patch_full[0]), 0);
        const uint32_t bar_af = __shfl_sync(0xffffffffu, ptx::smem_u32(&acc_full[0]), 0);
        const uint32_t bar_ae = __shfl_sync(0xffffffffu, ptx::smem_u32(&acc_empty[0]), 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint64_t a0 = desc(p_addr0, 16, C0D_PW * 16);
        uint64_t bd[6], aoff[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int r = j >> 1, s0 = (j & 1) * 2;  // filter row, first of the two taps of this K step
            bd[j] = desc(w_addr + (r * 4 + s0) * 512, 512, 128);
            aoff[j] = static_cast<uint64_t>(r * C0D_PW + s0);  // 16-byte units
        }
        const bool issuer = ptx::elect_one();
        uint32_t slot = 0, phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            ptx::mbar_wait_addr(bar_ae + 8u * slot, phase ^ 1);
            ptx::mbar_wait_addr(bar_pf + 8u * slot, phase);
            ptx::tc_fence_after();
            const uint64_t ad = a0 + slot * (C0D_PATCH_BYTES / 16);
            const uint32_t d = tmem_u + slot * 32;
            if (issuer) {
#pragma unroll
                for (int j = 0; j < 6; ++j) ptx::umma_bf16(d, ad + aoff[j], bd[j], idesc, j ? 1u : 0u);
                ptx::umma_commit_addr(bar_pe + 8u * slot);
                ptx::umma_commit_addr(bar_af + 8u * slot);
            }
            __syncwarp();
            if (++slot == C0W_ACCS) { slot = 0; phase ^= 1; }  // C0W_PATCHES == C0W_ACCS: one index walks both rings
        }
    } else {
        // ---------------------------------------------------------------- epilogue group g: tiles it = g, g + groups, ...
        static_assert(C0W_PATCHES == C0W_ACCS && C0W_ACCS == 8 && C0W_BUILD_WARPS == C0W_PATCHES, "ring indexing below assumes 8 + 8");
        const int group = warp >> 2, quarter = warp & 3;  // TMEM lane quarter = warp % 4
        float bv[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bv[c] = (c < cout) ? __ldg(bias + c) : 0.f;
        const float alpha_eff = act ? alpha : 1.0f;
        for (int it = group; it < my_tiles; it += C0W_GROUPS) {
            const int as = it & 7;
            const uint32_t aphase = (it >> 3) & 1;
            const int tile = blockIdx.x + it * gridDim.x;
            const int f = div_magic(tile, m_per_frame);
            const int rem = tile - f * per_frame;
            const int ty = div_magic(rem, m_tiles_x), tx = rem - ty * tiles_x;
            ptx::mbar_wait(&acc_full[as], aphase);
            ptx::tc_fence_after();
            uint32_t acc[32];
            ptx::tmem_ld_32x32(tmem + as * 32 + (static_cast<uint32_t>(quarter * 32) << 16), acc);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
            // lane = pixel (tile rows 4q .. 4q+3, 8 pixels each); bias + LeakyReLU as max(x, alpha x)
            uint32_t pk[16];
            const float2 a2 = make_float2(alpha_eff, alpha_eff);
#pragma unroll
            for (int c = 0; c < 16; ++c) {  // two columns per FADD2 / FMUL2
                float2 x = __fadd2_rn(make_float2(__uint_as_float(acc[2 * c]), __uint_as_float(acc[2 * c + 1])), make_float2(bv[2 * c], bv[2 * c + 1]));
                if (act != 2) {
                    const float2 m = __fmul2_rn(x, a2);
                    x.x = fmaxf(x.x, m.x);
                    x.y = fmaxf(x.y, m.y);
                } else {
                    x.x = x.x > 0.f ? x.x : x.x * alpha;
                    x.y = x.y > 0.f ? x.y : x.y * alpha;
                }
                pk[c] = c0_pack(x.x, x.y);
            }
            if (direct_store) {
                // straight from the registers: a lane owns one pixel's cout channels (64 contiguous bytes), 8 lanes an image
                // row segment of 512 bytes.  Per tile this costs 4 store instructions of 32 sectors each, against the
                // ~500 cycles a proxy fence + TMA store round costs this small-tile kernel (ncu source page)
                const int oy = ty * C0D_TH + 4 * quarter + (lane >> 3), ox = tx * C0D_TW + (lane & 7);
                if (oy < h && ox < wd) {
                    __nv_bfloat16* op = out + ((static_cast<long long>(f) * h + oy) * wd + ox) * out_pitch;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (8 * c < cout) *reinterpret_cast<uint4*>(op + 8 * c) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                }
                continue;
            }
            uint8_t* so = s_out[warp];
            if (lane == 0) ptx::tma_store_wait_read<0>();  // this warp's previous store (two tiles ago) has finished reading the buffer
            __syncwarp();
            // dense rows [4 image rows][8 pixels][cout]: with pitch == cout the 8 pixels of an image row are one contiguous
            // 8*cout*2-byte run in global memory: the TMA store moves 4 wide rows
            if (cout == 32) {
#pragma unroll
                for (int c = 0; c < 4; ++c)  // (un-swizzled: 4-way bank conflicts on 4 stores per tile, noise)
                    *reinterpret_cast<uint4*>(so + lane * 64 + (c << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    *reinterpret_cast<uint4*>(so + lane * 32 + (c << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&tm_out, so, tx * C0D_TW * cout, ty * C0D_TH + 4 * quarter, f);  // clipped at the frame edges
                ptx::tma_store_commit();
            }
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == C0W_MMA_WARP) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 32 * C0W_ACCS);
    }
}

int kernels_init() {
    return (cudaFuncSetAttribute(conv0_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) == cudaSuccess &&
            cudaFuncSetAttribute(conv0_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C0W_PATCHES * C0D_PATCH_BYTES) == cudaSuccess)
               ? 0
               : -1;
}

int launch_conv0_u8(const uint8_t* frames, const float* w, const float* bias, __nv_bfloat16* out, int n, int h,
                    int wd, int cout, int out_pitch, int act, float alpha, cudaStream_t s) {
    static const int variant = getenv("FASTDET_CONV0") ? atoi(getenv("FASTDET_CONV0")) : 0;  // 0 descriptor im2col, 1 thread-built im2col, 2 CUDA cores
    const int act_mode = act ? ((alpha >= 0.f && alpha <= 1.f) ? 1 : 2) : 0;
    if (variant == 0 && (cout == 16 || cout == 32) && out_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        static const bool lockstep = getenv("FASTDET_C0_LOCKSTEP") != nullptr;
        static const int c0_direct = getenv("FASTDET_C0_DIRECT") ? atoi(getenv("FASTDET_C0_DIRECT")) : 0;  // measured slower (293 vs 255 us)
        CUtensorMap tm;
        if (lockstep || out_pitch != cout) {
            if (!lockstep) return -1;  // (the planner gives the first layer a dense output)
            const unsigned long long dims[4] = {static_cast<unsigned long long>(cout), static_cast<unsigned long long>(wd),
                                                static_cast<unsigned long long>(h), static_cast<unsigned long long>(n)};
            const unsigned long long strides[3] = {2ULL * out_pitch, 2ULL * out_pitch * wd, 2ULL * out_pitch * wd * h};
            const unsigned box[4] = {static_cast<unsigned>(cout), C0D_TW, 4, 1};
            if (encode_tiled_bf16(&tm, out, 4, dims, strides, box, cout == 32 ? 2 : 1)) return -1;
        } else {
            // dense NHWC: an image row is one run of W*C elements; box = 8 pixels x 4 rows
            const unsigned long long dims[3] = {static_cast<unsigned long long>(cout) * wd, static_cast<unsigned long long>(h),
                                                static_cast<unsigned long long>(n)};
            const unsigned long long strides[2] = {2ULL * cout * wd, 2ULL * cout * wd * h};
            const unsigned box[3] = {static_cast<unsigned>(cout) * C0D_TW, 4, 1};
            if (encode_tiled_bf16(&tm, out, 3, dims, strides, box, 0)) return -1;
        }
        const long long tiles = 1LL * n * ((wd + C0D_TW - 1) / C0D_TW) * ((h + C0D_TH - 1) / C0D_TH);
        static const int per_sm = getenv("FASTDET_C0_CTAS") ? atoi(getenv("FASTDET_C0_CTAS")) : 1;
        const int blocks = static_cast<int>(tiles < 148LL * per_sm ? tiles : 148LL * per_sm);
        if (lockstep) conv0_desc_kernel<<<blocks, C0D_THREADS, 0, s>>>(frames, w, bias, tm, n, h, wd, cout, act_mode, alpha);
        else {
            const int tx = (wd + C0D_TW - 1) / C0D_TW, per_frame = tx * ((h + C0D_TH - 1) / C0D_TH);
            if (tiles >= (1LL << 24) || per_frame >= 4096 * 16) return -1;  // range of the multiply-shift division
            const unsigned long long one40 = 1ULL << 40;
            conv0_ws_kernel<<<blocks, C0W_THREADS, C0W_PATCHES * C0D_PATCH_BYTES, s>>>(frames, w, bias, tm, n, h, wd, cout, act_mode, alpha,
                                                          (one40 + per_frame - 1) / per_frame, (one40 + tx - 1) / tx, out, out_pitch,
                                                          c0_direct);
        }
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    if (variant <= 1 && (cout == 16 || cout == 32)) {
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
