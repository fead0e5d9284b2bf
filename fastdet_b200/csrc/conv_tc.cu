// conv_tc.cu — persistent, warp-specialised implicit-GEMM convolution on tcgen05 / TMEM / TMA.
//
//   D[M = N*Ho*Wo, Cout] = im2col(X)[M, K = kh*kw*Cin] * W[Cout, K]^T     (bf16 x bf16 -> fp32)
//   epilogue: + bias (BatchNorm folded) -> LeakyReLU -> (+ residual) -> bf16 NHWC slice | fp32 head rows
//
// Stands in for the Conv/BatchNormalization/LeakyRelu/Add (and the Resize+Concat that follow a
// branch conv) nodes ONNX Runtime executes at reference server/detector.py:135.
//
// CTA = 8 warps:  warp 0 lane 0  TMA producer (A: 2D tiled map for 1x1, im2col map for 3x3; B: 2D tiled)
//                 warp 1 lane 0  tcgen05.mma issuer (128 x BLOCK_N x 16 per instruction)
//                 warp 2         TMEM allocator
//                 warps 4..7     epilogue (TMEM lane quarter = warp % 4)
// Pipelines: smem ring (full/empty mbarriers, STAGES deep) between TMA and MMA; two TMEM accumulator
// stages (tmem_full/tmem_empty) between MMA and epilogue so tile i's epilogue overlaps tile i+1's mainloop.
#include "conv_tc.h"
#include "ptx.cuh"

#include <stdio.h>
#include <string.h>

namespace fd {

static constexpr int BLOCK_M = 128;
static constexpr int NUM_THREADS = 256;
static constexpr int A_STAGE_BYTES = BLOCK_M * 64 * 2;  // sized for block_k = 64

template <int BLOCK_N>
struct TileCfg {
    static constexpr int B_STAGE_BYTES = BLOCK_N * 64 * 2;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;  // 64,128,256,512: powers of two
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;
    static constexpr size_t SMEM_BYTES =
        1024 /*align slack*/ + size_t(STAGES) * (A_STAGE_BYTES + B_STAGE_BYTES) + BIAS_BYTES + BAR_BYTES;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    using Cfg = TileCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    float* sBias = reinterpret_cast<float*>(sB + STAGES * Cfg::B_STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        ptx::tma_prefetch_desc(&tmA);
        ptx::tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);
            ptx::mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t a_bytes = BLOCK_M * p.block_k * 2;
    const uint32_t b_bytes = BLOCK_N * p.block_k * 2;

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.num_n_tiles;
            const int n_tile = tile - m_tile * p.num_n_tiles;
            const int m0 = m_tile * BLOCK_M;
            const int n0 = n_tile * BLOCK_N;
            int img = 0, base_w = 0, base_h = 0;
            if (p.a_im2col) {
                const int hw = p.ho * p.wo;
                img = m0 / hw;
                const int rem = m0 - img * hw;
                const int oy = rem / p.wo;
                const int ox = rem - oy * p.wo;
                base_w = ox * p.stride - p.pad_lo;
                base_h = oy * p.stride - p.pad_lo;
            }
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
                const int tap = kb / p.cin_blocks;
                const int cb = kb - tap * p.cin_blocks;
                if (p.a_im2col) {
                    const int r = tap / p.ksize;
                    const int s = tap - r * p.ksize;
                    ptx::tma_load_im2col_4d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], cb * p.block_k,
                                            base_w, base_h, img, static_cast<uint16_t>(s),
                                            static_cast<uint16_t>(r));
                } else {
                    ptx::tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], cb * p.block_k, m0);
                }
                ptx::tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], kb * p.block_k, n0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = ptx::make_idesc_bf16_f32(BLOCK_M, BLOCK_N);
        // swizzle span = one K block: 128 B (block_k 64), 64 B (32) or 32 B (16)
        const uint32_t layout = (p.block_k == 64) ? 2u : (p.block_k == 32 ? 4u : 6u);
        const uint32_t sbo = 16u * p.block_k;                       // 8 rows x swizzle span
        const int k_steps = p.block_k / 16;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint64_t adesc =
                    ptx::make_kmajor_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES), sbo, layout);
                const uint64_t bdesc =
                    ptx::make_kmajor_desc(ptx::smem_u32(sB + stage * Cfg::B_STAGE_BYTES), sbo, layout);
                for (int k = 0; k < k_steps; ++k) {
                    // advance 16 elements (32 B) along K inside the swizzle span: +2 in 16-byte units
                    ptx::umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                }
                ptx::umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            ptx::umma_commit(&tmem_full_bar[as]);  // accumulator complete
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int quarter = warp & 3;
        const int et = threadIdx.x - 128;  // 0..127 within the epilogue group
        const int row_in_tile = quarter * 32 + lane;
        const int hw = p.ho * p.wo;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m_tile = tile / p.num_n_tiles;
            const int n_tile = tile - m_tile * p.num_n_tiles;
            const int n0 = n_tile * BLOCK_N;
            const long long m = static_cast<long long>(m_tile) * BLOCK_M + row_in_tile;
            const bool row_ok = m < p.M;

            // stage this tile's bias slice (buffer `as`; its previous reader finished two tiles ago,
            // and every epilogue thread passes the named barrier below once per tile)
            float* bias_s = sBias + as * BLOCK_N;
            for (int i = et; i < BLOCK_N; i += 128) bias_s[i] = __ldg(p.bias + n0 + i);
            asm volatile("bar.sync 1, 128;" ::: "memory");

            ptx::mbar_wait(&tmem_full_bar[as], aphase);
            ptx::tc_fence_after();

            long long out_row[4];
            int n_dst = 1;
            if (p.upsample2x) {
                const int img = static_cast<int>(m / hw);
                const int rem = static_cast<int>(m - static_cast<long long>(img) * hw);
                const int oy = rem / p.wo;
                const int ox = rem - oy * p.wo;
                const long long w2 = 2LL * p.wo;
                const long long r0 = (static_cast<long long>(img) * 2 * p.ho + 2 * oy) * w2 + 2 * ox;
                out_row[0] = r0;
                out_row[1] = r0 + 1;
                out_row[2] = r0 + w2;
                out_row[3] = r0 + w2 + 1;
                n_dst = 4;
            } else {
                out_row[0] = m;
            }

#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t acc[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BLOCK_N + c0;
                ptx::tmem_ld_32x32(taddr, acc);
                ptx::tmem_ld_wait();
                if (c0 + 32 >= BLOCK_N) {
                    // last TMEM read of this accumulator stage: hand it back to the MMA warp early
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
                }
                const int nbase = n0 + c0;
                if (!row_ok || nbase >= p.n_store_limit) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(acc[j]) + bias_s[c0 + j];
                    if (p.act) x = x > 0.f ? x : x * p.alpha;
                    v[j] = x;
                }
                if (p.residual != nullptr) {
                    const __nv_bfloat16* rp = p.residual + m * p.res_pitch + nbase;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (nbase + g * 8 < p.cout) {
                            const uint4 r = __ldg(reinterpret_cast<const uint4*>(rp + g * 8));
                            v[g * 8 + 0] += bf16_lo(r.x); v[g * 8 + 1] += bf16_hi(r.x);
                            v[g * 8 + 2] += bf16_lo(r.y); v[g * 8 + 3] += bf16_hi(r.y);
                            v[g * 8 + 4] += bf16_lo(r.z); v[g * 8 + 5] += bf16_hi(r.z);
                            v[g * 8 + 6] += bf16_lo(r.w); v[g * 8 + 7] += bf16_hi(r.w);
                        }
                    }
                }
                if (p.out_fp32) {
                    float* op = reinterpret_cast<float*>(p.out) + out_row[0] * p.out_pitch + nbase;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (nbase + g * 4 < p.n_store_limit)
                            *reinterpret_cast<float4*>(op + g * 4) =
                                make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                    }
                } else {
                    uint4 q[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        q[g].x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                        q[g].y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                        q[g].z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                        q[g].w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                    }
                    for (int d = 0; d < n_dst; ++d) {
                        __nv_bfloat16* op =
                            reinterpret_cast<__nv_bfloat16*>(p.out) + out_row[d] * p.out_pitch + nbase;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (nbase + g * 8 < p.n_store_limit) *reinterpret_cast<uint4*>(op + g * 8) = q[g];
                        }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ host
namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled g_encodeTiled = nullptr;
PFN_encodeIm2col g_encodeIm2col = nullptr;
int g_driver_version = 0;

void set_err(char* err, size_t n, const char* fmt, long long a = 0, long long b = 0, long long c = 0) {
    if (err && n) snprintf(err, n, fmt, a, b, c);
}

template <int BN>
int set_smem_attr() {
    return cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TileCfg<BN>::SMEM_BYTES)) == cudaSuccess
               ? 0
               : -1;
}

}  // namespace

int conv_tc_init(char* err, size_t errlen) {
    if (!g_encodeTiled) {
        cudaDriverEntryPointQueryResult q;
        void* f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
            set_err(err, errlen, "cuTensorMapEncodeTiled entry point unavailable");
            return -1;
        }
        g_encodeTiled = reinterpret_cast<PFN_encodeTiled>(f);
        f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
            set_err(err, errlen, "cuTensorMapEncodeIm2col entry point unavailable");
            return -1;
        }
        g_encodeIm2col = reinterpret_cast<PFN_encodeIm2col>(f);
        cudaDriverGetVersion(&g_driver_version);
    }
    if (set_smem_attr<32>() || set_smem_attr<64>() || set_smem_attr<128>() || set_smem_attr<256>()) {
        set_err(err, errlen, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %lld",
                static_cast<long long>(cudaGetLastError()));
        return -1;
    }
    return 0;
}

static int choose_block_n(int cout, long long m_tiles, int num_sms) {
    (void)m_tiles; (void)num_sms;
    // One 128 x N x 16 MMA reads A (4 KB) + B (N*32 B) from shared memory; at N = 128 that is exactly the
    // 128 B/clk/SM shared-memory bandwidth, so N = 256 is the only shape that keeps the tensor pipe fed
    // (measured: 1.15-1.22 PFLOP/s at N = 256 vs 0.68 at N = 128 on the hot 3x3 layers).
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    if (cout <= 128) return 128;
    return 256;
}

int conv_tc_prepare(const ConvDesc& d, int num_sms, int block_n_hint, ConvLaunch* L, char* err, size_t errlen) {
    if (!g_encodeTiled && conv_tc_init(err, errlen)) return -1;
    memset(L, 0, sizeof(*L));
    const int k = d.ksize;
    if (!(k == 1 || k == 3)) { set_err(err, errlen, "conv_tc: unsupported kernel size %lld", k); return -1; }
    if (d.cin % 16 != 0) { set_err(err, errlen, "conv_tc: Cin=%lld is not a multiple of 16", d.cin); return -1; }
    if (d.in_pitch % 8 != 0 || d.out_pitch % 4 != 0 || (!d.out_fp32 && (d.out_pitch % 8 != 0 || d.cout % 8 != 0))) {
        set_err(err, errlen, "conv_tc: pitches/channels must keep 16-byte alignment (in_pitch=%lld out_pitch=%lld)",
                d.in_pitch, d.out_pitch);
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(d.in) & 15) || (reinterpret_cast<uintptr_t>(d.out) & 15) ||
        (reinterpret_cast<uintptr_t>(d.w) & 15) || (d.residual && (reinterpret_cast<uintptr_t>(d.residual) & 15))) {
        set_err(err, errlen, "conv_tc: pointers must be 16-byte aligned");
        return -1;
    }
    const int ho = (d.hi + d.pad_lo + d.pad_hi - k) / d.stride + 1;
    const int wo = (d.wi + d.pad_lo + d.pad_hi - k) / d.stride + 1;
    const long long M = 1LL * d.n * ho * wo;
    if (M <= 0 || M > 0x7fffffffLL) { set_err(err, errlen, "conv_tc: bad M=%lld", M); return -1; }
    const int block_k = (d.cin % 64 == 0) ? 64 : (d.cin % 32 == 0 ? 32 : 16);
    const int cin_blocks = d.cin / block_k;
    const int K = k * k * d.cin;
    const long long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    int bn = block_n_hint ? block_n_hint : choose_block_n(d.cout, m_tiles, num_sms);
    if (!(bn == 32 || bn == 64 || bn == 128 || bn == 256)) { set_err(err, errlen, "conv_tc: bad block_n %lld", bn); return -1; }

    ConvParams& p = L->p;
    p.M = static_cast<int>(M);
    p.cout = d.cout;
    p.num_k_blocks = k * k * cin_blocks;
    p.cin_blocks = cin_blocks;
    p.ksize = k;
    p.stride = d.stride;
    p.pad_lo = d.pad_lo;
    p.ho = ho;
    p.wo = wo;
    p.block_k = block_k;
    p.a_im2col = !(k == 1 && d.stride == 1 && d.pad_lo == 0 && d.pad_hi == 0);
    p.num_m_tiles = static_cast<int>(m_tiles);
    p.num_n_tiles = (d.cout + bn - 1) / bn;
    p.bias = d.bias;
    p.act = d.act;
    p.alpha = d.alpha;
    p.residual = d.residual;
    p.res_pitch = d.res_pitch;
    p.out = d.out;
    p.out_pitch = d.out_pitch;
    p.out_fp32 = d.out_fp32;
    p.upsample2x = d.upsample2x;
    p.n_store_limit = d.out_fp32 ? d.out_pitch : d.cout;
    if (d.upsample2x && d.out_fp32) { set_err(err, errlen, "conv_tc: upsample2x needs a bf16 output"); return -1; }

    const CUtensorMapSwizzle swz = (block_k == 64)   ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (block_k == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r;
    if (!p.a_im2col) {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.cin), static_cast<cuuint64_t>(M)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.in_pitch) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(block_k), BLOCK_M};
        cuuint32_t estr[2] = {1, 1};
        r = g_encodeTiled(&L->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(d.in), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.cin), static_cast<cuuint64_t>(d.wi),
                              static_cast<cuuint64_t>(d.hi), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.in_pitch) * 2,
                                 static_cast<cuuint64_t>(d.in_pitch) * 2 * d.wi,
                                 static_cast<cuuint64_t>(d.in_pitch) * 2 * d.wi * d.hi};
        int lower[2] = {-d.pad_lo, -d.pad_lo};
        int upper[2] = {d.pad_hi - (k - 1), d.pad_hi - (k - 1)};
        cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(d.stride), static_cast<cuuint32_t>(d.stride), 1};
        r = g_encodeIm2col(&L->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(d.in), dims,
                           strides, lower, upper, static_cast<cuuint32_t>(block_k), BLOCK_M, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        // Same driver quirk CUTLASS works around for im2col maps over small tensors (< 128 KiB).
        const unsigned long long bytes = 2ULL * d.in_pitch * d.wi * d.hi * d.n;
        if (r == CUDA_SUCCESS && g_driver_version <= 13010 && bytes < 131072ULL)
            reinterpret_cast<uint64_t*>(&L->tmA)[1] &= ~(1ULL << 21);
    }
    if (r != CUDA_SUCCESS) { set_err(err, errlen, "conv_tc: tensor map A encode failed (CUresult %lld)", r); return -1; }
    {
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(d.cout)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(block_k), static_cast<cuuint32_t>(bn)};
        cuuint32_t estr[2] = {1, 1};
        r = g_encodeTiled(&L->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(d.w), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_err(err, errlen, "conv_tc: tensor map B encode failed (CUresult %lld)", r); return -1; }
    }
    L->block_n = bn;
    const long long tiles = m_tiles * p.num_n_tiles;
    L->grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
    L->smem_bytes = bn == 32 ? TileCfg<32>::SMEM_BYTES
                  : bn == 64 ? TileCfg<64>::SMEM_BYTES
                  : bn == 128 ? TileCfg<128>::SMEM_BYTES : TileCfg<256>::SMEM_BYTES;
    L->flops = 2.0 * double(M) * d.cout * K;
    return 0;
}

int conv_tc_launch(const ConvLaunch& L, cudaStream_t stream) {
    dim3 grid(L.grid), block(NUM_THREADS);
    switch (L.block_n) {
        case 32: conv_tc_kernel<32><<<grid, block, L.smem_bytes, stream>>>(L.tmA, L.tmB, L.p); break;
        case 64: conv_tc_kernel<64><<<grid, block, L.smem_bytes, stream>>>(L.tmA, L.tmB, L.p); break;
        case 128: conv_tc_kernel<128><<<grid, block, L.smem_bytes, stream>>>(L.tmA, L.tmB, L.p); break;
        case 256: conv_tc_kernel<256><<<grid, block, L.smem_bytes, stream>>>(L.tmA, L.tmB, L.p); break;
        default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace fd
