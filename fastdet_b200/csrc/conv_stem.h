// conv_stem.h — host interface of the fused network stem (see conv_stem.cu): /255 normalisation + first convolution
// (3 -> 32, 3x3 stride 1) + second convolution (32 -> 64, 3x3 stride 2) in one kernel; the first convolution's activation
// (709 MB per 64 frames of 416x416) never reaches HBM.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fd {

struct StemDesc {
    int n, h, w;              // u8 frames [n][h][w][3]
    const uint8_t* frames;
    // first convolution: 3x3, stride 1, pad 1, fp32 weights [3][3][3][c1] (BatchNorm folded), LeakyReLU(alpha1) if act1
    int c1;
    const float* w1;          // device
    const float* bias1_host;
    int act1;
    float alpha1;
    // second convolution: 3x3, stride 2, pad (1, pad_hi2), bf16 weights [c2][3*3*c1]
    int c2, pad_hi2;
    const __nv_bfloat16* w2;  // device
    const float* bias2_host;
    int act2;
    float alpha2;
    __nv_bfloat16* out;       // bf16 NHWC slice of the (ho, wo) map
    int out_pitch;
};

struct StemParams {
    const uint8_t* frames;
    int n, h, w, ho, wo;
    const float* w1;
    const __nv_bfloat16* w2;
    float alpha1, alpha2;     // effective slopes (1 = linear)
    int tiles_x, per_frame, total;
    unsigned long long m_per_frame, m_tiles_x;  // ceil(2^40 / d): x / d == (x * m) >> 40 for the ranges checked on the host
    float bias1[32], bias2[64];
    int debug;  // developer switches, honoured by the harness build only (0 in production)
    long long* prof;  // developer: per-CTA cycle counters [grid][16] (null in production)
};

struct StemLaunch {
    CUtensorMap tm_out;  // {C, W, H, N} of the output slice, box 32 channels x 8 pixels x 4 rows, SWIZZLE_64B
    StemParams p;
    int grid;
    size_t smem_bytes;
    double flops;        // algorithmic, both convolutions
};

// One-time per device: opt in to the large dynamic shared memory the kernel needs.
int conv_stem_init();
bool conv_stem_supported(const StemDesc& d);
int conv_stem_prepare(const StemDesc& d, int num_sms, StemLaunch* out, char* err, size_t errlen);
int conv_stem_launch(const StemLaunch& L, cudaStream_t stream);

}  // namespace fd
