// post.cu — YOLO head decode + threshold + compaction, and per-frame Gaussian Soft-NMS.
//
// Replaces the two pure-Python hot loops of the reference:
//   decode_kernel    : ONNXDetector.process_yolo, server/detector.py:148-166 (+ sigmoid :12-13)
//   soft_nms_kernel  : soft_nms :45-59 with YOLOObject.get_iou :38-42 / rect_intersect :15-22, and the
//                      pixel scaling of the result tuples :142-144
// All arithmetic is float64 in the reference (Python floats); it is float64 here too, in the same
// operation order, so decisions at the threshold agree except where exp() differs in its last bit.
//
// decode: one warp per grid cell.  Lanes 0..2 test the three anchors' objectness (1 sector of the cell's
// row); only if one passes does the warp read the class logits (coalesced) for a first-maximum arg-max.
// Candidates are appended with an atomic per-frame counter and carry their insertion-order index `box`
// (head, row, column, anchor) — the order the reference's list/dict iteration has — so later tie-breaks
// do not depend on the append order.
#include <limits.h>
#include <math.h>

#include "kernels.h"

namespace fd {

struct DecodeParams {
    HeadDesc heads[4];
    int cell_start[5];  // prefix sum of h*w per head
    int n_heads, num_classes, n, net_w, net_h, boxes_per_frame;
    double threshold;
    float obj_cut;  // logits below this cannot reach the threshold (logit(threshold) minus a safety margin); -inf: no pre-test
};

__device__ __forceinline__ double logistic(float v) { return 1.0 / (1.0 + exp(-static_cast<double>(v))); }

__global__ void __launch_bounds__(256)
decode_kernel(const DecodeParams p, Candidate* __restrict__ cand, int* __restrict__ cand_count) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (blockIdx.x * 256LL + threadIdx.x) >> 5;
    const int cells_per_frame = p.cell_start[p.n_heads];
    if (warp_global >= 1LL * p.n * cells_per_frame) return;
    const int f = static_cast<int>(warp_global / cells_per_frame);
    const int cell_all = static_cast<int>(warp_global - 1LL * f * cells_per_frame);
    int hi = 0;
    while (hi + 1 < p.n_heads && cell_all >= p.cell_start[hi + 1]) ++hi;
    const HeadDesc& H = p.heads[hi];
    const int cell = cell_all - p.cell_start[hi];
    const int gy = cell / H.w, gx = cell - gy * H.w;
    const float* row = H.data + (1LL * f * H.h * H.w + cell) * H.pitch;
    const int span = 5 + p.num_classes;

    double obj = 0.0;
    bool pass = false;
    if (lane < 3) {
        // float pre-test (sigmoid is monotonic, the margin dwarfs any rounding): ~99 % of the boxes end here and never
        // pay for the float64 exp below; the decision itself is taken in float64 exactly as the reference does
        const float t4 = __ldg(row + lane * span + 4);
        if (!(t4 < p.obj_cut)) {
            obj = logistic(t4);
            pass = !(obj < p.threshold);  // reference: `if conf < threshold: continue`
        }
    }
    unsigned mask = __ballot_sync(0xffffffffu, pass);
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const float* cls = row + k * span + 5;
        // first-maximum arg-max over the raw logits (np.argmax)
        float best = -INFINITY;
        int best_i = INT_MAX;
        for (int c = lane; c < p.num_classes; c += 32) {
            const float v = __ldg(cls + c);
            if (v > best || best_i == INT_MAX) { best = v; best_i = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (oi != INT_MAX && (best_i == INT_MAX || ov > best || (ov == best && oi < best_i))) { best = ov; best_i = oi; }
        }
        const double objk = __shfl_sync(0xffffffffu, obj, k);
        if (lane == 0) {
            const double conf = objk * logistic(best);
            if (!(conf < p.threshold)) {
                const float* b = row + k * span;
                const double x = (gx + logistic(__ldg(b + 0))) / H.w;
                const double y = (gy + logistic(__ldg(b + 1))) / H.h;
                const double w = static_cast<double>(H.anchor_w[k]) * exp(static_cast<double>(__ldg(b + 2))) / p.net_w;
                const double h = static_cast<double>(H.anchor_h[k]) * exp(static_cast<double>(__ldg(b + 3))) / p.net_h;
                const int slot = atomicAdd(cand_count + f, 1);
                Candidate c;
                c.conf = conf;
                c.x = x - w / 2;
                c.y = y - h / 2;
                c.w = w;
                c.h = h;
                c.box = H.first_box + cell * 3 + k;
                c.klass = best_i + 1;
                cand[1LL * f * p.boxes_per_frame + slot] = c;
            }
        }
    }
}

int launch_decode(const HeadDesc* heads, int n_heads, int num_classes, int n, int net_w, int net_h,
                  double threshold, Candidate* cand, int* cand_count, int boxes_per_frame, cudaStream_t s) {
    if (n_heads < 1 || n_heads > 4) return -1;
    DecodeParams p;
    p.cell_start[0] = 0;
    for (int i = 0; i < n_heads; ++i) {
        p.heads[i] = heads[i];
        p.cell_start[i + 1] = p.cell_start[i] + heads[i].h * heads[i].w;
    }
    p.n_heads = n_heads; p.num_classes = num_classes; p.n = n; p.net_w = net_w; p.net_h = net_h;
    p.boxes_per_frame = boxes_per_frame; p.threshold = threshold;
    p.obj_cut = (threshold > 0.0 && threshold < 1.0) ? static_cast<float>(log(threshold / (1.0 - threshold)) - 1e-3) : -INFINITY;
    if (cudaMemsetAsync(cand_count, 0, sizeof(int) * n, s) != cudaSuccess) return -1;
    const long long warps = 1LL * n * p.cell_start[n_heads];
    const int blocks = static_cast<int>((warps + 7) / 8);
    decode_kernel<<<blocks, 256, 0, s>>>(p, cand, cand_count);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------ Soft-NMS
static constexpr int NMS_THREADS = 256;

__global__ void __launch_bounds__(NMS_THREADS)
soft_nms_kernel(const Candidate* __restrict__ cand_all, const int* __restrict__ cand_count, double* __restrict__ score_all,
                int cap, int net_w, int net_h, double threshold, Detection* __restrict__ out, int* __restrict__ out_count,
                int* __restrict__ total_count, int max_det) {
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = min(cand_count[f], cap);
    const Candidate* cand = cand_all + 1LL * f * cap;
    double* sc = score_all + 1LL * f * cap;
    __shared__ double s_score[NMS_THREADS / 32];
    __shared__ int s_box[NMS_THREADS / 32], s_pos[NMS_THREADS / 32];
    __shared__ double sel_box[4];
    __shared__ int sel_pos;

    for (int i = tid; i < C; i += NMS_THREADS) sc[i] = cand[i].conf;
    __syncthreads();
    int kept = 0;
    while (true) {
        // arg-max of the live scores; ties go to the lowest insertion index (strict '<' at detector.py:51)
        double best = -1.0;
        int best_box = INT_MAX, best_pos = -1;
        for (int i = tid; i < C; i += NMS_THREADS) {
            const double s = sc[i];
            if (s < 0.0) continue;  // already selected
            const int b = cand[i].box;
            if (s > best || (s == best && b < best_box)) { best = s; best_box = b; best_pos = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, best, o);
            const int ob = __shfl_xor_sync(0xffffffffu, best_box, o);
            const int op = __shfl_xor_sync(0xffffffffu, best_pos, o);
            if (os > best || (os == best && ob < best_box)) { best = os; best_box = ob; best_pos = op; }
        }
        if (lane == 0) { s_score[warp] = best; s_box[warp] = best_box; s_pos[warp] = best_pos; }
        __syncthreads();
        if (tid == 0) {
            for (int w2 = 1; w2 < NMS_THREADS / 32; ++w2)
                if (s_score[w2] > best || (s_score[w2] == best && s_box[w2] < best_box)) {
                    best = s_score[w2]; best_box = s_box[w2]; best_pos = s_pos[w2];
                }
            if (best_pos < 0 || best < threshold) {  // `if mconf < threshold: break`
                sel_pos = -1;
            } else {
                sel_pos = best_pos;
                const Candidate c = cand[best_pos];
                sel_box[0] = c.x; sel_box[1] = c.y; sel_box[2] = c.w; sel_box[3] = c.h;
                if (kept < max_det) {
                    Detection d;
                    d.klass = c.klass;
                    d.box = c.box;
                    d.conf = c.conf;  // the reference reports the original, undecayed score (detector.py:142)
                    d.x = c.x * net_w; d.y = c.y * net_h; d.w = c.w * net_w; d.h = c.h * net_h;
                    out[1LL * f * max_det + kept] = d;
                }
                sc[best_pos] = -2.0;
            }
        }
        __syncthreads();
        if (sel_pos < 0) break;
        ++kept;
        const double x0 = sel_box[0], y0 = sel_box[1], w0 = sel_box[2], h0 = sel_box[3];
        for (int i = tid; i < C; i += NMS_THREADS) {
            const double s = sc[i];
            if (s < 0.0) continue;
            const Candidate c = cand[i];
            const double iw = fmin(x0 + w0, c.x + c.w) - fmax(x0, c.x);
            const double ih = fmin(y0 + h0, c.y + c.h) - fmax(y0, c.y);
            if (iw <= 0.0 || ih <= 0.0) continue;  // overlap 0 -> factor exp(0) = 1
            const double ov = (iw * ih) / (w0 * h0);  // area(sel ∩ other) / area(sel): asymmetric, as the reference
            sc[i] = s * exp(-3.0 * (ov * ov));
        }
        __syncthreads();
    }
    if (tid == 0) {
        out_count[f] = kept < max_det ? kept : max_det;
        total_count[f] = kept;
    }
}

int launch_soft_nms(Candidate* cand, const int* cand_count, double* score_scratch, int boxes_per_frame, int n,
                    int net_w, int net_h, double threshold, Detection* out, int* out_count, int* total_count,
                    int max_det, cudaStream_t s) {
    soft_nms_kernel<<<n, NMS_THREADS, 0, s>>>(cand, cand_count, score_scratch, boxes_per_frame, net_w, net_h, threshold,
                                              out, out_count, total_count, max_det);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace fd
