// Validation bookkeeping (sm_100a): greedy, class-matched TP / FP / FN counting of decoded
// predictions against the ground truth.  Replaces DetectionMetrics.update
// (src/training/metrics.py:68-160), a pure-Python O(P*G) double loop with one .item() per element,
// called once per image at src/training/train_model.py:326-328.  Here: one warp per image for the
// whole batch, counters accumulated on the device with integer atomics (exact, order-independent).
//
// Per image, predictions are visited in row order; prediction i takes the still-unmatched target of
// its class with the largest IoU (strictly above the best so far, so the lowest index wins ties; the
// IoU must be > 0); it is a true positive when that IoU >= iou_threshold and the target is then
// consumed.  IoU as box_iou_batch (metrics.py:6-41): xywh boxes, 1e-6 added to the union.
// The reference's early returns are kept: with no prediction (or no target) only the FN (FP) side is
// counted and total_predictions / total_ground_truths are NOT advanced (metrics.py:87-104).
#include "common.cuh"

namespace yb {

constexpr int kMaxTargetsPerLane = 32;      // a lane's "consumed" bits: up to 1024 targets per image

// counters layout (int64): [0] tp [1] fp [2] fn [3] total_predictions [4] total_ground_truths
//                          [8 + 0*nc ..] class_tp  [8 + 1*nc ..] class_fp  [8 + 2*nc ..] class_fn  [8 + 3*nc ..] class_gt
__device__ __forceinline__ void bump(unsigned long long *c, int nc, int which, int cls) {
    if (cls >= 0 && cls < nc) atomicAdd(c + 8 + (size_t)which * nc + cls, 1ull);
}

__device__ __forceinline__ float iou_xywh(const float4 &a, const float4 &b) {
    const float ax1 = __fsub_rn(a.x, __fmul_rn(a.z, 0.5f)), ay1 = __fsub_rn(a.y, __fmul_rn(a.w, 0.5f));
    const float ax2 = __fadd_rn(a.x, __fmul_rn(a.z, 0.5f)), ay2 = __fadd_rn(a.y, __fmul_rn(a.w, 0.5f));
    const float bx1 = __fsub_rn(b.x, __fmul_rn(b.z, 0.5f)), by1 = __fsub_rn(b.y, __fmul_rn(b.w, 0.5f));
    const float bx2 = __fadd_rn(b.x, __fmul_rn(b.z, 0.5f)), by2 = __fadd_rn(b.y, __fmul_rn(b.w, 0.5f));
    const float iw = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float a1 = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
    const float a2 = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-6f));
}

__global__ void __launch_bounds__(128)
detection_match_kernel(const float *__restrict__ pred_rows, int row_stride, const int *__restrict__ pred_count,
                       const float *__restrict__ pred_scores, float score_thr, const float *__restrict__ gt,
                       const int *__restrict__ gt_off, int n_images, int nc, float iou_thr, int skip_empty_targets,
                       unsigned long long *__restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (n >= n_images) return;
    const float *rows = pred_rows + (size_t)n * row_stride * 5;
    const float *sc = pred_scores ? pred_scores + (size_t)n * row_stride : nullptr;
    const int p_all = pred_count[n];
    const int g0 = gt_off[n], m = gt_off[n + 1] - g0;
    // the reference's validation loop only calls update() for images that have targets (train_model.py:326-328)
    if (m == 0 && skip_empty_targets) return;

    // predictions that pass the score filter (metrics.py:82-86)
    int p = 0;
    for (int i = lane; i < p_all; i += 32) p += (!sc || sc[i] >= score_thr) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (p == 0 && m == 0) return;
    if (p == 0) {                                          // :89-96
        if (lane == 0) atomicAdd(counters + 2, (unsigned long long)m);
        for (int j = lane; j < m; j += 32) {
            const int c = (int)gt[(size_t)(g0 + j) * 5 + 4];
            bump(counters, nc, 2, c);
            bump(counters, nc, 3, c);
        }
        return;
    }
    if (m == 0) {                                          // :98-104
        if (lane == 0) atomicAdd(counters + 1, (unsigned long long)p);
        for (int i = lane; i < p_all; i += 32)
            if (!sc || sc[i] >= score_thr) bump(counters, nc, 1, (int)rows[(size_t)i * 5 + 4]);
        return;
    }
    unsigned used = 0;                                     // bit t: my t-th strided target is consumed
    int tp = 0, fp = 0;
    // the lane's first two targets stay in registers (images with up to 64 targets never re-read them)
    float4 gbox[2];
    int gcls[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = lane + 32 * t;
        gbox[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        gcls[t] = -0x7fffffff;
        if (j < m) {
            const float *g5 = gt + (size_t)(g0 + j) * 5;
            gbox[t] = make_float4(g5[0], g5[1], g5[2], g5[3]);
            gcls[t] = (int)g5[4];
        }
    }
    for (int i0 = 0; i0 < p_all; i0 += 32) {               // 32 predictions at a time: lane l fetches row i0 + l
        float4 my_pb = make_float4(0.f, 0.f, 0.f, 0.f);
        int my_pc = 0;
        bool my_ok = false;
        if (i0 + lane < p_all) {
            const float *r5 = rows + (size_t)(i0 + lane) * 5;
            my_pb = make_float4(r5[0], r5[1], r5[2], r5[3]);
            my_pc = (int)r5[4];
            my_ok = !sc || sc[i0 + lane] >= score_thr;
        }
        const unsigned ok_mask = __ballot_sync(0xffffffffu, my_ok);
        const int n_here = min(32, p_all - i0);
        for (int k = 0; k < n_here; ++k) {                 // prediction order matters: sequential
            if (!((ok_mask >> k) & 1u)) continue;          // warp-uniform
            const float4 pb = make_float4(__shfl_sync(0xffffffffu, my_pb.x, k), __shfl_sync(0xffffffffu, my_pb.y, k),
                                          __shfl_sync(0xffffffffu, my_pb.z, k), __shfl_sync(0xffffffffu, my_pb.w, k));
            const int pc = __shfl_sync(0xffffffffu, my_pc, k);
            float best = 0.f;
            int bj = 0x7fffffff;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (!((used >> t) & 1u) && gcls[t] == pc) {
                    const float v = iou_xywh(pb, gbox[t]);
                    if (v > best) { best = v; bj = lane + 32 * t; }           // strict: first maximum within the lane
                }
            }
            for (int j = lane + 64, t = 2; j < m; j += 32, ++t) {
                if ((used >> t) & 1u) continue;
                const float *g5 = gt + (size_t)(g0 + j) * 5;
                if ((int)g5[4] != pc) continue;
                const float v = iou_xywh(pb, make_float4(g5[0], g5[1], g5[2], g5[3]));
                if (v > best) { best = v; bj = j; }
            }
            // IoUs are >= 0, so their bit patterns order like integers; ties go to the lowest target index
            const int wbi = __reduce_max_sync(0xffffffffu, __float_as_int(best));
            const int wj = __reduce_min_sync(0xffffffffu, __float_as_int(best) == wbi ? bj : 0x7fffffff);
            const float wb = __int_as_float(wbi);
            if (wj != 0x7fffffff && wb >= iou_thr) {       // :136-141 (an IoU of 0 never selects a target)
                ++tp;
                if ((wj & 31) == lane) used |= 1u << (wj >> 5);
                if (lane == 0) bump(counters, nc, 0, pc);
            } else {
                ++fp;
                if (lane == 0) bump(counters, nc, 1, pc);
            }
        }
    }
    int fn = 0;
    for (int j = lane, t = 0; j < m; j += 32, ++t) {       // :146-155
        const int c = (int)gt[(size_t)(g0 + j) * 5 + 4];
        bump(counters, nc, 3, c);
        if (!((used >> t) & 1u)) {
            ++fn;
            bump(counters, nc, 2, c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fn += __shfl_xor_sync(0xffffffffu, fn, o);
    if (lane == 0) {
        atomicAdd(counters + 0, (unsigned long long)tp);
        atomicAdd(counters + 1, (unsigned long long)fp);
        atomicAdd(counters + 2, (unsigned long long)fn);
        atomicAdd(counters + 3, (unsigned long long)p);
        atomicAdd(counters + 4, (unsigned long long)m);
    }
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_detection_counters_bytes(int nc) { return nc > 0 ? sizeof(unsigned long long) * (8 + 4 * (size_t)nc) : 0; }

extern "C" int yb_detection_match(const float *pred_rows, int row_stride, const int32_t *pred_count,
                                  const float *pred_scores, float score_threshold, const float *gt,
                                  const int32_t *gt_offsets, int gmax, int n_images, int nc, float iou_threshold,
                                  int skip_empty_targets, uint64_t *counters, void *stream) {
    YB_NVTX("yb_detection_match");
    YB_REQUIRE(pred_count && gt_offsets && counters, "yb_detection_match: null pointer");
    YB_REQUIRE(n_images > 0 && nc > 0 && row_stride >= 0, "yb_detection_match: bad sizes");
    YB_REQUIRE(row_stride == 0 || pred_rows, "yb_detection_match: pred_rows is null");
    YB_REQUIRE(gmax <= 32 * kMaxTargetsPerLane, "yb_detection_match: at most %d targets per image", 32 * kMaxTargetsPerLane);
    YB_REQUIRE(gmax == 0 || gt, "yb_detection_match: gt is null");
    detection_match_kernel<<<(n_images + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        pred_rows, row_stride, pred_count, pred_scores, score_threshold, gt, gt_offsets, n_images, nc, iou_threshold,
        skip_empty_targets, reinterpret_cast<unsigned long long *>(counters));
    YB_LAUNCH_CHECK();
    return YB_OK;
}
