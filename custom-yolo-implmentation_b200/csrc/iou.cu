// Box / IoU utilities and the stand-alone loss helpers of the reference's public API (sm_100a).
//
//   xywh2xyxy_kernel      src/utils/model_utils.py:153-172
//   bbox_iou_kernel       src/model/losses.py:9-40 (element-wise, b1_y2 slip kept) + backward w.r.t. box1
//   box_iou_kernel        src/utils/model_utils.py:131-151 (pairwise xyxy) and
//                         src/training/metrics.py:6-41     (pairwise xywh, eps 1e-6)
//   qfl_dense_kernel      src/model/losses.py:46-57 on dense (M, C) targets
//   dfl_rows_kernel       src/model/losses.py:63-78
#include "common.cuh"

namespace yb {

__global__ void __launch_bounds__(256) xywh2xyxy_kernel(const float4 *__restrict__ in, size_t n, float4 *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 b = __ldg(in + i);
    const float dw = __fmul_rn(b.z, 0.5f), dh = __fmul_rn(b.w, 0.5f);
    out[i] = make_float4(__fsub_rn(b.x, dw), __fsub_rn(b.y, dh), __fadd_rn(b.x, dw), __fadd_rn(b.y, dh));
}

__global__ void __launch_bounds__(256)
bbox_iou_kernel(const float4 *__restrict__ box1, const float4 *__restrict__ box2, int m, float *__restrict__ out,
                const float *__restrict__ go, float4 *__restrict__ g1) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const float4 a = __ldg(box1 + i), b = __ldg(box2 + i);
    const float ax1 = __fsub_rn(a.x, __fmul_rn(a.z, 0.5f)), ay1 = __fsub_rn(a.y, __fmul_rn(a.w, 0.5f));
    const float ax2 = __fadd_rn(a.x, __fmul_rn(a.z, 0.5f));
    const float ay2 = __fadd_rn(a.w, __fmul_rn(a.y, 0.5f));          // sic: losses.py:20
    const float bx1 = __fsub_rn(b.x, __fmul_rn(b.z, 0.5f)), by1 = __fsub_rn(b.y, __fmul_rn(b.w, 0.5f));
    const float bx2 = __fadd_rn(b.x, __fmul_rn(b.z, 0.5f)), by2 = __fadd_rn(b.y, __fmul_rn(b.w, 0.5f));
    const float iw_raw = __fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1));
    const float ih_raw = __fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1));
    const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float aw = __fsub_rn(ax2, ax1), ah = __fsub_rn(ay2, ay1);
    const float area1 = __fmul_rn(aw, ah);
    const float area2 = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(area1, area2), inter), kEpsIou);
    out[i] = __fdiv_rn(inter, den);
    if (g1 == nullptr) return;
    const float g = __ldg(go + i);
    const float d_inter = g * (1.f / den + inter / (den * den));
    const float d_area1 = -g * inter / (den * den);
    const float d_iw = (iw_raw >= 0.f) ? d_inter * ih : 0.f;
    const float d_ih = (ih_raw >= 0.f) ? d_inter * iw : 0.f;
    auto pick_max = [](float p, float o) { return p > o ? 1.f : (p == o ? 0.5f : 0.f); };
    auto pick_min = [](float p, float o) { return p < o ? 1.f : (p == o ? 0.5f : 0.f); };
    const float d_ax1 = -d_iw * pick_max(ax1, bx1) - d_area1 * ah;
    const float d_ax2 = d_iw * pick_min(ax2, bx2) + d_area1 * ah;
    const float d_ay1 = -d_ih * pick_max(ay1, by1) - d_area1 * aw;
    const float d_ay2 = d_ih * pick_min(ay2, by2) + d_area1 * aw;
    g1[i] = make_float4(d_ax1 + d_ax2, d_ay1 + 0.5f * d_ay2, 0.5f * (d_ax2 - d_ax1), d_ay2 - 0.5f * d_ay1);
}

// pairwise IoU; XYWH selects box_iou_batch's corner conversion, eps is added to the union
template <bool XYWH>
__global__ void __launch_bounds__(256)
box_iou_kernel(const float4 *__restrict__ b1, int n, const float4 *__restrict__ b2, int m, float eps,
               float *__restrict__ out) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= m) return;
    float4 a = __ldg(b1 + i), b = __ldg(b2 + j);
    if (XYWH) {
        a = make_float4(__fsub_rn(a.x, __fmul_rn(a.z, 0.5f)), __fsub_rn(a.y, __fmul_rn(a.w, 0.5f)),
                        __fadd_rn(a.x, __fmul_rn(a.z, 0.5f)), __fadd_rn(a.y, __fmul_rn(a.w, 0.5f)));
        b = make_float4(__fsub_rn(b.x, __fmul_rn(b.z, 0.5f)), __fsub_rn(b.y, __fmul_rn(b.w, 0.5f)),
                        __fadd_rn(b.x, __fmul_rn(b.z, 0.5f)), __fadd_rn(b.y, __fmul_rn(b.w, 0.5f)));
    }
    const float iw = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float area1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    out[(size_t)i * m + j] = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area1, area2), inter), eps));
}

// dense QFL: per-CTA partial sums -> workspace, finished by qfl_finish_kernel in a fixed order
constexpr int kQflThreads = 256;
__global__ void __launch_bounds__(kQflThreads)
qfl_dense_kernel(const float *__restrict__ x, const float *__restrict__ t, size_t n, float inv_m, float beta,
                 float *__restrict__ grad, float *__restrict__ part) {
    __shared__ float s_red[kQflThreads / 32];
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * kQflThreads;
    for (size_t i = (size_t)blockIdx.x * kQflThreads + threadIdx.x; i < n; i += stride) {
        const float xv = __ldg(x + i), tv = __ldg(t + i);
        const float p = __fdiv_rn(1.f, 1.f + expf(-xv));
        const float q = 1.f - p;
        const float lp = logf(p + kEpsLog), lq = logf(q + kEpsLog);
        const float u = 1.f - tv;
        // (1 - p)^beta and p^beta (src/model/losses.py:53-54); beta == 2 is the reference's only use
        const bool sq = beta == 2.f;
        const float qb = sq ? q * q : powf(q, beta), pb = sq ? p * p : powf(p, beta);
        acc += tv * qb * lp + u * pb * lq;
        if (grad) {
            const float qb1 = sq ? q : powf(q, beta - 1.f), pb1 = sq ? p : powf(p, beta - 1.f);
            const float dpos = tv * (-beta * qb1 * lp + qb / (p + kEpsLog));
            const float dneg = u * (beta * pb1 * lq - pb / (q + kEpsLog));
            grad[i] = -inv_m * (dpos + dneg) * p * q;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kQflThreads / 32; ++w) s += s_red[w];
        part[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) qfl_finish_kernel(const float *__restrict__ part, int n_part, float inv_m, float *out) {
    __shared__ double s[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < n_part; i += 256) a += (double)part[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(-s[0] * (double)inv_m);
}

// distribution_focal_loss: one thread per row (rows are short), single CTA, fixed-order reduction
__global__ void __launch_bounds__(256)
dfl_rows_kernel(const float *__restrict__ z, const float *__restrict__ target, int m, int r, float *__restrict__ out,
                float *__restrict__ grad) {
    __shared__ double s[256];
    double acc = 0.0;
    const float inv_m = 1.f / (float)m;
    for (int i = threadIdx.x; i < m; i += 256) {
        const float *row = z + (size_t)i * r;
        float mx = row[0];
        for (int j = 1; j < r; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
        for (int j = 0; j < r; ++j) sum += expf(row[j] - mx);
        const float lse = mx + logf(sum);
        const float t = target[i];
        const int bl = (int)t;
        const int br = min(bl + 1, r - 1);
        const float wl = (float)(bl + 1) - t, wr = t - (float)bl;
        acc += (double)((lse - row[bl]) * wl + (lse - row[br]) * wr);
        if (grad) {
            for (int j = 0; j < r; ++j) {
                const float p = expf(row[j] - lse);
                grad[(size_t)i * r + j] = inv_m * ((wl + wr) * p - (j == bl ? wl : 0.f) - (j == bl + 1 ? wr : 0.f));
            }
        }
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(s[0] / (double)m);
}

static int qfl_blocks(size_t n) {
    const size_t want = (n + kQflThreads - 1) / kQflThreads;
    return (int)(want < (size_t)(148 * 8) ? (want ? want : 1) : (size_t)(148 * 8));
}

}  // namespace yb

using namespace yb;

extern "C" int yb_xywh2xyxy(const float *in, size_t n_boxes, float *out, void *stream) {
    YB_REQUIRE(in && out, "yb_xywh2xyxy: null pointer");
    if (n_boxes == 0) return YB_OK;
    if (!aligned16(in) || !aligned16(out)) {
        set_error("yb_xywh2xyxy: (n, 4) fp32 buffers must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    xywh2xyxy_kernel<<<(unsigned)((n_boxes + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (const float4 *)in, n_boxes, (float4 *)out);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

extern "C" int yb_bbox_iou(const float *box1, const float *box2, int m, float *out_iou, const float *grad_out,
                           float *grad_box1, void *stream) {
    YB_REQUIRE(box1 && box2 && out_iou, "yb_bbox_iou: null pointer");
    YB_REQUIRE(m >= 0, "yb_bbox_iou: negative size");
    YB_REQUIRE((grad_box1 == nullptr) || (grad_out != nullptr), "yb_bbox_iou: grad_box1 needs grad_out");
    if (m == 0) return YB_OK;
    if (!aligned16(box1) || !aligned16(box2) || (grad_box1 && !aligned16(grad_box1))) {
        set_error("yb_bbox_iou: (M, 4) fp32 buffers must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    bbox_iou_kernel<<<(m + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (const float4 *)box1, (const float4 *)box2, m, out_iou, grad_out, (float4 *)grad_box1);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

static int pairwise(const float *box1, int n, const float *box2, int m, float eps, float *out, bool xywh, void *stream) {
    YB_REQUIRE(n >= 0 && m >= 0 && n <= 65535, "pairwise IoU: bad sizes (n <= 65535)");
    if (n == 0 || m == 0) return YB_OK;
    YB_REQUIRE(box1 && box2 && out, "pairwise IoU: null pointer");
    if (!aligned16(box1) || !aligned16(box2)) {
        set_error("pairwise IoU: (n, 4) fp32 buffers must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    dim3 grid((m + 255) / 256, n);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (xywh)
        box_iou_kernel<true><<<grid, 256, 0, st>>>((const float4 *)box1, n, (const float4 *)box2, m, eps, out);
    else
        box_iou_kernel<false><<<grid, 256, 0, st>>>((const float4 *)box1, n, (const float4 *)box2, m, eps, out);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

extern "C" int yb_box_iou(const float *box1, int n, const float *box2, int m, float eps, float *out, void *stream) {
    return pairwise(box1, n, box2, m, eps, out, false, stream);
}

extern "C" int yb_box_iou_batch(const float *box1, int n, const float *box2, int m, float *out, void *stream) {
    return pairwise(box1, n, box2, m, 1e-6f, out, true, stream);
}

extern "C" size_t yb_qfl_workspace_bytes(size_t n_elements) { return sizeof(float) * (size_t)qfl_blocks(n_elements); }

extern "C" int yb_quality_focal_loss(const float *pred_scores, const float *target_scores, int m, int c, float beta,
                                     float *out_loss, float *grad_scores, void *workspace, size_t workspace_bytes,
                                     void *stream) {
    YB_REQUIRE(pred_scores && target_scores && out_loss && workspace, "yb_quality_focal_loss: null pointer");
    YB_REQUIRE(m > 0 && c > 0, "yb_quality_focal_loss: bad sizes");
    YB_REQUIRE(beta > 0.f, "yb_quality_focal_loss: beta must be positive");
    const size_t n = (size_t)m * c;
    const int blocks = qfl_blocks(n);
    if (workspace_bytes < sizeof(float) * (size_t)blocks) {
        set_error("yb_quality_focal_loss: workspace too small");
        return YB_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv_m = 1.f / (float)m;
    qfl_dense_kernel<<<blocks, kQflThreads, 0, st>>>(pred_scores, target_scores, n, inv_m, beta, grad_scores,
                                                     (float *)workspace);
    YB_LAUNCH_CHECK();
    qfl_finish_kernel<<<1, 256, 0, st>>>((const float *)workspace, blocks, inv_m, out_loss);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

extern "C" int yb_distribution_focal_loss(const float *pred_dist, const float *target, int m, int r, float *out_loss,
                                          float *grad_dist, void *stream) {
    YB_REQUIRE(pred_dist && target && out_loss, "yb_distribution_focal_loss: null pointer");
    YB_REQUIRE(m > 0 && r > 1, "yb_distribution_focal_loss: bad sizes");
    dfl_rows_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred_dist, target, m, r, out_loss, grad_dist);
    YB_LAUNCH_CHECK();
    return YB_OK;
}
