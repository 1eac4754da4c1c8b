// Fused decode + nearest-centre assignment + DFL/QFL loss + backward (sm_100a).
//
// Replaces YoloDFLQFLoss.forward and its autograd backward (src/model/losses.py:93-281).
// Three launches per step, every head-output byte read once and every gradient byte written once;
// the (M x A) distance matrix, the decoded boxes and the dense (A x nc) QFL target never exist:
//
//   assign_kernel    reads the 4*16 box channels (128-bit loads), decodes each anchor's predicted
//                    centre, scans it against the image's GT centres (one GT per thread, anchors
//                    streamed from shared memory), merges the per-CTA winners with a 64-bit
//                    atomicMax on (distance, anchor) keys, and zero-fills the box-channel gradient.
//   match_kernel     one warp per GT: gathers the 64 logits of the matched anchor, DFL loss and its
//                    gradient, IoU soft target (reference formula, slip included) and the gradient
//                    that flows through it, duplicate-anchor resolution; writes the match table.
//   cls_loss_kernel  reads the nc class channels, QFL loss + gradient in one pass with the target
//                    looked up from a per-tile table in shared memory; the last CTA reduces the
//                    per-CTA partial sums in a fixed order and writes the loss scalars.
#include "common.cuh"

namespace yb {

constexpr int kAssignThreads = 128;
constexpr int kClsThreads = 128;
constexpr unsigned long long kNoKey = ~0ull;

struct LossWorkspace {
    unsigned int *ticket;          // [1]   cls_loss_kernel completion counter      } zeroed
    unsigned long long *best;      // [gt_total] inverted (distance, anchor) keys   } every call
    int *m_idx;                    // [gt_total] matched anchor
    int *m_cls;                    // [gt_total] class id, or -1 when this GT does not own its anchor's target row
    float *m_iou;                  // [gt_total]
    float *m_dfl;                  // [gt_total] sum over the 4 sides of the DFL term
    float *part;                   // [N * tiles] per-CTA sums of p^2 log(1-p) etc.
    size_t zero_bytes;
    size_t total_bytes;
};

static LossWorkspace carve(void *base, int n_images, int cls_tiles, int gt_total) {
    LossWorkspace w;
    char *p = static_cast<char *>(base);
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned int *>(p + off);
    off += 64;
    w.best = reinterpret_cast<unsigned long long *>(p + off);
    off += round_up(sizeof(unsigned long long) * (size_t)(gt_total > 0 ? gt_total : 1), 64);
    w.zero_bytes = off;
    const size_t g4 = round_up(sizeof(int) * (size_t)(gt_total > 0 ? gt_total : 1), 64);
    w.m_idx = reinterpret_cast<int *>(p + off);
    off += g4;
    w.m_cls = reinterpret_cast<int *>(p + off);
    off += g4;
    w.m_iou = reinterpret_cast<float *>(p + off);
    off += g4;
    w.m_dfl = reinterpret_cast<float *>(p + off);
    off += g4;
    w.part = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * (size_t)n_images * cls_tiles, 64);
    w.total_bytes = off;
    return w;
}

// ------------------------------------------------------------------------------------------
// assign_kernel
// ------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(kAssignThreads)
assign_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, const float *__restrict__ anchors,
              const float *__restrict__ strides, const float *__restrict__ gt, const int *__restrict__ gt_off,
              unsigned long long *__restrict__ best, T *__restrict__ grad) {
    constexpr int TILE = kAssignThreads * VW;
    __shared__ float4 s_ctr[TILE];                        // cx, cy, cx^2+cy^2 of the tile's anchors
    __shared__ unsigned long long s_key[kAssignThreads];

    const int n = blockIdx.y;
    const int tile0 = blockIdx.x * TILE;
    const int a0 = tile0 + threadIdx.x * VW;
    const size_t img = (size_t)n * n_ch * n_anchors;
    const int g_begin = gt_off[n];
    const int m_img = gt_off[n + 1] - g_begin;

    if (a0 < n_anchors) {
        float dist[4][VW];
#pragma unroll
        for (int side = 0; side < 4; ++side) {
            Group<T, VW> row[kRegMax];
#pragma unroll
            for (int j = 0; j < kRegMax; ++j)
                row[j].load(preds + img + (size_t)(side * kRegMax + j) * n_anchors + a0);
            if (grad != nullptr) {
#pragma unroll
                for (int j = 0; j < kRegMax; ++j)
                    Group<T, VW>::store_zero(grad + img + (size_t)(side * kRegMax + j) * n_anchors + a0);
            }
            if (m_img > 0) {
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    float x[kRegMax], p[kRegMax];
#pragma unroll
                    for (int j = 0; j < kRegMax; ++j) x[j] = row[j].get(v);
                    dist[side][v] = dfl_expectation16(x, p);
                }
            }
        }
        if (m_img > 0) {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                const float ax = __ldg(anchors + a0 + v), ay = __ldg(anchors + n_anchors + a0 + v);
                const float s = __ldg(strides + a0 + v);
                const PredBox b = decode_box(ax, ay, s, dist[0][v], dist[1][v], dist[2][v], dist[3][v]);
                const float pn = __fadd_rn(__fmul_rn(b.cx, b.cx), __fmul_rn(b.cy, b.cy));
                s_ctr[threadIdx.x * VW + v] = make_float4(b.cx, b.cy, pn, 0.f);
            }
        }
    }
    if (m_img == 0) return;                                // uniform per CTA
    __syncthreads();

    const int tile_n = min(TILE, n_anchors - tile0);
    // GT chunks of up to 128: thread <-> one GT; when the chunk is small the anchors of the tile are
    // split into slices so that all warps have work.  Lanes of a warp share the slice, so every
    // shared-memory read below is a broadcast.
    for (int g0 = 0; g0 < m_img; g0 += kAssignThreads) {
        const int m_chunk = min(kAssignThreads, m_img - g0);
        const int m_pad = (m_chunk + 31) & ~31;
        const int n_slice = kAssignThreads / m_pad;
        const int g = threadIdx.x % m_pad;
        const int slice = threadIdx.x / m_pad;
        unsigned long long key = kNoKey;
        if (g < m_chunk && slice < n_slice) {
            const float gx = __ldg(gt + (size_t)(g_begin + g0 + g) * 5 + 0);
            const float gy = __ldg(gt + (size_t)(g_begin + g0 + g) * 5 + 1);
            // ATen _euclidean_dist: [-2gx, -2gy, |g|^2, 1] . [px, py, 1, |p|^2], K = 4, FMA chain k = 0..3
            const float c0 = __fmul_rn(-2.f, gx), c1 = __fmul_rn(-2.f, gy);
            const float gn = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
            const int per = (tile_n + n_slice - 1) / n_slice;
            const int j_begin = slice * per, j_end = min(tile_n, j_begin + per);
            float best_d2 = __int_as_float(0x7f800000), best_s = __int_as_float(0x7f800000);
            int best_j = -1;
#pragma unroll 4
            for (int j = j_begin; j < j_end; ++j) {
                const float4 c = s_ctr[j];
                float d2 = __fmul_rn(c0, c.x);
                d2 = __fmaf_rn(c1, c.y, d2);
                d2 = __fadd_rn(d2, gn);
                d2 = __fadd_rn(d2, c.z);
                d2 = d2 < 0.f ? 0.f : d2;                  // clamp_min(0)
                if (d2 < best_d2) {                        // sqrt is monotone: only then can it win
                    const float s = __fsqrt_rn(d2);
                    if (s < best_s) { best_s = s; best_d2 = d2; best_j = j; }
                }
            }
            if (best_j >= 0)
                key = ((unsigned long long)(__float_as_uint(best_s) & 0x7fffffffu) << 32) |
                      (unsigned int)(tile0 + best_j);
        }
        s_key[threadIdx.x] = key;
        __syncthreads();
        if (threadIdx.x < m_chunk) {
            unsigned long long k = s_key[threadIdx.x];
            for (int sl = 1; sl < n_slice; ++sl) {
                const unsigned long long o = s_key[sl * m_pad + threadIdx.x];
                k = o < k ? o : k;
            }
            // smallest (distance, anchor) wins == first-min argmin; stored inverted so that the
            // workspace can be armed with a plain memset(0).
            if (k != kNoKey) atomicMax(best + g_begin + g0 + threadIdx.x, ~k);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// match_kernel: one warp per GT
// ------------------------------------------------------------------------------------------
struct GtTerms {
    float g_lo, g_hi;   // gradient of the two logits this lane holds (bins lane%16 of sides lane/16 and 2+lane/16)
    float dfl;          // sum over the four sides of the DFL term (all lanes)
    float iou;          // soft target (all lanes)
};

// d(total)/d(iou_m): the QFL is linear in the target, so this is independent of the target value
// (src/model/losses.py:53-56 differentiated w.r.t. target_scores).
__device__ __forceinline__ float qfl_dloss_dtarget(float logit, float k_cls) {
    const float p = __fdiv_rn(1.f, 1.f + expf(-logit));
    const float q = 1.f - p;
    return -k_cls * (q * q * logf(p + kEpsLog) - p * p * logf(q + kEpsLog));
}

template <typename T>
__device__ __forceinline__ GtTerms gt_terms(const T *__restrict__ img, int n_anchors, int idx,
                                            const float *__restrict__ anchors, const float *__restrict__ strides,
                                            const float *__restrict__ g5, float k_dfl, float k_cls, int nc) {
    const int lane = threadIdx.x & 31;
    const int bin = lane & 15;
    const int half = lane >> 4;
    const float z_lo = load_as_float(img + (size_t)lane * n_anchors + idx);          // sides 0,1
    const float z_hi = load_as_float(img + (size_t)(lane + 32) * n_anchors + idx);   // sides 2,3
    const float gcx = __ldg(g5 + 0), gcy = __ldg(g5 + 1), gw = __ldg(g5 + 2), gh = __ldg(g5 + 3);
    int cls = (int)__ldg(g5 + 4);                                                     // .long(): truncation
    cls = min(max(cls, 0), nc - 1);
    const float z_cls = load_as_float(img + (size_t)(4 * kRegMax + cls) * n_anchors + idx);
    const float ax = __ldg(anchors + idx), ay = __ldg(anchors + n_anchors + idx), s = __ldg(strides + idx);

    // softmax over each 16-lane half (xor offsets 8,4,2,1 stay inside a half)
    float m_lo = z_lo, m_hi = z_hi;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, o));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, o));
    }
    const float e_lo = expf(z_lo - m_lo), e_hi = expf(z_hi - m_hi);
    float s_lo = e_lo, s_hi = e_hi;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        s_lo = __fadd_rn(s_lo, __shfl_xor_sync(0xffffffffu, s_lo, o));
        s_hi = __fadd_rn(s_hi, __shfl_xor_sync(0xffffffffu, s_hi, o));
    }
    const float p_lo = __fdiv_rn(e_lo, s_lo), p_hi = __fdiv_rn(e_hi, s_hi);
    float d_lo = __fmul_rn(p_lo, (float)bin), d_hi = __fmul_rn(p_hi, (float)bin);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        d_lo += __shfl_xor_sync(0xffffffffu, d_lo, o);
        d_hi += __shfl_xor_sync(0xffffffffu, d_hi, o);
    }
    const float dl = __shfl_sync(0xffffffffu, d_lo, 0), dt = __shfl_sync(0xffffffffu, d_lo, 16);
    const float dr = __shfl_sync(0xffffffffu, d_hi, 0), db = __shfl_sync(0xffffffffu, d_hi, 16);
    const PredBox b = decode_box(ax, ay, s, dl, dt, dr, db);

    // ---- IoU soft target, reference formula (src/model/losses.py:17-40) and its gradient ------
    const float hw = b.w * 0.5f, hh = b.h * 0.5f;
    const float ax1 = b.cx - hw, ay1 = b.cy - hh, ax2 = b.cx + hw;
    const float ay2 = b.h + b.cy * 0.5f;                  // sic: losses.py:20
    const float bx1 = gcx - gw * 0.5f, by1 = gcy - gh * 0.5f, bx2 = gcx + gw * 0.5f, by2 = gcy + gh * 0.5f;
    const float ix1 = fmaxf(ax1, bx1), iy1 = fmaxf(ay1, by1), ix2 = fminf(ax2, bx2), iy2 = fminf(ay2, by2);
    const float iw_raw = ix2 - ix1, ih_raw = iy2 - iy1;
    const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
    const float inter = iw * ih;
    const float aw = ax2 - ax1, ah = ay2 - ay1;
    const float area1 = aw * ah, area2 = (bx2 - bx1) * (by2 - by1);
    const float uni = area1 + area2 - inter;
    const float den = uni + kEpsIou;
    const float iou = inter / den;

    // backward of the above in autograd's conventions: clamp passes where the input >= 0, max/min
    // route to the larger/smaller argument and split evenly on ties.
    const float g_iou = qfl_dloss_dtarget(z_cls, k_cls);
    const float d_inter = g_iou * (1.f / den + inter / (den * den));   // d iou/d inter, incl. the -inter in the union
    const float d_area1 = -g_iou * inter / (den * den);
    const float d_iw = (iw_raw >= 0.f) ? d_inter * ih : 0.f;
    const float d_ih = (ih_raw >= 0.f) ? d_inter * iw : 0.f;
    auto pick_max = [](float a, float o) { return a > o ? 1.f : (a == o ? 0.5f : 0.f); };   // weight on `a`
    auto pick_min = [](float a, float o) { return a < o ? 1.f : (a == o ? 0.5f : 0.f); };
    float d_ax1 = -d_iw * pick_max(ax1, bx1) - d_area1 * ah;
    float d_ax2 = d_iw * pick_min(ax2, bx2) + d_area1 * ah;
    float d_ay1 = -d_ih * pick_max(ay1, by1) - d_area1 * aw;
    float d_ay2 = d_ih * pick_min(ay2, by2) + d_area1 * aw;
    // ax1 = cx - w/2, ax2 = cx + w/2, ay1 = cy - h/2, ay2 = h + cy/2
    const float d_cx = d_ax1 + d_ax2;
    const float d_w = 0.5f * (d_ax2 - d_ax1);
    const float d_cy = d_ay1 + 0.5f * d_ay2;
    const float d_h = d_ay2 - 0.5f * d_ay1;
    // cx = (x1+x2)/2, w = x2-x1 ;  x1 = (ax-dl)*s, x2 = (ax+dr)*s
    const float d_x1 = 0.5f * d_cx - d_w, d_x2 = 0.5f * d_cx + d_w;
    const float d_y1 = 0.5f * d_cy - d_h, d_y2 = 0.5f * d_cy + d_h;
    const float d_dl = -d_x1 * s, d_dt = -d_y1 * s, d_dr = d_x2 * s, d_db = d_y2 * s;

    // ---- DFL target bins (src/model/losses.py:226-246) and loss rows (:63-78) -----------------
    const float t_l = ax - bx1 / s, t_t = ay - by1 / s, t_r = bx2 / s - ax, t_b = by2 / s - ay;
    const float hi_clamp = (float)(kRegMax - 1 - 0.01);
    const float t_lo = fminf(fmaxf(half == 0 ? t_l : t_t, 0.f), hi_clamp);   // this lane's side in the low register
    const float t_hi = fminf(fmaxf(half == 0 ? t_r : t_b, 0.f), hi_clamp);   // ... and in the high register
    const int bl_lo = (int)t_lo, bl_hi = (int)t_hi;
    const float wl_lo = (float)(bl_lo + 1) - t_lo, wr_lo = t_lo - (float)bl_lo;
    const float wl_hi = (float)(bl_hi + 1) - t_hi, wr_hi = t_hi - (float)bl_hi;
    // log-softmax value of this lane's bin
    const float lp_lo = (z_lo - m_lo) - logf(s_lo), lp_hi = (z_hi - m_hi) - logf(s_hi);
    const int base = lane & 16;
    const float ce_lo = -(__shfl_sync(0xffffffffu, lp_lo, base + bl_lo) * wl_lo +
                          __shfl_sync(0xffffffffu, lp_lo, base + bl_lo + 1) * wr_lo);
    const float ce_hi = -(__shfl_sync(0xffffffffu, lp_hi, base + bl_hi) * wl_hi +
                          __shfl_sync(0xffffffffu, lp_hi, base + bl_hi + 1) * wr_hi);
    // ce_lo is uniform within a half: lanes 0-15 hold side 0 / 2, lanes 16-31 side 1 / 3
    const float ce_half = ce_lo + ce_hi;
    const float dfl = ce_half + __shfl_xor_sync(0xffffffffu, ce_half, 16);

    GtTerms r;
    r.dfl = dfl;
    r.iou = iou;
    const float dd_lo = half == 0 ? d_dl : d_dt, dd_hi = half == 0 ? d_dr : d_db;
    const float dk_lo = half == 0 ? dl : dt, dk_hi = half == 0 ? dr : db;
    const float oh_lo = (bin == bl_lo ? wl_lo : 0.f) + (bin == bl_lo + 1 ? wr_lo : 0.f);
    const float oh_hi = (bin == bl_hi ? wl_hi : 0.f) + (bin == bl_hi + 1 ? wr_hi : 0.f);
    r.g_lo = k_dfl * ((wl_lo + wr_lo) * p_lo - oh_lo) + dd_lo * p_lo * ((float)bin - dk_lo);
    r.g_hi = k_dfl * ((wl_hi + wr_hi) * p_hi - oh_hi) + dd_hi * p_hi * ((float)bin - dk_hi);
    return r;
}

template <typename T>
__global__ void __launch_bounds__(128)
match_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, int nc, const float *__restrict__ anchors,
             const float *__restrict__ strides, const float *__restrict__ gt, const int *__restrict__ gt_off,
             const unsigned long long *__restrict__ best, float k_dfl_num, float k_cls, T *__restrict__ grad,
             int *__restrict__ m_idx, int *__restrict__ m_cls, float *__restrict__ m_iou, float *__restrict__ m_dfl,
             int *__restrict__ out_idx, float *__restrict__ out_iou) {
    const int n = blockIdx.y;
    const int g_begin = gt_off[n];
    const int m_img = gt_off[n + 1] - g_begin;
    const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (m >= m_img) return;
    const int lane = threadIdx.x & 31;
    const T *img = preds + (size_t)n * n_ch * n_anchors;
    auto idx_of = [&](int mm) {
        const unsigned long long inv = best[g_begin + mm];
        return inv == 0ull ? 0 : (int)(unsigned int)(~inv & 0xffffffffull);   // no finite distance at all -> anchor 0
    };
    const int idx = idx_of(m);
    const float k_dfl = k_dfl_num / (float)m_img;         // lambda_dfl / (N * 4 * M)

    // Which GTs of this image share my anchor?  owner = lowest such m (writes the summed gradient),
    // winner = highest (its class row is the anchor's QFL target: "last write wins", losses.py:261).
    int first = m, last = m;
    for (int mm = lane; mm < m_img; mm += 32) {
        if (idx_of(mm) == idx) { first = min(first, mm); last = max(last, mm); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }

    GtTerms mine = gt_terms(img, n_anchors, idx, anchors, strides, gt + (size_t)(g_begin + m) * 5, k_dfl, k_cls, nc);
    if (lane == 0) {
        int cls = (int)__ldg(gt + (size_t)(g_begin + m) * 5 + 4);
        cls = min(max(cls, 0), nc - 1);
        m_idx[g_begin + m] = idx;
        m_cls[g_begin + m] = (last == m) ? cls : -1;
        m_iou[g_begin + m] = mine.iou;
        m_dfl[g_begin + m] = mine.dfl;
        if (out_idx) out_idx[g_begin + m] = idx;
        if (out_iou) out_iou[g_begin + m] = mine.iou;
    }
    if (grad == nullptr || first != m) return;
    float g_lo = mine.g_lo, g_hi = mine.g_hi;
    if (last != m) {
        for (int mm = m + 1; mm <= last; ++mm) {           // warp-uniform loop
            if (idx_of(mm) != idx) continue;
            const GtTerms o = gt_terms(img, n_anchors, idx, anchors, strides, gt + (size_t)(g_begin + mm) * 5, k_dfl,
                                       k_cls, nc);
            g_lo += o.g_lo;
            g_hi += o.g_hi;
        }
    }
    T *gimg = grad + (size_t)n * n_ch * n_anchors;
    store_from_float(gimg + (size_t)lane * n_anchors + idx, g_lo);
    store_from_float(gimg + (size_t)(lane + 32) * n_anchors + idx, g_hi);
}

// ------------------------------------------------------------------------------------------
// cls_loss_kernel
// ------------------------------------------------------------------------------------------
// One element of the quality focal loss and its gradient w.r.t. the logit
// (src/model/losses.py:51-56):  loss += t (1-p)^2 log(p+e) + (1-t) p^2 log(1-p+e)   (sign applied later)
__device__ __forceinline__ void qfl_elem(float x, float t, float k_cls, float &acc, float &g) {
    const float e = expf(-x);
    const float p = __fdividef(1.f, 1.f + e);
    const float q = 1.f - p;
    const float lq = logf(q + kEpsLog);
    if (t == 0.f && q > 1e-4f) {
        // target 0 and 1-p far above the 1e-12 guard: -(d/dx) = k p^2 (p - 2 q log q)
        acc += p * p * lq;
        g = k_cls * p * p * (p - 2.f * q * lq);
    } else {
        const float lp = logf(p + kEpsLog);
        const float u = 1.f - t;
        acc += t * q * q * lp + u * p * p * lq;
        const float dpos = t * (-2.f * q * lp + q * q / (p + kEpsLog));
        const float dneg = u * (2.f * p * lq - p * p / (q + kEpsLog));
        g = -k_cls * (dpos + dneg) * p * q;
    }
}

template <typename T, int VW, bool WRITE_GRAD>
__global__ void __launch_bounds__(kClsThreads)
cls_loss_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, int nc, const int *__restrict__ gt_off,
                const int *__restrict__ m_idx, const int *__restrict__ m_cls, const float *__restrict__ m_iou,
                const float *__restrict__ m_dfl, float k_cls, float lambda_cls, float lambda_dfl,
                T *__restrict__ grad, float *__restrict__ part, unsigned int *__restrict__ ticket,
                float *__restrict__ out_loss, float *__restrict__ out_per_image) {
    constexpr int TILE = kClsThreads * VW;
    __shared__ int s_cls[TILE];
    __shared__ float s_iou[TILE];
    __shared__ float s_red[kClsThreads / 32];
    __shared__ bool s_last;

    const int n = blockIdx.y;
    const int tile0 = blockIdx.x * TILE;
    const int a0 = tile0 + threadIdx.x * VW;
    const int g_begin = gt_off[n];
    const int m_img = gt_off[n + 1] - g_begin;

#pragma unroll
    for (int v = 0; v < VW; ++v) s_cls[threadIdx.x * VW + v] = -1;
    __syncthreads();
    for (int m = threadIdx.x; m < m_img; m += kClsThreads) {
        const int c = m_cls[g_begin + m];
        const int j = m_idx[g_begin + m] - tile0;
        if (c >= 0 && j >= 0 && j < TILE) {               // one owner per anchor: no write conflict
            s_cls[j] = c;
            s_iou[j] = m_iou[g_begin + m];
        }
    }
    __syncthreads();

    float acc = 0.f;
    if (a0 < n_anchors) {
        int t_cls[VW];
        float t_iou[VW];
        bool any = false;
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            t_cls[v] = s_cls[threadIdx.x * VW + v];
            t_iou[v] = t_cls[v] >= 0 ? s_iou[threadIdx.x * VW + v] : 0.f;
            any |= t_cls[v] >= 0;
        }
        const size_t base = ((size_t)n * n_ch + 4 * kRegMax) * n_anchors + a0;
        constexpr int U = 4;
        int c = 0;
        for (; c + U <= nc; c += U) {
            Group<T, VW> row[U];
#pragma unroll
            for (int u = 0; u < U; ++u) row[u].load(preds + base + (size_t)(c + u) * n_anchors);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float g[VW];
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    const float t = (any && t_cls[v] == c + u) ? t_iou[v] : 0.f;
                    qfl_elem(row[u].get(v), t, k_cls, acc, g[v]);
                }
                if (WRITE_GRAD) Group<T, VW>::store(grad + base + (size_t)(c + u) * n_anchors, g);
            }
        }
        for (; c < nc; ++c) {
            Group<T, VW> row;
            row.load(preds + base + (size_t)c * n_anchors);
            float g[VW];
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                const float t = (any && t_cls[v] == c) ? t_iou[v] : 0.f;
                qfl_elem(row.get(v), t, k_cls, acc, g[v]);
            }
            if (WRITE_GRAD) Group<T, VW>::store(grad + base + (size_t)c * n_anchors, g);
        }
    }

    // CTA partial -> workspace; the last CTA to finish reduces everything in a fixed order.
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kClsThreads / 32; ++w) s += s_red[w];
        part[(size_t)n * gridDim.x + blockIdx.x] = s;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        s_last = (done == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // final reduction: thread i owns images i, i+128, ...; then a fixed-shape tree over the CTA.
    const int n_images = gridDim.y;
    const int tiles = gridDim.x;
    double dfl_sum = 0.0, cls_sum = 0.0;
    float fg = 0.f;
    for (int b = threadIdx.x; b < n_images; b += kClsThreads) {
        double c_img = 0.0;
        for (int t = 0; t < tiles; ++t) c_img += (double)__ldcg(part + (size_t)b * tiles + t);
        const float cls_b = (float)(-c_img / (double)n_anchors);
        const int gb = gt_off[b], mb = gt_off[b + 1] - gb;
        double d_img = 0.0;
        for (int m = 0; m < mb; ++m) {
            d_img += (double)__ldcg(m_dfl + gb + m);
            fg += (__ldcg(m_cls + gb + m) >= 0) ? 1.f : 0.f;
        }
        const float dfl_b = mb > 0 ? (float)(d_img / (4.0 * (double)mb)) : 0.f;
        if (out_per_image) {
            out_per_image[b] = dfl_b;
            out_per_image[n_images + b] = cls_b;
        }
        dfl_sum += (double)dfl_b;
        cls_sum += (double)cls_b;
    }
    __shared__ double s_d[kClsThreads], s_c[kClsThreads];
    __shared__ float s_f[kClsThreads];
    s_d[threadIdx.x] = dfl_sum;
    s_c[threadIdx.x] = cls_sum;
    s_f[threadIdx.x] = fg;
    __syncthreads();
    for (int o = kClsThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_d[threadIdx.x] += s_d[threadIdx.x + o];
            s_c[threadIdx.x] += s_c[threadIdx.x + o];
            s_f[threadIdx.x] += s_f[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mean_dfl = (float)(s_d[0] / (double)n_images);
        const float mean_cls = (float)(s_c[0] / (double)n_images);
        out_loss[0] = lambda_dfl * mean_dfl + lambda_cls * mean_cls;   // losses.py:275
        out_loss[1] = mean_dfl;
        out_loss[2] = mean_cls;
        out_loss[3] = s_f[0];
        out_loss[4] = out_loss[5] = out_loss[6] = out_loss[7] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// grad *= *scale
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T *__restrict__ g, size_t n, const float *__restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.f) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        store_from_float(g + i, load_as_float(g + i) * s);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <typename T>
static bool vector_ok(const void *preds, const void *grad, int n_anchors) {
    constexpr int VW = ElemsPer16<T>::value;
    return n_anchors % VW == 0 && aligned16(preds) && (grad == nullptr || aligned16(grad));
}

// Optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launching
// stream around each of the three kernels.  Off by default; never on during a timed benchmark step.
static bool g_stage_timing = false;
static cudaEvent_t g_stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};

static int stage_mark(int i, cudaStream_t st) {
    if (!g_stage_timing) return YB_OK;
    if (g_stage_ev[i] == nullptr) YB_CUDA(cudaEventCreate(&g_stage_ev[i]));
    YB_CUDA(cudaEventRecord(g_stage_ev[i], st));
    return YB_OK;
}

template <typename T, int VW>
static int launch_loss(const T *preds, int n_images, int nc, int n_anchors, const float *anchors, const float *strides,
                       const float *gt, const int32_t *gt_off, int gt_total, int gmax, float lambda_cls,
                       float lambda_dfl, T *grad, float *out_loss, int32_t *out_idx, float *out_iou,
                       float *out_per_image, const LossWorkspace &w, cudaStream_t st) {
    const int n_ch = 4 * kRegMax + nc;
    const float k_cls = lambda_cls / ((float)n_images * (float)n_anchors);
    const float k_dfl_num = lambda_dfl / ((float)n_images * 4.f);
    YB_CUDA(cudaMemsetAsync(w.ticket, 0, w.zero_bytes, st));
    if (int rc = stage_mark(0, st)) return rc;
    {
        constexpr int TILE = kAssignThreads * VW;
        dim3 grid((n_anchors + TILE - 1) / TILE, n_images);
        assign_kernel<T, VW><<<grid, kAssignThreads, 0, st>>>(preds, n_ch, n_anchors, anchors, strides, gt, gt_off,
                                                              w.best, grad);
        YB_CUDA(cudaGetLastError());
    }
    if (int rc = stage_mark(1, st)) return rc;
    if (gt_total > 0 && gmax > 0) {
        dim3 grid((gmax + 3) / 4, n_images);
        match_kernel<T><<<grid, 128, 0, st>>>(preds, n_ch, n_anchors, nc, anchors, strides, gt, gt_off, w.best,
                                              k_dfl_num, k_cls, grad, w.m_idx, w.m_cls, w.m_iou, w.m_dfl, out_idx,
                                              out_iou);
        YB_CUDA(cudaGetLastError());
    }
    if (int rc = stage_mark(2, st)) return rc;
    {
        constexpr int TILE = kClsThreads * VW;
        dim3 grid((n_anchors + TILE - 1) / TILE, n_images);
        if (grad != nullptr)
            cls_loss_kernel<T, VW, true><<<grid, kClsThreads, 0, st>>>(preds, n_ch, n_anchors, nc, gt_off, w.m_idx,
                                                                       w.m_cls, w.m_iou, w.m_dfl, k_cls, lambda_cls,
                                                                       lambda_dfl, grad, w.part, w.ticket, out_loss,
                                                                       out_per_image);
        else
            cls_loss_kernel<T, VW, false><<<grid, kClsThreads, 0, st>>>(preds, n_ch, n_anchors, nc, gt_off, w.m_idx,
                                                                        w.m_cls, w.m_iou, w.m_dfl, k_cls, lambda_cls,
                                                                        lambda_dfl, grad, w.part, w.ticket, out_loss,
                                                                        out_per_image);
        YB_CUDA(cudaGetLastError());
    }
    if (int rc = stage_mark(3, st)) return rc;
    return YB_OK;
}

static int cls_tiles_for(int n_anchors, int dtype, bool vec) {
    const int vw = vec ? (dtype == YB_BF16 ? 8 : 4) : 1;
    const int tile = kClsThreads * vw;
    return (n_anchors + tile - 1) / tile;
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_loss_workspace_bytes(int n_images, int n_anchors, int gt_total, int dtype) {
    if (n_images <= 0 || n_anchors <= 0 || gt_total < 0) return 0;
    // sized for the scalar fall-back (most tiles); the vector path needs less
    return carve(nullptr, n_images, cls_tiles_for(n_anchors, dtype, false), gt_total).total_bytes;
}

extern "C" int yb_loss_fwd_bwd(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                               const float *anchors, const float *strides, const float *gt,
                               const int32_t *gt_offsets, int gt_total, int gmax, float lambda_cls, float lambda_dfl,
                               void *grad_preds, float *out_loss, int32_t *out_idx, float *out_iou,
                               float *out_per_image, void *workspace, size_t workspace_bytes, void *stream) {
    YB_REQUIRE(preds && anchors && strides && gt_offsets && out_loss && workspace, "yb_loss_fwd_bwd: null pointer");
    YB_REQUIRE(gt_total == 0 || gt != nullptr, "yb_loss_fwd_bwd: gt is null but gt_total > 0");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0 && gt_total >= 0 && gmax >= 0, "yb_loss_fwd_bwd: bad sizes");
    YB_REQUIRE(reg_max == kRegMax, "yb_loss_fwd_bwd: reg_max must be %d (got %d)", kRegMax, reg_max);
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_loss_fwd_bwd: dtype must be YB_F32 or YB_BF16");
    YB_REQUIRE(n_images <= 65535, "yb_loss_fwd_bwd: at most 65535 images per call");
    if (workspace_bytes < yb_loss_workspace_bytes(n_images, n_anchors, gt_total, dtype)) {
        set_error("yb_loss_fwd_bwd: workspace %zu B < required %zu B", workspace_bytes,
                  yb_loss_workspace_bytes(n_images, n_anchors, gt_total, dtype));
        return YB_ERR_WORKSPACE;
    }
    if (!aligned16(workspace)) {
        set_error("yb_loss_fwd_bwd: workspace must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == YB_F32) {
        const bool vec = vector_ok<float>(preds, grad_preds, n_anchors);
        const LossWorkspace w = carve(workspace, n_images, cls_tiles_for(n_anchors, dtype, vec), gt_total);
        if (vec)
            return launch_loss<float, 4>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                         gt_offsets, gt_total, gmax, lambda_cls, lambda_dfl, (float *)grad_preds,
                                         out_loss, out_idx, out_iou, out_per_image, w, st);
        return launch_loss<float, 1>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt, gt_offsets,
                                     gt_total, gmax, lambda_cls, lambda_dfl, (float *)grad_preds, out_loss, out_idx,
                                     out_iou, out_per_image, w, st);
    }
    const bool vec = vector_ok<__nv_bfloat16>(preds, grad_preds, n_anchors);
    const LossWorkspace w = carve(workspace, n_images, cls_tiles_for(n_anchors, dtype, vec), gt_total);
    if (vec)
        return launch_loss<__nv_bfloat16, 8>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides,
                                             gt, gt_offsets, gt_total, gmax, lambda_cls, lambda_dfl,
                                             (__nv_bfloat16 *)grad_preds, out_loss, out_idx, out_iou, out_per_image, w,
                                             st);
    return launch_loss<__nv_bfloat16, 1>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                         gt_offsets, gt_total, gmax, lambda_cls, lambda_dfl,
                                         (__nv_bfloat16 *)grad_preds, out_loss, out_idx, out_iou, out_per_image, w, st);
}

extern "C" int yb_stage_timing(int enable) {
    g_stage_timing = enable != 0;
    return YB_OK;
}

extern "C" int yb_loss_last_stage_ms(float *out_ms_host) {
    YB_REQUIRE(out_ms_host != nullptr, "yb_loss_last_stage_ms: null pointer");
    for (int i = 0; i < 4; ++i) YB_REQUIRE(g_stage_ev[i] != nullptr, "yb_loss_last_stage_ms: no timed call has run");
    YB_CUDA(cudaEventSynchronize(g_stage_ev[3]));
    for (int i = 0; i < 3; ++i) YB_CUDA(cudaEventElapsedTime(out_ms_host + i, g_stage_ev[i], g_stage_ev[i + 1]));
    return YB_OK;
}

extern "C" int yb_scale_grad(void *grad, int dtype, size_t n_elements, const float *scale, void *stream) {
    YB_REQUIRE(grad && scale, "yb_scale_grad: null pointer");
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_scale_grad: bad dtype");
    if (n_elements == 0) return YB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = 148 * 8;
    if (dtype == YB_F32)
        scale_kernel<float><<<blocks, 256, 0, st>>>((float *)grad, n_elements, scale);
    else
        scale_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16 *)grad, n_elements, scale);
    YB_CUDA(cudaGetLastError());
    return YB_OK;
}

extern "C" int yb_loss_fwd_bwd_host(const void *preds_host, int dtype, int n_images, int nc, int reg_max,
                                    int n_anchors, const float *anchors, const float *strides, const float *gt_host,
                                    const int32_t *gt_offsets_host, int gt_total, int gmax, float lambda_cls,
                                    float lambda_dfl, void *preds_dev, float *gt_dev, int32_t *gt_offsets_dev,
                                    void *grad_dev, float *out_loss_dev, float *out_loss_host, void *grad_host,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    YB_REQUIRE(preds_host && preds_dev && gt_offsets_host && gt_offsets_dev && out_loss_dev && out_loss_host,
               "yb_loss_fwd_bwd_host: null pointer");
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_loss_fwd_bwd_host: bad dtype");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0, "yb_loss_fwd_bwd_host: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t esz = dtype == YB_BF16 ? 2 : 4;
    const size_t n_el = (size_t)n_images * (4 * kRegMax + nc) * n_anchors;
    YB_CUDA(cudaMemcpyAsync(preds_dev, preds_host, n_el * esz, cudaMemcpyHostToDevice, st));
    if (gt_total > 0) {
        YB_REQUIRE(gt_host && gt_dev, "yb_loss_fwd_bwd_host: gt is null but gt_total > 0");
        YB_CUDA(cudaMemcpyAsync(gt_dev, gt_host, sizeof(float) * 5 * (size_t)gt_total, cudaMemcpyHostToDevice, st));
    }
    YB_CUDA(cudaMemcpyAsync(gt_offsets_dev, gt_offsets_host, sizeof(int32_t) * (size_t)(n_images + 1),
                            cudaMemcpyHostToDevice, st));
    const int rc = yb_loss_fwd_bwd(preds_dev, dtype, n_images, nc, reg_max, n_anchors, anchors, strides, gt_dev,
                                   gt_offsets_dev, gt_total, gmax, lambda_cls, lambda_dfl, grad_dev, out_loss_dev,
                                   nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
    if (rc != YB_OK) return rc;
    YB_CUDA(cudaMemcpyAsync(out_loss_host, out_loss_dev, sizeof(float) * 8, cudaMemcpyDeviceToHost, st));
    if (grad_host != nullptr && grad_dev != nullptr)
        YB_CUDA(cudaMemcpyAsync(grad_host, grad_dev, n_el * esz, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    return YB_OK;
}
