// Task-aligned variant of the training path (sm_100a): DFL decode -> task-aligned assignment
// (pairwise CIoU over anchors x GT, metric = score^alpha * IoU^beta, per-GT top-k, conflicts to the
// larger IoU) -> CIoU + DFL + BCE loss and its backward.
//
// The reference has NO counterpart for this tier: the specification is the in-repo
// oracle `oracle/tal_oracle.py` (SURVEY.md §8(a')); results are "parity vs the in-repo oracle".
//
// Two ABI calls so that the normaliser can be exchanged between them (the path's one real exchange step).
// Everything that does not need the normaliser runs in the FIRST call, so a collective issued between the
// two has nothing left to hide behind but also nothing left to wait for:
//   yb_tal_assign   tal_decode_kernel      streaming: reads the 4 x 16 box rows once (128-bit loads) and leaves the
//                                          decoded pixel box of every anchor in the workspace (16 B per anchor,
//                                          L2-resident for the next kernel)
//                   tal_gt_kernel          one warp per GT over the whole image.  (1) candidates: the anchors whose centre
//                                          lies inside the GT (rectangle enumeration on a verified grid hint, group-extent
//                                          skip + ballot compaction otherwise), alignment metric 64 anchors at a time
//                                          (approximate-reciprocal arithmetic: it only RANKS), running top-k (metric desc,
//                                          anchor asc) by REDUX rounds with the candidate list in registers, 64-bit
//                                          atomicMax (overlap, ~gt) per anchor for conflicts.  (2) the foreground terms of
//                                          the k selected anchors, two at a time (one per half-warp), before anybody knows
//                                          whether the GT keeps them: every term is LINEAR in the anchor's target score t,
//                                          so CIoU / DFL loss and the gradient of the anchor's 64 box logits are left per
//                                          unit of t (compact buffer); the scattered gathers of one warp run under the
//                                          ranking arithmetic of the others
//                                          (3) once every selection is in: which selected anchors did each GT keep, the
//                                          GT's max metric / overlap -> target score t per slot (or "not kept"), anchor ->
//                                          slot map; target scores summed in fixed point (integer atomics: order-
//                                          independent); the last unit -> [sum of target scores, #foreground] of this rank
//                                          (+ the peer stores)
//   yb_tal_loss     tal_cls_kernel         dense BCE-with-logits at target 0 + gradient; writes the box rows of the
//                                          gradient too: zero, or the foreground anchor's 64 values * t / normaliser
//                                          (predicated loads through the anchor -> slot map, no scattered stores)
//                   tal_finalize_kernel    patches the one positive class cell of every foreground anchor, scales the
//                                          per-slot terms by t, fixed-order reduction -> loss scalars
// The anchors x GT overlap / metric matrices never exist.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace yb {

int launch_peer_wait(const yb_peer_exchange &px, float *out2, unsigned int *timed_out, cudaStream_t st);   // csrc/peer.cu
int check_peer(const yb_peer_exchange *px, const char *who);
constexpr int kPeerSlotsTal = 4;                 // = kPeerSlots of csrc/peer.cu

#ifndef YB_TAL_DECODE_MINBLOCKS      // 0: 6 resident CTAs per SM for fp32 rows, 4 for bf16 rows (measured)
#define YB_TAL_DECODE_MINBLOCKS 0
#endif
#ifndef YB_TAL_PDL                    // programmatic dependent launch between the four kernels of the step
#define YB_TAL_PDL 1
#endif
constexpr int kTalThreads = 128;
constexpr int kTalMaxK = 16;
constexpr float kEpsCiou = 1e-7f;
constexpr float kEpsIn = 1e-9f;
constexpr float kEpsNorm = 1e-9f;
constexpr float kFourOverPi2 = 0.40528473456935109f;
constexpr int kTalFinThreadsDecl = 256;
constexpr int kTalStatAcc = 64;            // sub-accumulators of the target-score sum / foreground count (spreads same-address atomics)
constexpr double kTalFix = 4294967296.0;   // 2^32: target scores are summed in fixed point (order-independent, exact)
// CTAs per anchor tile in the dense pass (each takes 1/SPLIT of the class and box rows): measured best 8 for
// 512-anchor (fp32) tiles, 4 for 1024-anchor (bf16) tiles; the scalar fall-back (128-anchor tiles) uses 8 too
__host__ __device__ constexpr int tal_cls_split(int tile) { return tile >= 1024 ? 4 : 8; }
#ifndef YB_TAL_CLS_UNROLL
#define YB_TAL_CLS_UNROLL 4
#endif
struct TalWorkspace {
    // the counters: zero when yb_tal_assign starts -- memset there, or (YB_TAL_WS_CLEAN) wiped by the previous step's tal_finalize_kernel
    unsigned int *ticket;               // [0] finalize ticket, [1] / [7] / [8] next unit of tal_gt_kernel's three kinds of work, [9] units resolved, [2] GT rows with a class id outside [0, nc), [3] grid hint rejected, [4] a peer's entry never arrived, [10] a wait on img_done timed out
    unsigned long long *stat_acc;       // [kTalStatAcc] fixed-point sums of the target scores, [kTalStatAcc] foreground counts, [1] total #foreground
    unsigned int *img_done;             // [N]  tal_decode_kernel CTAs that have finished with the image (tal_gt_kernel starts on an image behind this count)
    // armed (zeroed) by tal_decode_kernel, every call
    unsigned long long *akey;           // [N * A]  (overlap bits << 32) | ~gt_local   (0 = nobody)
    int *aslot;                         // [N * A]  1 + (g * topk + r) of the GT slot that owns the anchor (0 = background)
    // plain scratch
    float4 *dbox;                       // [N * A]  decoded pixel box (xyxy) of every anchor
    float4 *gext;                       // [ceil(A / 32)]  extent of the anchor centres of each group of 32 anchors
    float2 *ctr;                        // [A]  anchor centres in pixels
    float *peer_tss;                    // [2]  normaliser and #foreground averaged over the ranks (peer exchange, csrc/peer.cu)
    float4 *sel;                        // [gt_total * kTalMaxK]  anchor bits, metric, overlap, -
    int *sel_count;                     // [gt_total]  (image << 8) | number of selected anchors; -1 until the GT's selection is published
    // per slot = (GT, r-th selected anchor), written by tal_gt_kernel PER UNIT of the target score t:
    float4 *fterm;                      // [gt_total * kTalMaxK]  1 - CIoU, DFL term, class logit of the GT's class, its sigmoid
    float *fgrad;                       // [gt_total * kTalMaxK * 64] gradient of the anchor's 64 box logits
    long long *fcell_off;               // [gt_total * kTalMaxK] element offset of the anchor's positive class cell
    float *tsc;                         // [gt_total * kTalMaxK] target score t of the slot, < 0: the GT did not keep the anchor (tal_gt_resolve)
    float *part;                        // [N * cls_tiles]
    double *cta_sums;                   // [4 * finalize CTAs]
    int cls_tiles;
    size_t small_zero_bytes, zero_bytes, total_bytes;
};

static int tal_tile(int dtype, bool vec) { return kTalThreads * (vec ? (dtype == YB_BF16 ? 8 : 4) : 1); }

static TalWorkspace carve_tal(void *base, int n_images, int n_anchors, int gt_total, int topk, int tile) {
    TalWorkspace w;
    char *p = static_cast<char *>(base);
    const size_t g = (size_t)(gt_total > 0 ? gt_total : 1);
    const int tiles = (n_anchors + tile - 1) / tile;
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned int *>(p + off);
    off += 64;
    w.stat_acc = reinterpret_cast<unsigned long long *>(p + off);
    off += round_up(sizeof(unsigned long long) * (2 * kTalStatAcc + 1), 64);
    w.img_done = reinterpret_cast<unsigned int *>(p + off);
    off += round_up(sizeof(unsigned int) * (size_t)n_images, 64);
    w.small_zero_bytes = off;
    w.akey = reinterpret_cast<unsigned long long *>(p + off);
    off += round_up(sizeof(unsigned long long) * (size_t)n_images * n_anchors, 64);
    w.aslot = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images * n_anchors, 64);
    w.zero_bytes = off;
    (void)topk;
    w.cls_tiles = tiles * tal_cls_split(tile);             // partial sums per image
    w.dbox = reinterpret_cast<float4 *>(p + off);
    off += round_up(sizeof(float4) * (size_t)n_images * n_anchors, 64);
    w.gext = reinterpret_cast<float4 *>(p + off);
    off += round_up(sizeof(float4) * (size_t)((n_anchors + 31) / 32), 64);
    w.ctr = reinterpret_cast<float2 *>(p + off);
    off += round_up(sizeof(float2) * (size_t)n_anchors, 64);
    w.peer_tss = reinterpret_cast<float *>(p + off);
    off += 64;
    w.sel = reinterpret_cast<float4 *>(p + off);
    off += round_up(sizeof(float4) * g * kTalMaxK, 64);
    w.sel_count = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * g, 64);
    w.fterm = reinterpret_cast<float4 *>(p + off);
    off += round_up(sizeof(float4) * g * kTalMaxK, 64);
    w.fgrad = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * g * kTalMaxK * 4 * kRegMax, 64);
    w.fcell_off = reinterpret_cast<long long *>(p + off);
    off += round_up(sizeof(long long) * g * kTalMaxK, 64);
    w.tsc = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * g * kTalMaxK, 64);
    w.part = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * (size_t)n_images * w.cls_tiles, 64);
    w.cta_sums = reinterpret_cast<double *>(p + off);
    off += round_up(sizeof(double) * 4 * (((size_t)n_images * w.cls_tiles + g * kTalMaxK) / kTalFinThreadsDecl + 2), 64);
    w.total_bytes = off;
    return w;
}

// Complete-IoU of a GT box g and a predicted box p (xyxy), spec: oracle/tal_oracle.py::ciou
struct Ciou {
    float value, iou, v, alpha, inter, uni, c2, rho2, cw, ch, w1, h1, iw_raw, ih_raw, dxs, dys, at;
};
__device__ __forceinline__ float gt_atan(const float4 &g) { return atanf((g.z - g.x) / (g.w - g.y + kEpsCiou)); }

__device__ __forceinline__ Ciou ciou_eval(const float4 &p, const float4 &g, float atan_g) {
    Ciou r;
    r.w1 = p.z - p.x;
    r.h1 = p.w - p.y + kEpsCiou;
    const float w2 = g.z - g.x, h2 = g.w - g.y + kEpsCiou;
    r.iw_raw = fminf(p.z, g.z) - fmaxf(p.x, g.x);
    r.ih_raw = fminf(p.w, g.w) - fmaxf(p.y, g.y);
    r.inter = fmaxf(r.iw_raw, 0.f) * fmaxf(r.ih_raw, 0.f);
    r.uni = r.w1 * r.h1 + w2 * h2 - r.inter + kEpsCiou;
    r.iou = r.inter / r.uni;
    r.cw = fmaxf(p.z, g.z) - fminf(p.x, g.x);
    r.ch = fmaxf(p.w, g.w) - fminf(p.y, g.y);
    r.c2 = r.cw * r.cw + r.ch * r.ch + kEpsCiou;
    r.dxs = g.x + g.z - p.x - p.z;
    r.dys = g.y + g.w - p.y - p.w;
    r.rho2 = (r.dxs * r.dxs + r.dys * r.dys) * 0.25f;
    r.at = atan_g - atanf(r.w1 / r.h1);
    r.v = kFourOverPi2 * r.at * r.at;
    r.alpha = r.v / (r.v - r.iou + (1.f + kEpsCiou));
    r.value = r.iou - (r.rho2 / r.c2 + r.v * r.alpha);
    return r;
}

// ------------------------------------------------------------------------------------------
// approximate-reciprocal versions for RANKING the candidates (tal_candidates_kernel): the overlap and the
// metric agree with ciou_eval / the oracle to a few 1e-7 absolute, so only numerical near-ties of the
// metric can rank differently; the loss itself is evaluated with ciou_eval on the selected anchors.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_div(float a, float b) { return a * fast_rcp(b); }

// atan(x), x >= 0 (Cephes atanf: two range reductions + a degree-4 polynomial in x^2, ~2 ulp)
__device__ __forceinline__ float fast_atan_pos(float x) {
    const bool hi = x > 2.414213562373095f, mid = x > 0.4142135623730950f;
    const float y0 = hi ? 1.5707963267948966f : (mid ? 0.7853981633974483f : 0.f);
    const float num = hi ? -1.f : (mid ? x - 1.f : x);
    const float den = hi ? x : (mid ? x + 1.f : 1.f);
    const float r = fast_div(num, den);
    const float z = r * r;
    float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    return y0 + fmaf(p * z, r, r);
}
__device__ __forceinline__ float gt_atan_fast(const float4 &g) { return fast_atan_pos(fast_div(g.z - g.x, g.w - g.y + kEpsCiou)); }

// plain IoU of predicted box p and GT box g, the very value overlap_fast starts from (0 when they do not intersect)
__device__ __forceinline__ float iou_fast(const float4 &p, const float4 &g, float area_g) {
    const float w1 = p.z - p.x, h1 = p.w - p.y + kEpsCiou;
    const float iw = fmaxf(fminf(p.z, g.z) - fmaxf(p.x, g.x), 0.f), ih = fmaxf(fminf(p.w, g.w) - fmaxf(p.y, g.y), 0.f);
    const float inter = iw * ih;
    if (inter <= 0.f) return 0.f;
    return fast_div(inter, w1 * h1 + area_g - inter + kEpsCiou);
}

// overlap = max(CIoU, 0) of predicted box p and GT box g
__device__ __forceinline__ float overlap_fast(const float4 &p, const float4 &g, float atan_g, float area_g) {
    const float w1 = p.z - p.x, h1 = p.w - p.y + kEpsCiou;
    const float iou = iou_fast(p, g, area_g);
    if (iou <= 0.f) return 0.f;                            // IoU 0: the penalties can only push the CIoU below zero
    const float cw = fmaxf(p.z, g.z) - fminf(p.x, g.x), ch = fmaxf(p.w, g.w) - fminf(p.y, g.y);
    const float c2 = cw * cw + ch * ch + kEpsCiou;
    const float dxs = g.x + g.z - p.x - p.z, dys = g.y + g.w - p.y - p.w;
    const float rho2 = (dxs * dxs + dys * dys) * 0.25f;
    const float at = atan_g - fast_atan_pos(fast_div(w1, h1));
    const float v = kFourOverPi2 * at * at;
    const float alpha = fast_div(v, v - iou + (1.f + kEpsCiou));
    return fmaxf(iou - (fast_div(rho2, c2) + v * alpha), 0.f);
}

__device__ __forceinline__ float metric_fast(float logit, float ov, float alpha, float beta) {
    const float sc = fast_rcp(1.f + fast_ex2(-1.4426950408889634f * logit));
    if (alpha == 0.5f && beta == 6.f) {                    // the defaults: sqrt and three multiplications
        const float o2 = ov * ov;
        float rs;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(sc));
        return rs * (o2 * o2 * o2);
    }
    return powf(sc, alpha) * powf(ov, beta);
}

// An upper bound of metric_fast over every class score, from the plain IoU alone: score^alpha <= 1 and the CIoU
// penalties only lower the overlap (alpha, beta >= 0; the caller does not filter otherwise).  The margin covers the
// approximate square root of a score that rounds to 1.
__device__ __forceinline__ float metric_bound(float iou, float beta) {
    if (beta == 6.f) {
        const float o2 = iou * iou;
        return (o2 * o2 * o2) * 1.0001f;
    }
    return powf(iou, beta) * 1.0001f;
}

// ------------------------------------------------------------------------------------------
// Grid hint.  Do the anchors form the reference's pyramid of regular grids (src/utils/model_utils.py:60-70: per level
// x fastest, x = x0 + col, y = y0 + row, one stride per level)?  Anchors are an INPUT of the loss (they may be
// bf16-rounded, SURVEY Q13, or anything else), so nothing is assumed: the caller may pass the structure it believes in
// (yb_tal_grid, a host struct handed to the kernels by value), and tal_decode_kernel VERIFIES it against the anchor /
// stride arrays element by element, bit for bit, every call.  When it holds, tal_topk_kernel enumerates the anchors
// inside a GT as one rectangle of cells per level instead of scanning groups of anchors; when it does not (or no hint
// is given) the generic scan runs and out_stats[2] reports the rejection.
// ------------------------------------------------------------------------------------------
typedef yb_tal_grid TalGrid;
constexpr int kGridMaxLevels = YB_TAL_MAX_LEVELS;

__device__ __forceinline__ bool grid_matches(const TalGrid &gr, int a, float ax, float ay, float s) {
    int l = 0;
    while (l + 1 < gr.n_levels && a >= gr.start[l + 1]) ++l;
    const int j = a - gr.start[l], row = j / gr.w[l], col = j - row * gr.w[l];
    return row < gr.h[l] && ax == gr.x0[l] + (float)col && ay == gr.y0[l] + (float)row && s == gr.stride[l];
}

// ------------------------------------------------------------------------------------------
// tal_decode_kernel: the streaming half of the assignment.  One thread per VW consecutive anchors reads the
// 4 x 16 box rows (128-bit loads, 8 rows in flight), and leaves the decoded pixel box of every anchor in the
// workspace (16 B per anchor: 17 MB at cfg2, it stays in L2 for tal_gt_kernel; loading the box rows with an L2
// evict-first policy to protect it changed nothing, 171.2 vs 171.3 us for the assign phase).  The CTAs of image 0 also
// write the anchor CENTRES in pixels and their extent per group of 32 consecutive anchors (the same for all images).
// ------------------------------------------------------------------------------------------
template <typename T, int VW>
__device__ __forceinline__ void tal_decode_body(int n, int tile, const T *__restrict__ preds, int n_ch, int n_anchors,
                                                const float *__restrict__ anchors, const float *__restrict__ strides,
                                                const int *__restrict__ gt_off, float4 *dbox, float4 *gext, float2 *ctr,
                                                unsigned long long *akey, int *aslot, int *sel_count, const TalGrid &grid,
                                                unsigned int *grid_rejected) {
    constexpr int GL = 32 / VW;                            // lanes that share one group of 32 anchors
    const int a0 = (tile * kTalThreads + threadIdx.x) * VW;
    // arm the "selection published" words of the image's GTs (the image's tiles share the job)
    for (int m = tile * kTalThreads + threadIdx.x, m_img = gt_off[n + 1] - gt_off[n]; m < m_img; m += gridDim.x * kTalThreads)
        sel_count[gt_off[n] + m] = -1;
    // arm the per-anchor conflict keys and the anchor -> slot map of this thread's anchors (all images: the dense
    // pass reads the map everywhere)
#pragma unroll
    for (int v = 0; v < VW; ++v)
        if (a0 + v < n_anchors) { akey[(size_t)n * n_anchors + a0 + v] = 0ull; aslot[(size_t)n * n_anchors + a0 + v] = 0; }
    const bool want_ext = n == 0;                          // uniform per CTA
    if (gt_off[n + 1] == gt_off[n] && !want_ext) return;   // an image without GT has no candidates (uniform per CTA)
    const bool have_gt = gt_off[n + 1] != gt_off[n];
    const size_t img = (size_t)n * n_ch * n_anchors;
    float lo_x = __int_as_float(0x7f800000), lo_y = lo_x, hi_x = -lo_x, hi_y = -lo_x;
    if (a0 < n_anchors) {
        float dist[4][VW];
        if (have_gt) {
#pragma unroll
            for (int side = 0; side < 4; ++side) {
                DflPartial part[VW];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    Group<T, VW> row[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) row[j].load(preds + img + (size_t)(side * kRegMax + h * 8 + j) * n_anchors + a0);
#pragma unroll
                    for (int v = 0; v < VW; ++v) {
                        float x[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = row[j].get(v);
                        const DflPartial ph = dfl_half8(x, h * 8);
                        if (h == 0) part[v] = ph;
                        else dist[side][v] = dfl_merge(part[v], ph);
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            const float ax = __ldg(anchors + a0 + v), ay = __ldg(anchors + n_anchors + a0 + v), s = __ldg(strides + a0 + v);
            if (have_gt) {
                const PredBox b = decode_box(ax, ay, s, dist[0][v], dist[1][v], dist[2][v], dist[3][v]);
                dbox[(size_t)n * n_anchors + a0 + v] = make_float4(b.x1, b.y1, b.x2, b.y2);
            }
            const float cx = ax * s, cy = ay * s;
            if (want_ext) {
                ctr[a0 + v] = make_float2(cx, cy);
                if (grid.n_levels > 0 && !grid_matches(grid, a0 + v, ax, ay, s)) atomicOr(grid_rejected, 1u);
            }
            lo_x = fminf(lo_x, cx); hi_x = fmaxf(hi_x, cx);
            lo_y = fminf(lo_y, cy); hi_y = fmaxf(hi_y, cy);
        }
    }
    if (want_ext) {
#pragma unroll
        for (int o = 1; o < GL; o <<= 1) {
            lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
            lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
            hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
            hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
        }
        if ((threadIdx.x & (GL - 1)) == 0 && a0 < n_anchors) gext[a0 >> 5] = make_float4(lo_x, lo_y, hi_x, hi_y);
    }
}

template <typename T, int VW>
__global__ void __launch_bounds__(kTalThreads, YB_TAL_DECODE_MINBLOCKS ? YB_TAL_DECODE_MINBLOCKS : (VW == 8 ? 4 : 6))
tal_decode_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, const float *__restrict__ anchors,
                  const float *__restrict__ strides, const int *__restrict__ gt_off, float4 *__restrict__ dbox,
                  float4 *__restrict__ gext, float2 *__restrict__ ctr, unsigned long long *__restrict__ akey,
                  int *__restrict__ aslot, int *__restrict__ sel_count, const TalGrid grid,
                  unsigned int *__restrict__ grid_rejected, unsigned int *__restrict__ img_done) {
#if YB_TAL_PDL
    pdl_wait();                                            // behind the previous step's tal_finalize_kernel (YB_TAL_WS_CLEAN): its wiped counters,
                                                           // and the workspace arrays that step's kernels still read
    // ... and only then may tal_gt_kernel's CTAs move in: they do NOT wait for this grid to complete but for the per-image
    // counts below, so nothing of the previous step may still be in flight when the first of them starts.  By the time
    // every CTA of this grid has passed this line, every CTA of this grid is resident or done: a waiting successor can
    // never keep a producer out of the machine.
    pdl_launch_dependents();
#endif
    tal_decode_body<T, VW>(blockIdx.y, blockIdx.x, preds, n_ch, n_anchors, anchors, strides, gt_off, dbox, gext, ctr, akey, aslot,
                           sel_count, grid, grid_rejected);
    __syncthreads();
    if (threadIdx.x == 0) dep_signal(img_done + blockIdx.y);
}

// ------------------------------------------------------------------------------------------
// tal_gt_kernel: everything that is per GT.  GTs are handed out to WARPS from a global counter (no wave quantisation, no
// idle warps inside a CTA).
// (1) The GT's candidates are the anchors whose centre lies strictly inside it; in ascending anchor order they are
// evaluated 64 at a time (decoded boxes from L2, class logit from HBM; approximate-reciprocal arithmetic: the metric only
// RANKS) against a running top-k in registers, lane r holding the r-th best (metric desc, anchor asc): entries that beat
// the current k-th best are merged in, a few by insertion (ballot + shuffle-up), many at once by REDUX rounds.  A later
// entry that only TIES the k-th best can never displace it (ties -> lowest anchor), so the strict test is exact.
// Conflicts: 64-bit atomicMax (overlap, ~gt) per anchor.
// (2) The foreground terms of the selected anchors, one HALF-warp per anchor (lane & 15 is the DFL bin and the lane holds
// that bin of all four sides, so the scalar part -- CIoU, its gradient, the class cell -- is issued once for two
// anchors), while it is still open whether the GT keeps the anchor and what its target score t will be: CIoU loss, DFL
// loss and the gradient of the 64 box logits are all proportional to t, so they are stored per unit of t and
// tal_cls_kernel / tal_finalize_kernel apply t / normaliser.  (The anchors a GT loses to another GT -- a few per cent --
// are computed for nothing.)  As a kernel of its own this half was bound by its scattered 64-byte DRAM accesses (60 us
// after a 73 us ranking kernel bound by instruction issue); in one kernel the two overlap across warps.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int gt_image(const int *__restrict__ gt_off, int n_images, int g) {
    int lo = 0, hi = n_images;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(gt_off + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int kTopkWarps = 4;
constexpr int kTopkQueue = 128;                  // inside-anchor queue per warp: evaluated 64 at a time (< 64 waiting + 64 new)
constexpr int kEmptyKey = (int)0x80000000;       // below every metric bit pattern (metrics are >= 0)

struct TopK {                                    // lane r: the r-th best so far (r < topk), or empty
    int m, a;
    float o;
};

// one candidate (the same values in all lanes) into the sorted winners
__device__ __forceinline__ void topk_insert(TopK &w, int lane, int topk, int m, int a, float o) {
    const int p = __popc(__ballot_sync(0xffffffffu, lane < topk && w.m >= m));   // winners that stay ahead (ties: earlier anchor)
    const int um = __shfl_up_sync(0xffffffffu, w.m, 1), ua = __shfl_up_sync(0xffffffffu, w.a, 1);
    const float uo = __shfl_up_sync(0xffffffffu, w.o, 1);
    if (lane > p) { w.m = um; w.a = ua; w.o = uo; }
    if (lane == p) { w.m = m; w.a = a; w.o = o; }
}

// the k best of NE entries per lane (metric bits, anchor, overlap; kEmptyKey = none), by REDUX rounds: lane r <- r-th best
template <int NE>
__device__ __forceinline__ TopK topk_select(int (&vm)[NE], const int (&va)[NE], const float (&vo)[NE], int lane, int topk) {
    TopK nw = {kEmptyKey, 0x7fffffff, 0.f};
    for (int r = 0; r < topk; ++r) {
        int bm = vm[0], ba = va[0], bi = 0;
#pragma unroll
        for (int i = 1; i < NE; ++i)
            if (vm[i] > bm || (vm[i] == bm && va[i] < ba)) { bm = vm[i]; ba = va[i]; bi = i; }
        const int wm = __reduce_max_sync(0xffffffffu, bm);
        if (wm == kEmptyKey) break;                        // fewer than k entries in all (warp-uniform)
        const int wa = __reduce_min_sync(0xffffffffu, bm == wm ? ba : 0x7fffffff);
        const bool mine = bm == wm && ba == wa;            // exactly one lane: an anchor appears once
        const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
        float bo = vo[0];
#pragma unroll
        for (int i = 1; i < NE; ++i) bo = i == bi ? vo[i] : bo;
        const float wo = __shfl_sync(0xffffffffu, bo, src);
        if (mine) {
#pragma unroll
            for (int i = 0; i < NE; ++i) vm[i] = i == bi ? kEmptyKey : vm[i];
        }
        if (lane == r) { nw.m = wm; nw.a = wa; nw.o = wo; }
    }
    return nw;
}

// what a half-warp fetches for one selected anchor: this lane's bin of the four sides and the class logit
struct FgFetch {
    float z[4], z_cls;
};
template <typename T>
__device__ __forceinline__ FgFetch fg_fetch(const T *__restrict__ img, const T *__restrict__ cls_row, int n_anchors, int idx,
                                            int bin) {
    FgFetch f;
#pragma unroll
    for (int k = 0; k < 4; ++k) f.z[k] = load_as_float(img + (size_t)(k * kRegMax + bin) * n_anchors + idx);
    f.z_cls = load_as_float(cls_row + idx);
    return f;
}

#ifndef YB_TAL_FILTER                 // candidate filter: 0 off, 1 IoU-only bound, 2 IoU and the anchor's own class score
#define YB_TAL_FILTER 2
#endif
#ifndef YB_TAL_FENCE_ALL
#define YB_TAL_FENCE_ALL 0
#endif
#ifndef YB_TAL_SPIN_NS                // a waiting warp sleeps between polls: it must not take issue slots from the working ones
#define YB_TAL_SPIN_NS 100
#endif
#ifndef YB_TOPK_MINBLOCKS
#define YB_TOPK_MINBLOCKS 6
#endif
// what the per-GT work reads and writes (one struct so that the stand-alone kernel and the fused launch share the body)
template <typename T>
struct TalGtArgs {
    const T *preds;
    int n_images, n_ch, n_anchors;
    const float *anchors, *strides, *gt;
    const int *gt_off;
    int gt_total, topk;
    float alpha, beta, lambda_box, lambda_dfl;
    const float4 *dbox, *gext;
    const float2 *ctr;
    TalGrid grid;
    float4 *sel;
    int *sel_count;
    unsigned long long *akey;
    float4 *fterm;
    float *fgrad;
    long long *fcell_off;
    unsigned int *bad_cls;
};

// GT g (image n, g_local-th box of it) by one warp; aq = the warp's queue of kTopkQueue ints in shared memory
template <typename T>
__device__ __forceinline__ void tal_gt_body(const TalGtArgs<T> &A, int g, int n, int g_local, bool regular, int *aq) {
    const int lane = threadIdx.x & 31;
    const T *__restrict__ preds = A.preds;
    const float *__restrict__ anchors = A.anchors, *__restrict__ strides = A.strides, *__restrict__ gt = A.gt;
    const int n_ch = A.n_ch, n_anchors = A.n_anchors, topk = A.topk;
    const float alpha = A.alpha, beta = A.beta, lambda_box = A.lambda_box, lambda_dfl = A.lambda_dfl;
    const float4 *dbox = A.dbox, *gext = A.gext;
    const float2 *ctr = A.ctr;
    const TalGrid &grid = A.grid;
    float4 *__restrict__ sel = A.sel;
    int *__restrict__ sel_count = A.sel_count;
    unsigned long long *akey = A.akey;
    float4 *__restrict__ fterm = A.fterm;
    float *__restrict__ fgrad = A.fgrad;
    long long *__restrict__ fcell_off = A.fcell_off;
    unsigned int *bad_cls = A.bad_cls;
    const int n_cls = n_ch - 4 * kRegMax;
    const int n_groups = (n_anchors + 31) >> 5;
    const int n_levels = grid.n_levels;
    {
        const float *g5 = gt + (size_t)g * 5;
        const float gcx = __ldg(g5), gcy = __ldg(g5 + 1), gw = __ldg(g5 + 2), gh = __ldg(g5 + 3);
        const float4 gb = make_float4(gcx - gw * 0.5f, gcy - gh * 0.5f, gcx + gw * 0.5f, gcy + gh * 0.5f);
        const int cls_raw = (int)__ldg(g5 + 4);
        const int cls = min(max(cls_raw, 0), n_cls - 1);
        if (lane == 0 && cls_raw != cls) atomicAdd(bad_cls, 1u);   // clamped (memory-safe), counted: the caller raises
        const float at_g = gt_atan_fast(gb);
        const float area_g = (gb.z - gb.x) * (gb.w - gb.y + kEpsCiou);
        const T *img = preds + (size_t)n * n_ch * n_anchors;
        const T *cls_row = img + (size_t)(4 * kRegMax + cls) * n_anchors;
        const float4 *box_row = dbox + (size_t)n * n_anchors;

        TopK win = {kEmptyKey, 0x7fffffff, 0.f};
        int thr = -1;                                      // bit pattern of the k-th best metric once k winners exist
        int nq = 0;                                        // queue fill (warp-uniform)

        int thr_seed = -1;                                 // k-th best metric of the seed cells (below), once known
        // evaluate up to 64 anchors, two per lane (a < 0: none), in ascending order lane-major within each of the two
        auto evaluate2 = [&](const int (&a)[2]) {
            float4 pb[2];
            float lg[2];
#pragma unroll
            for (int u = 0; u < 2; ++u)                    // all four loads in flight together
                if (a[u] >= 0) { pb[u] = box_row[a[u]]; lg[u] = load_as_float(cls_row + a[u]); }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float ov = 0.f;
                int m = kEmptyKey;
                if (a[u] >= 0) {
                    ov = overlap_fast(pb[u], gb, at_g, area_g);
                    m = __float_as_int(metric_fast(lg[u], ov, alpha, beta));
                }
                // (an exact metric strictly below the seed's k-th best cannot be selected either: it never enters the list,
                // so the list is mostly built by cheap insertions instead of REDUX rounds)
                unsigned keep = __ballot_sync(0xffffffffu, (m > thr) & (m >= thr_seed));
                if (keep == 0u) continue;                  // warp-uniform
                if (__popc(keep) > 6) {
                    int vm[2] = {lane < topk ? win.m : kEmptyKey, ((keep >> lane) & 1u) ? m : kEmptyKey};
                    const int va[2] = {win.a, a[u]};
                    const float vo[2] = {win.o, ov};
                    win = topk_select<2>(vm, va, vo, lane, topk);
                    const int kth = __shfl_sync(0xffffffffu, win.m, topk - 1);
                    thr = kth == kEmptyKey ? -1 : kth;
                } else {
                    while (keep) {                         // ascending lane = ascending anchor
                        const int src = __ffs(keep) - 1;
                        keep &= keep - 1;
                        const int sm = __shfl_sync(0xffffffffu, m, src);
                        const int sa = __shfl_sync(0xffffffffu, a[u], src);
                        const float so = __shfl_sync(0xffffffffu, ov, src);
                        if (sm > thr) {                    // thr may have risen since the ballot (warp-uniform)
                            topk_insert(win, lane, topk, sm, sa, so);
                            const int kth = __shfl_sync(0xffffffffu, win.m, topk - 1);
                            thr = kth == kEmptyKey ? -1 : kth;
                        }
                    }
                }
            }
        };

        // (Counted on bench.py's input: 308 inside anchors per GT, 138 pass the filter, 38 enter the list, 2 REDUX merges.)
        // Cheap filter in front of the evaluation: a candidate whose plain IoU already bounds its metric at or below the
        // k-th best so far (or strictly below a k-th best known from a seed of central cells) can never be selected; only the
        // survivors, compacted into the queue in ascending order, pay for the CIoU, the class logit and the merge.
        const bool can_bound = YB_TAL_FILTER && beta >= 0.f;
        auto survives = [&](const float4 &pb, float logit) -> bool {     // the candidate's decoded box and class logit
#if YB_TAL_FILTER == 2
            // the metric itself with the plain IoU in place of the overlap: the class score is the anchor's own (the head's
            // scores sit around 0.01, a bound of 1 would be ten times too loose), the CIoU penalties only lower the overlap
            // and the metric does not fall as the overlap grows (beta >= 0)
            const int ub = __float_as_int(metric_fast(logit, iou_fast(pb, gb, area_g), alpha, beta) * 1.0001f);
#else
            const int ub = __float_as_int(metric_bound(iou_fast(pb, gb, area_g), beta));
#endif
            return (ub > thr) & (ub >= thr_seed);
        };

        // lane l works out the rectangle of level l (the others fetch it by shuffle)
        int r_x0 = 0, r_y0 = 0, r_nx = 0, r_ny = 0;
        if (regular) {
            // ---- the anchors form regular grids: the inside anchors of a level are a rectangle of cells --------
            if (lane < n_levels) {
                const int W = grid.w[lane], H = grid.h[lane];
                const float s = grid.stride[lane], x0 = grid.x0[lane], y0 = grid.y0[lane];
                // first / last column and row whose centre is strictly inside: a guess from the division, made exact
                // with the very comparison the generic path applies to the stored centres ((x0 + col) * s is the
                // stored value bit for bit: tal_decode_kernel verified that)
                auto first_in = [&](float lo, float c0, int n) {
                    int i = min(max((int)floorf(lo / s - c0), 0), n);
                    while (i > 0 && (c0 + (float)(i - 1)) * s - lo > kEpsIn) --i;
                    while (i < n && !((c0 + (float)i) * s - lo > kEpsIn)) ++i;
                    return i;
                };
                auto last_in = [&](float hi, float c0, int n) {
                    int i = min(max((int)ceilf(hi / s - c0), -1), n - 1);
                    while (i < n - 1 && hi - (c0 + (float)(i + 1)) * s > kEpsIn) ++i;
                    while (i >= 0 && !(hi - (c0 + (float)i) * s > kEpsIn)) --i;
                    return i;
                };
                r_x0 = first_in(gb.x, x0, W);
                r_y0 = first_in(gb.y, y0, H);
                r_nx = max(last_in(gb.z, x0, W) - r_x0 + 1, 0);
                r_ny = max(last_in(gb.w, y0, H) - r_y0 + 1, 0);
            }
            __syncwarp();
            if (can_bound) {
                // seed: the 3 x 3 cells around the GT's centre on (up to) three levels, one per lane; the k-th best metric
                // among them is a lower bound of the final k-th best.  Only used to DROP candidates strictly below it: the
                // ranking itself starts empty and still sees these cells, in their place in the anchor order.
                const int sl = min(lane / 9, n_levels - 1), k9 = lane % 9;
                const int sx0 = __shfl_sync(0xffffffffu, r_x0, sl), sy0 = __shfl_sync(0xffffffffu, r_y0, sl);
                const int snx = __shfl_sync(0xffffffffu, r_nx, sl), sny = __shfl_sync(0xffffffffu, r_ny, sl);
                int v = kEmptyKey;
                if (lane < 27 && lane / 9 < n_levels && snx > 0 && sny > 0) {
                    const float s = grid.stride[sl];
                    const int cc = min(max((int)floorf(gcx / s - grid.x0[sl] + 0.5f), sx0), sx0 + snx - 1) + (k9 % 3) - 1;
                    const int cr = min(max((int)floorf(gcy / s - grid.y0[sl] + 0.5f), sy0), sy0 + sny - 1) + (k9 / 3) - 1;
                    if (cc >= sx0 && cc < sx0 + snx && cr >= sy0 && cr < sy0 + sny) {
                        const int a = grid.start[sl] + cr * grid.w[sl] + cc;
                        v = __float_as_int(metric_fast(load_as_float(cls_row + a), overlap_fast(box_row[a], gb, at_g, area_g), alpha, beta));
                    }
                }
                int kth = kEmptyKey;
                for (int r = 0; r < topk; ++r) {           // k-th largest of the lanes' values
                    kth = __reduce_max_sync(0xffffffffu, v);
                    if (kth == kEmptyKey) break;           // fewer than k seeds
                    const unsigned eq = __ballot_sync(0xffffffffu, v == kth);
                    if (lane == __ffs(eq) - 1) v = kEmptyKey;
                }
                // (a hair below the k-th best seed: the seeds are evaluated by another copy of the metric arithmetic than the
                // candidates, and the last bit of the two may differ -- the seed itself must never fall below its own mark)
                thr_seed = kth == kEmptyKey ? -1 : __float_as_int(__int_as_float(kth) * 0.9999f);
            }
        }
        // One loop for both forms of the enumeration, so that the evaluation below exists ONCE in the kernel (one copy of
        // its arithmetic: the two forms give bit-identical metrics; and a third of the code).  Each round yields up to 32
        // candidates in ascending anchor order -- the next cells of the current level's rectangle, or the next group of 32
        // anchors whose centre extent meets the GT -- which pass the cheap filter into the queue.
        int lvl = -1, c0 = 0, cells = 0, ix0 = 0, iy0 = 0, nx = 1, lw = 0, lst = 0;        // grid form
        float inv_nx = 1.f;
        int gb0 = -32;                                                                      // group form
        unsigned groups = 0u;
        for (bool more = true; more;) {
            // up to two rounds of 32 candidates per trip (grid form), so that two boxes and two class logits per lane are
            // in flight together: the filter is a chain of dependent L2 / DRAM round trips otherwise (10 us per GT)
            int a[2] = {-1, -1};
            bool in[2] = {false, false};
            if (regular) {
                while (c0 >= cells) {                      // warp-uniform: on to the next level with cells inside the GT
                    if (++lvl >= n_levels) { more = false; break; }
                    ix0 = __shfl_sync(0xffffffffu, r_x0, lvl); iy0 = __shfl_sync(0xffffffffu, r_y0, lvl);
                    nx = __shfl_sync(0xffffffffu, r_nx, lvl);
                    cells = nx * __shfl_sync(0xffffffffu, r_ny, lvl);
                    lw = grid.w[lvl]; lst = grid.start[lvl];
                    inv_nx = 1.f / (float)max(nx, 1);
                    c0 = 0;
                }
                if (more) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c = c0 + 32 * u + lane;
                        int row = (int)(((float)c + 0.5f) * inv_nx), col = c - row * nx;
                        if (col < 0) { --row; col += nx; } else if (col >= nx) { ++row; col -= nx; }
                        a[u] = lst + (iy0 + row) * lw + ix0 + col;
                        in[u] = c < cells;
                    }
                    c0 += 64;
                }
            } else {
                while (groups == 0u) {                     // warp-uniform: lane l looks at the centre extent of group gb0 + l
                    gb0 += 32;
                    if (gb0 >= n_groups) { more = false; break; }
                    float4 ext = make_float4(0.f, 0.f, -1.f, -1.f);   // an empty extent never intersects
                    if (gb0 + lane < n_groups) ext = gext[gb0 + lane];
                    groups = __ballot_sync(0xffffffffu, gb.x < ext.z && gb.z > ext.x && gb.y < ext.w && gb.w > ext.y);
                }
                if (more) {
                    a[0] = ((gb0 + (__ffs(groups) - 1)) << 5) + lane;
                    groups &= groups - 1;
                    if (a[0] < n_anchors) {
                        const float2 c = ctr[a[0]];
                        in[0] = fminf(fminf(c.x - gb.x, c.y - gb.y), fminf(gb.z - c.x, gb.w - c.y)) > kEpsIn;
                    }
                }
            }
            // the cheap filter on both rounds (loads first), then append the kept anchors in round / lane order
            float4 fb[2];
            float fl[2];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (in[u] && can_bound) { fb[u] = box_row[a[u]]; fl[u] = YB_TAL_FILTER == 2 ? load_as_float(cls_row + a[u]) : 0.f; }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                bool keep = in[u];
                if (keep && can_bound) keep = survives(fb[u], fl[u]);
                const unsigned mask = __ballot_sync(0xffffffffu, keep);
                if (keep) aq[nq + __popc(mask & ((1u << lane) - 1u))] = a[u];
                nq += __popc(mask);
            }
            __syncwarp();
            // evaluate 64 at a time (and what is left once the enumeration ends)
            while (nq >= 64 || (!more && nq > 0)) {        // warp-uniform
                const int cnt = min(nq, 64);
                const int a2[2] = {lane < cnt ? aq[lane] : -1, lane + 32 < cnt ? aq[lane + 32] : -1};
                evaluate2(a2);
                const int rest = nq - cnt;                 // < 64
                const int k0 = lane < rest ? aq[64 + lane] : 0, k1 = lane + 32 < rest ? aq[96 + lane] : 0;
                __syncwarp();
                if (lane < rest) aq[lane] = k0;
                if (lane + 32 < rest) aq[32 + lane] = k1;
                nq = rest;
                __syncwarp();
            }
        }
        __syncwarp();                                      // the queue is reused by the warp's next GT

        // ---- publish the GT's list; conflicts: the anchor goes to the GT with the largest overlap, ties -> lowest GT ----
        const int n_sel = __popc(__ballot_sync(0xffffffffu, lane < topk && win.m != kEmptyKey));
        if (lane < n_sel) {
            sel[(size_t)g * kTalMaxK + lane] = make_float4(__int_as_float(win.a), __int_as_float(win.m), win.o, 0.f);
            atomicMax(akey + (size_t)n * n_anchors + win.a,
                      ((unsigned long long)__float_as_uint(win.o) << 32) | (unsigned int)(~(unsigned int)g_local));
        }
#if YB_TAL_FENCE_ALL
        __threadfence();                                   // the list before the word that announces it
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int *>(sel_count + g) = (n << 8) | n_sel;
#else
        // the list before the word that announces it: the lanes' stores are ordered before lane 0's fence by the warp
        // barrier, and the fence is cumulative
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            *reinterpret_cast<volatile int *>(sel_count + g) = (n << 8) | n_sel;
        }
#endif
    }
}

// (2) The foreground terms of GT g's selected anchors, per unit of the target score, by one warp -- any warp, once the GT's
// selection is published.
template <typename T>
__device__ __forceinline__ void tal_gt_terms(const TalGtArgs<T> &A, int g) {
    const int lane = threadIdx.x & 31;
    const T *__restrict__ preds = A.preds;
    const float *__restrict__ anchors = A.anchors, *__restrict__ strides = A.strides, *__restrict__ gt = A.gt;
    const int n_ch = A.n_ch, n_anchors = A.n_anchors, topk = A.topk;
    const float lambda_box = A.lambda_box, lambda_dfl = A.lambda_dfl;
    float4 *__restrict__ fterm = A.fterm;
    float *__restrict__ fgrad = A.fgrad;
    long long *__restrict__ fcell_off = A.fcell_off;
    const int n_cls = n_ch - 4 * kRegMax;
    {
        // wait for the selection (the warp that owns it drew the GT before this one did, so it is running or done)
        int packed;
        while ((packed = (int)ld_acquire_u32(reinterpret_cast<const unsigned int *>(A.sel_count + g))) < 0) __nanosleep(YB_TAL_SPIN_NS);
        const int n = packed >> 8, n_sel = packed & 0xff;
        if (n_sel == 0) return;                            // warp-uniform
        const float *g5 = gt + (size_t)g * 5;
        const float gcx = __ldg(g5), gcy = __ldg(g5 + 1), gw = __ldg(g5 + 2), gh = __ldg(g5 + 3);
        const float4 gb = make_float4(gcx - gw * 0.5f, gcy - gh * 0.5f, gcx + gw * 0.5f, gcy + gh * 0.5f);
        const int cls = min(max((int)__ldg(g5 + 4), 0), n_cls - 1);
        const T *img = preds + (size_t)n * n_ch * n_anchors;
        const T *cls_row = img + (size_t)(4 * kRegMax + cls) * n_anchors;
        const float4 *box_row = A.dbox + (size_t)n * n_anchors;
        // lane r: the r-th selected anchor
        const int my_a = __float_as_int(__ldcg(A.sel + (size_t)g * kTalMaxK + min(lane, n_sel - 1)).x);
        // (2a) what is scalar per anchor -- CIoU, its gradient w.r.t. the four distances, the DFL target of each side --
        // once, lane r for the r-th selected anchor (spec: oracle/tal_oracle.py::ciou; the exact arithmetic from here on).
        // The box is the one tal_decode_kernel left in the workspace.
        float s_dd[4], s_tk[4], s_box;
        {
            const int idx = my_a;
            const float4 pb = box_row[idx];
            const float ax = __ldg(anchors + idx), ay = __ldg(anchors + n_anchors + idx), s = __ldg(strides + idx);
            const Ciou c = ciou_eval(pb, gb, gt_atan(gb));
            auto w_gt = [](float a, float o) { return a > o ? 1.f : (a == o ? 0.5f : 0.f); };     // d max(a,o)/da
            auto w_lt = [](float a, float o) { return a < o ? 1.f : (a == o ? 0.5f : 0.f); };     // d min(a,o)/da
            const float iw = fmaxf(c.iw_raw, 0.f), ih = fmaxf(c.ih_raw, 0.f);
            const float inv_u2 = 1.f / (c.uni * c.uni);
            const float d_inter = (c.uni + c.inter) * inv_u2;                // d iou / d inter (union contains -inter)
            const float d_area1 = -c.inter * inv_u2;
            const float d_iw = c.iw_raw >= 0.f ? d_inter * ih : 0.f;
            const float d_ih = c.ih_raw >= 0.f ? d_inter * iw : 0.f;
            // iou part
            float gx1 = -d_iw * w_gt(pb.x, gb.x) - d_area1 * c.h1;
            float gx2 = d_iw * w_lt(pb.z, gb.z) + d_area1 * c.h1;
            float gy1 = -d_ih * w_gt(pb.y, gb.y) - d_area1 * c.w1;
            float gy2 = d_ih * w_lt(pb.w, gb.w) + d_area1 * c.w1;
            // - rho2 / c2
            const float inv_c2 = 1.f / c.c2;
            const float k_r = c.rho2 * inv_c2 * inv_c2;                      // rho2 / c2^2
            // d rho2/dx1 = d rho2/dx2 = -dxs/2 ;  d c2/dx2 = 2 cw [x2 > u2], d c2/dx1 = -2 cw [x1 < u1]
            gx1 -= (-0.5f * c.dxs) * inv_c2 - k_r * (-2.f * c.cw * w_lt(pb.x, gb.x));
            gx2 -= (-0.5f * c.dxs) * inv_c2 - k_r * (2.f * c.cw * w_gt(pb.z, gb.z));
            gy1 -= (-0.5f * c.dys) * inv_c2 - k_r * (-2.f * c.ch * w_lt(pb.y, gb.y));
            gy2 -= (-0.5f * c.dys) * inv_c2 - k_r * (2.f * c.ch * w_gt(pb.w, gb.w));
            // - alpha v :  v = k at^2, at = atan(w2/h2) - atan(w1/h1)
            const float dv_dA1 = -2.f * kFourOverPi2 * c.at;
            const float inv_hyp = 1.f / (c.h1 * c.h1 + c.w1 * c.w1);
            const float dv_dw1 = dv_dA1 * (c.h1 * inv_hyp), dv_dh1 = dv_dA1 * (-c.w1 * inv_hyp);
            gx1 -= c.alpha * (-dv_dw1);
            gx2 -= c.alpha * dv_dw1;
            gy1 -= c.alpha * (-dv_dh1);
            gy2 -= c.alpha * dv_dh1;
            // L_box = (1 - ciou) * t / tss * lambda_box   (t / tss applied later)
            const float kb = -lambda_box * s;
            s_dd[0] = -kb * gx1; s_dd[1] = -kb * gy1; s_dd[2] = kb * gx2; s_dd[3] = kb * gy2;   // d / d (dl, dt, dr, db)
            s_box = 1.f - c.value;
            // DFL target of each side (same rule as the reference, src/model/losses.py:226-246)
            const float inv_s = 1.f / s;
            const float hi_clamp = (float)(kRegMax - 1 - 0.01);
            s_tk[0] = fminf(fmaxf(ax - gb.x * inv_s, 0.f), hi_clamp);
            s_tk[1] = fminf(fmaxf(ay - gb.y * inv_s, 0.f), hi_clamp);
            s_tk[2] = fminf(fmaxf(gb.z * inv_s - ax, 0.f), hi_clamp);
            s_tk[3] = fminf(fmaxf(gb.w * inv_s - ay, 0.f), hi_clamp);
        }
        // (2b) what is per bin, one HALF-warp per anchor, two anchors per round
        const int half = lane >> 4, bin = lane & 15, base = lane & 16;
        FgFetch cur;
        cur = fg_fetch<T>(img, cls_row, n_anchors, __shfl_sync(0xffffffffu, my_a, min(half, n_sel - 1)), bin);
        for (int r0 = 0; r0 < n_sel; r0 += 2) {            // warp-uniform
            const int r = r0 + half;
            const bool live = r < n_sel;                   // an idle half walks through the same shuffles and writes nothing
            const int src = min(r, n_sel - 1);
            const int idx = __shfl_sync(0xffffffffu, my_a, src);
            FgFetch nxt = cur;
            if (r0 + 2 < n_sel)                            // the next round's gathers fly under this round's arithmetic
                nxt = fg_fetch<T>(img, cls_row, n_anchors, __shfl_sync(0xffffffffu, my_a, min(r + 2, n_sel - 1)), bin);
            const float kd = lambda_dfl * 0.25f;
            float dfl4 = 0.f, gk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // softmax and expectation of the side over the 16 lanes of the half (xor offsets <= 8 stay inside it)
                float m = cur.z[k];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                const float zs = cur.z[k] - m;
                const float ex = fast_ex2(zs * 1.4426950408889634f);
                float sum = ex;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float p = ex * fast_rcp(sum);
                float d = p * (float)bin;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                // DFL row (src/model/losses.py:63-78) and the gradient of this lane's logit
                const float tk = __shfl_sync(0xffffffffu, s_tk[k], src), ddk = __shfl_sync(0xffffffffu, s_dd[k], src);
                const int bl = (int)tk;
                const float wl = (float)(bl + 1) - tk, wr = tk - (float)bl;
                const float lpk = zs - __logf(sum);                          // log-softmax of this lane's bin
                dfl4 -= __shfl_sync(0xffffffffu, lpk, base + bl) * wl + __shfl_sync(0xffffffffu, lpk, base + bl + 1) * wr;
                const float oh = (bin == bl ? wl : 0.f) + (bin == bl + 1 ? wr : 0.f);
                gk[k] = kd * ((wl + wr) * p - oh) + ddk * p * ((float)bin - d);
            }
            const float u_box = __shfl_sync(0xffffffffu, s_box, src);
            if (live) {                                            // no shuffles below
                const size_t slot = (size_t)g * topk + r;
                // compact, coalesced: the dense kernel merges these 64 values (times t / normaliser) into the anchor's box rows
#pragma unroll
                for (int k = 0; k < 4; ++k) fgrad[slot * (4 * kRegMax) + k * kRegMax + bin] = gk[k];
                if (bin == 0) {
                    const float sg = __fdiv_rn(1.f, 1.f + expf(-cur.z_cls));
                    fterm[slot] = make_float4(u_box, dfl4 * 0.25f, cur.z_cls, sg);
                    fcell_off[slot] = (long long)((size_t)n * n_ch * n_anchors + (size_t)(4 * kRegMax + cls) * n_anchors + idx);
                }
            }
            cur = nxt;
        }
    }
}

// [sum of target scores, #foreground] of this rank from the fixed-point sub-accumulators, and the producer side of the
// exchange (csrc/peer.cu).  One whole WARP, after every slot has been resolved.
__device__ __forceinline__ void tal_publish_stats(unsigned long long *stat_acc, const unsigned int *grid_rejected, int have_hint,
                                                  float *__restrict__ out_stats, const yb_peer_exchange &px) {
    static_assert(kTalStatAcc == 64, "two sub-accumulators per lane");
    const int lane = threadIdx.x & 31;
    long long s_t = (long long)__ldcg(stat_acc + lane) + (long long)__ldcg(stat_acc + 32 + lane);
    long long s_n = (long long)__ldcg(stat_acc + kTalStatAcc + lane) + (long long)__ldcg(stat_acc + kTalStatAcc + 32 + lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_t += __shfl_xor_sync(0xffffffffu, s_t, o);
        s_n += __shfl_xor_sync(0xffffffffu, s_n, o);
    }
    const float t = (float)((double)s_t / kTalFix), nf = (float)s_n;
    if (lane == 0) stat_acc[2 * kTalStatAcc] = (unsigned long long)s_n;
    if (lane < 8)
        out_stats[lane] = lane == 0 ? t                                                 // local sum of target scores (un-clamped)
                        : lane == 1 ? nf                                                // foreground anchors
                        : lane == 2 ? (have_hint && __ldcg(grid_rejected) ? 1.f : 0.f)  // the grid hint did not describe the anchors
                                    : 0.f;
    // the exchange, producer side (csrc/peer.cu): lane r stores this rank's entry of step px.seq into rank r's mailbox —
    // a peer store over NVLink / NVSwitch for r != rank.  Two self-validating 8-byte words: (seq << 32) | payload.
    for (int r = lane; r < px.world; r += 32) {
        volatile unsigned long long *e = static_cast<unsigned long long *>(px.mailbox[r]) +
                                         ((size_t)(px.seq % kPeerSlotsTal) * YB_PEER_MAX_WORLD + px.rank) * 2;
        e[0] = ((unsigned long long)px.seq << 32) | __float_as_uint(t);
        e[1] = ((unsigned long long)px.seq << 32) | __float_as_uint(nf);
    }
}

// a rank without boxes still owes its peers an entry
__global__ void __launch_bounds__(32)
tal_stats_kernel(unsigned long long *__restrict__ stat_acc, const unsigned int *__restrict__ grid_rejected, int have_hint,
                 float *__restrict__ out_stats, const yb_peer_exchange px) {
    tal_publish_stats(stat_acc, grid_rejected, have_hint, out_stats, px);
}

// (3) What stays open until every GT of an image has published its selection: did the GT keep the anchor (conflicts went
// to the larger overlap: the atomicMax on akey), and from the GT's largest metric / overlap over the anchors it kept, the
// target score t of every slot (t = -1: not kept, or never filled).  One HALF-warp per GT, lane & 15 = the GT's r-th
// selected anchor; GTs 2u and 2u + 1 by one warp.  The target scores are summed in fixed point with integer atomics, so
// the statistics do not depend on the order in which warps finish.
struct TalResolveArgs {
    float *tsc;
    int *aslot;
    int *out_assigned;
    float *out_tscore;
    unsigned long long *stat_acc;
};
template <typename T>
__device__ __forceinline__ void tal_gt_resolve(const TalGtArgs<T> &A, const TalResolveArgs &R, int unit) {
    const int lane = threadIdx.x & 31, bin = lane & 15;
    const int g = 2 * unit + (lane >> 4);
    const bool in_range = g < A.gt_total;
    const int gg = in_range ? g : A.gt_total - 1;
    // the image(s) of the two GTs must be complete: every selection of theirs published (all were drawn before this unit)
    const int n_lo = gt_image(A.gt_off, A.n_images, 2 * unit), n_hi = gt_image(A.gt_off, A.n_images, min(2 * unit + 1, A.gt_total - 1));
    for (int i = __ldg(A.gt_off + n_lo) + lane, end = __ldg(A.gt_off + n_hi + 1); __any_sync(0xffffffffu, i < end); i += 32)
        if (i < end)
            while ((int)ld_acquire_u32(reinterpret_cast<const unsigned int *>(A.sel_count + i)) < 0) __nanosleep(YB_TAL_SPIN_NS);
    __syncwarp();
    const int packed = __ldcg(A.sel_count + gg);           // (image << 8) | count
    const int n = packed >> 8;
    const int g_local = gg - __ldg(A.gt_off + n);
    const int ns = in_range ? packed & 0xff : 0;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    bool pos = false;
    if (bin < ns) {
        e = __ldcg(A.sel + (size_t)gg * kTalMaxK + bin);
        const unsigned long long k = __ldcg(A.akey + (size_t)n * A.n_anchors + __float_as_int(e.x));
        pos = (unsigned int)(k & 0xffffffffull) == (unsigned int)(~(unsigned int)g_local);
    }
    float mm = pos ? e.y : 0.f, mo = pos ? e.z : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    }
    const float t = pos ? e.y * (mo / (mm + kEpsNorm)) : -1.f;
    if (in_range && bin < A.topk) {
        const int slot = gg * A.topk + bin;
        R.tsc[slot] = t;
        if (pos) {
            const int idx = __float_as_int(e.x);
            atomicAdd(R.stat_acc + (slot & (kTalStatAcc - 1)), (unsigned long long)__double2ll_rn((double)t * kTalFix));
            atomicAdd(R.stat_acc + kTalStatAcc + (slot & (kTalStatAcc - 1)), 1ull);
            R.aslot[(size_t)n * A.n_anchors + idx] = slot + 1;
            if (R.out_assigned) R.out_assigned[(size_t)n * A.n_anchors + idx] = g_local;
            if (R.out_tscore) R.out_tscore[(size_t)n * A.n_anchors + idx] = t;
        }
    }
}

// GTs are handed out to warps from a global counter (no wave quantisation, no idle warps inside a CTA).
// Measured and not kept: decode CTAs and per-GT CTAs as two roles of ONE launch behind per-image in-kernel dependencies
// (common.cuh), GT CTAs of image i placed `lag` CTAs behind the image's decode CTAs: 255 / 244 / 225 / 216 us for the
// assign phase at a lag of 100 / 300 / 1000 / all CTAs against 206 us for the two launches -- the long-lived GT CTAs take
// the resident slots the streaming role needs to keep HBM busy, and a per-image GT queue balances worse than a global one.
#ifdef YB_TAL_TRACE                   // measurement aid (scratch/trace_tal.py): when every warp of tal_gt_kernel passed its phases
__device__ ulonglong4 g_tal_trace[1 << 13];
__device__ unsigned long long g_tal_trace2[1 << 13];
__device__ __forceinline__ unsigned long long tal_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
extern "C" int yb_tal_trace_dump(void *dst, void *dst2) {
    int rc = (int)cudaMemcpyFromSymbol(dst, g_tal_trace, sizeof(g_tal_trace));
    return rc ? rc : (int)cudaMemcpyFromSymbol(dst2, g_tal_trace2, sizeof(g_tal_trace2));
}
#endif
template <typename T>
__global__ void __launch_bounds__(32 * kTopkWarps, YB_TOPK_MINBLOCKS)
tal_gt_kernel(const TalGtArgs<T> A, const TalResolveArgs R, const unsigned int *grid_rejected,
              unsigned int *__restrict__ next_gt, int have_hint, float *__restrict__ out_stats, const yb_peer_exchange px,
              const unsigned int *img_done, int decode_tiles) {
    __shared__ int s_aq[kTopkWarps][kTopkQueue];           // queue of inside anchors
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if YB_TAL_PDL
    pdl_launch_dependents();
#endif
    // No wait for tal_decode_kernel as a whole: its last wave fills only part of the machine (2 176 CTAs on 888 slots at
    // cfg2), and these CTAs move into the rest and start on the images that are already decoded -- a warp waits for the
    // decode CTAs of the image it is about to read (boxes, armed keys and "published" words), image 0 first (the anchor
    // centres, their extents and the verdict on the grid hint are written by image 0's CTAs).  The acquire of dep_wait
    // orders the warp's reads behind the producers' release; the units further down read behind the per-GT / per-unit
    // acquires that chain back to it.  griddepcontrol.wait at the END of the kernel keeps the stream's completion order
    // (tal_cls_kernel waits for THIS grid only).
#ifdef YB_TAL_TRACE
    const unsigned long long tr0 = tal_now();
#endif
    if (lane == 0) dep_wait(img_done, (unsigned int)decode_tiles, next_gt + 9);
    __syncwarp();
    const bool regular = A.grid.n_levels > 0 && __ldcg(grid_rejected) == 0u;   // uniform over the launch
    int n_ready = 0;                                       // the image this warp last waited for (GTs come in image order)
    // (Measured and not kept: the next round's boxes requested one trip ahead of the filter -- the restructured loop cost
    // more than the hidden L2 latency gained, 180 vs 174 us for the assign phase; asking for the next unit one unit ahead:
    // no change; one batch-wide "selections done" counter for the target scores: +18 us, they then start only when the
    // last selection of the batch is through.)
    // Three kinds of work units, three counters: first every GT's selection (the long kind), then every GT's foreground
    // terms (the short kind) by whichever warp comes free -- a warp that runs out of selections while others are still
    // ranking goes on with terms instead of idling --, last the target scores (tiny: the launch ends on them), whose
    // last unit publishes the rank's statistics.  (Measured and not kept: the warp that finishes an image's last GT
    // resolving that image, one lane per GT: a release fence per GT and a serial tail per image, 216 us against 185 us
    // for the assign phase; the target scores as a launch of their own: 10 us more.)
    // (Measured and not kept: the large GTs first -- a pass over the draws that only takes GTs with >= 600 anchors inside, then
    // the rest.  The slowest selection ends at 80 us instead of 89, but the second round of draws (an atomic and a dependent
    // load per skipped GT) delays every warp: selections done at 48 us instead of 42 for the median warp, kernel 103 us
    // instead of 97, step 313 vs 307 us; scratch/trace_tal.py prints the phase times per warp.)
    for (;;) {
        int g = 0;
        if (lane == 0) g = (int)atomicAdd(next_gt, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= A.gt_total) break;
        const int n = gt_image(A.gt_off, A.n_images, g);
        if (n != n_ready) {
            if (lane == 0) dep_wait(img_done + n, (unsigned int)decode_tiles, next_gt + 9);
            __syncwarp();
            n_ready = n;
        }
        tal_gt_body<T>(A, g, n, g - __ldg(A.gt_off + n), regular, s_aq[warp]);
    }
#ifdef YB_TAL_TRACE
    const unsigned long long tr1 = tal_now();
#endif
    for (;;) {
        int g = 0;
        if (lane == 0) g = (int)atomicAdd(next_gt + 6, 1u);     // ticket[7]
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= A.gt_total) break;
        tal_gt_terms<T>(A, g);
    }
#ifdef YB_TAL_TRACE
    const unsigned long long tr2 = tal_now();
#endif
    const int n_units = (A.gt_total + 1) >> 1;
    for (;;) {
        int u = 0;
        if (lane == 0) u = (int)atomicAdd(next_gt + 7, 1u);     // ticket[8]
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        tal_gt_resolve<T>(A, R, u);
        __threadfence();                                   // this unit's statistics and slots before the count below
        __syncwarp();
        unsigned int before = 0;
        if (lane == 0) before = atomicAdd(next_gt + 8, 1u);     // ticket[9]: units resolved
        before = __shfl_sync(0xffffffffu, before, 0);
        if ((int)before == n_units - 1) {                   // the launch's last unit
            __threadfence();
            tal_publish_stats(R.stat_acc, grid_rejected, have_hint, out_stats, px);
        }
    }
#ifdef YB_TAL_TRACE
    if (lane == 0 && blockIdx.x * kTopkWarps + warp < (1 << 13)) {
        g_tal_trace[blockIdx.x * kTopkWarps + warp] = make_ulonglong4(tr0, tr1, tr2, tal_now());
    }
#endif
#if YB_TAL_PDL
    pdl_wait();                                            // (returns at once: every image this grid read was complete)
#endif
#ifdef YB_TAL_TRACE
    if (lane == 0 && blockIdx.x * kTopkWarps + warp < (1 << 13)) g_tal_trace2[blockIdx.x * kTopkWarps + warp] = tal_now();
#endif
}

// ------------------------------------------------------------------------------------------
// phase 2: dense BCE-with-logits at target 0 (+ zero box gradient), foreground terms, finalize
// ------------------------------------------------------------------------------------------
// softplus(x) = max(x,0) + log1p(exp(-|x|));  d/dx = sigmoid(x)
__device__ __forceinline__ void bce_bg_elem(float x, float kc, float &acc, float &g) {
    const float e = __expf(-fabsf(x));
    acc += fmaxf(x, 0.f) + log1pf(e);
    const float r = fast_rcp(1.f + e);
    g = kc * (x >= 0.f ? r : e * r);
}

// The same for two cells at a time, branch-free, on the packed fp32 pipe.  With t = exp(-|x|) in (0, 1] and
// r = 1 / (1 + t) in [0.5, 1):  softplus(x) = max(x, 0) - log(r)  and  sigmoid(x) = r (x >= 0) or t r (x < 0).
// log(r) comes from the degree-5 polynomial of log(1 + f) on [-0.293, 0]: f = r sqrt2 - 1 (and -ln2/2 added) when
// r < 1/sqrt2, else f = -t r, which is r - 1 without the cancellation.  acc2 accumulates softplus.
__device__ __forceinline__ void softplus_sigmoid_pair(float x0, float x1, f32x2 &sp, f32x2 &sg) {
    float a0, a1;
    unpack2(mul2(pack2(fabsf(x0), fabsf(x1)), pack2(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
    const float t0 = fast_ex2(a0), t1 = fast_ex2(a1);
    float u0, u1;
    unpack2(add2(pack2(t0, t1), pack2(1.f, 1.f)), u0, u1);
    const float r0 = fast_rcp(u0), r1 = fast_rcp(u1);
    float tr0, tr1;
    unpack2(mul2(pack2(t0, t1), pack2(r0, r1)), tr0, tr1);                    // t r = 1 - r, exactly enough
    const bool lo0 = r0 < 0.70710678f, lo1 = r1 < 0.70710678f;
    const float f0 = lo0 ? fmaf(r0, 1.41421356f, -1.f) : -tr0, f1 = lo1 ? fmaf(r1, 1.41421356f, -1.f) : -tr1;
    const f32x2 lr = add2(log1p_neg_small2(pack2(f0, f1)), pack2(lo0 ? -0.34657359f : 0.f, lo1 ? -0.34657359f : 0.f));
    sp = fma2(lr, pack2(-1.f, -1.f), pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));     // max(x, 0) - log(r)
    sg = pack2(x0 >= 0.f ? r0 : tr0, x1 >= 0.f ? r1 : tr1);
}
__device__ __forceinline__ void bce_bg_pair(float x0, float x1, f32x2 k2, f32x2 &acc2, float &g0, float &g1) {
    f32x2 sp, sg;
    softplus_sigmoid_pair(x0, x1, sp, sg);
    acc2 = add2(acc2, sp);
    unpack2(mul2(sg, k2), g0, g1);
}

// Varifocal weighting of the background cells (label 0, target 0): w = va * sigmoid(x)^vg, loss = w * softplus(x);
// the weight is differentiated too:  d/dx = w * (vg * (1 - sigmoid) * softplus + sigmoid).
struct VflParams {
    float alpha, gamma;
};
__device__ __forceinline__ float vfl_bg_weight(float sg, const VflParams &vp) {
    return vp.alpha * (vp.gamma == 2.f ? sg * sg : exp2f(vp.gamma * __log2f(sg)));
}
__device__ __forceinline__ void vfl_bg_elem(float x, float kc, const VflParams &vp, float &acc, float &g) {
    const float e = __expf(-fabsf(x));
    const float r = fast_rcp(1.f + e);
    const float sg = x >= 0.f ? r : e * r;
    const float sp = fmaxf(x, 0.f) + log1pf(e);
    const float w = vfl_bg_weight(sg, vp);
    acc += w * sp;
    g = kc * w * fmaf(vp.gamma * (1.f - sg), sp, sg);
}

// two cells, packed: w = alpha s^gamma, loss w sp, gradient k w (gamma (1 - s) sp + s)
__device__ __forceinline__ void vfl_bg_pair(float x0, float x1, f32x2 k2, const VflParams &vp, f32x2 &acc2, float &g0,
                                            float &g1) {
    f32x2 sp, sg;
    softplus_sigmoid_pair(x0, x1, sp, sg);
    f32x2 w;
    if (vp.gamma == 2.f) {
        w = mul2(mul2(sg, sg), pack2(vp.alpha, vp.alpha));
    } else {
        float s0, s1;
        unpack2(sg, s0, s1);
        w = pack2(vfl_bg_weight(s0, vp), vfl_bg_weight(s1, vp));
    }
    acc2 = fma2(w, sp, acc2);
    const f32x2 one_minus = fma2(sg, pack2(-1.f, -1.f), pack2(1.f, 1.f));
    const f32x2 inner = fma2(mul2(one_minus, pack2(vp.gamma, vp.gamma)), sp, sg);
    unpack2(mul2(mul2(w, k2), inner), g0, g1);
}

template <typename T, int VW, bool WRITE_GRAD, bool VFL>
__global__ void __launch_bounds__(kTalThreads)
tal_cls_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, int nc, const float *tss_dev, float lambda_cls,
               VflParams vp, const int *aslot, const float *fgrad, const float *tsc, T *__restrict__ grad,
               float *__restrict__ part) {
    // (tss_dev, aslot, fgrad, tsc are written by the kernel before this one.  Under programmatic dependent launch a load of
    // "read-only" data may be scheduled above pdl_wait(): the normaliser and the slot map, whose addresses are known up
    // front, are therefore read with plain L2 loads right after the wait; fgrad and tsc are addressed THROUGH the slot
    // map, so their loads (common.cuh::ld_dependent_f32) cannot move above it -- and unlike plain loads they may still be
    // scheduled across the streaming stores of the row loop)
    __shared__ float s_red[kTalThreads / 32];
#if YB_TAL_PDL
    pdl_launch_dependents();
    pdl_wait();                                            // the normaliser, the slot map and the per-slot terms
#endif
    // kTalClsSplit CTAs share an anchor tile: each takes a slice of the class rows and of the box rows, so the CTAs
    // are short and the grid has several waves (one CTA per tile left a 1.6-wave grid with a long tail)
    constexpr int kTalClsSplit = tal_cls_split(kTalThreads * VW);
    const int n = blockIdx.y;
    const int tile = blockIdx.x / kTalClsSplit, split = blockIdx.x - tile * kTalClsSplit;
    const int a0 = (tile * kTalThreads + threadIdx.x) * VW;
    const int c_per = (nc + kTalClsSplit - 1) / kTalClsSplit, c_lo = min(split * c_per, nc), c_hi = min(c_lo + c_per, nc);
    constexpr int B_PER = (4 * kRegMax + kTalClsSplit - 1) / kTalClsSplit;
    const int b_lo = min(split * B_PER, 4 * kRegMax), b_hi = min(b_lo + B_PER, 4 * kRegMax);
    const float inv_tss = 1.f / fmaxf(__ldcg(tss_dev), 1.f);
    const float kc = lambda_cls * inv_tss;
    const f32x2 k2 = pack2(kc, kc);
    float acc = 0.f;
    f32x2 acc2 = pack2(0.f, 0.f);                          // packed BCE path: two running sums of softplus
    if (a0 < n_anchors) {
        const size_t img = (size_t)n * n_ch * n_anchors + a0;
        // box rows of the gradient: zero, except the foreground anchors, whose 64 values tal_gt_kernel left in fgrad per
        // unit of the target score (predicated loads, no divergence), here times t / normaliser.  One box row is written per
        // class row of the loop below, so that the kernel's reads and writes stay interleaved instead of opening with a
        // write-only burst.
        int fo[VW];                                        // offset of the anchor's 64 values in fgrad, < 0 = background
        float fs[VW];                                      // t / normaliser of the anchor's slot
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            const int slot = WRITE_GRAD ? __ldcg(aslot + (size_t)n * n_anchors + a0 + v) - 1 : -1;
            fo[v] = slot * (4 * kRegMax);
            fs[v] = slot >= 0 ? ld_dependent_f32(tsc + slot) * inv_tss : 0.f;
        }
        auto box_row = [&](int c) {
            float vals[VW];
#pragma unroll
            for (int v = 0; v < VW; ++v) vals[v] = fo[v] >= 0 ? ld_dependent_f32(fgrad + fo[v] + c) * fs[v] : 0.f;
            Group<T, VW>::store(grad + img + (size_t)c * n_anchors, vals);
        };
        const size_t base = img + (size_t)4 * kRegMax * n_anchors;
        constexpr int U = YB_TAL_CLS_UNROLL;
        Group<T, VW> cur[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c_lo + u < c_hi) cur[u].load(preds + base + (size_t)(c_lo + u) * n_anchors);
        int b_next = b_lo;                                 // next box row of this CTA's slice
        for (int c = c_lo; c < c_hi; c += U) {
            Group<T, VW> nxt[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c + U + u < c_hi) nxt[u].load(preds + base + (size_t)(c + U + u) * n_anchors);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (c + u < c_hi) {
                    float g[VW];
#pragma unroll
                    for (int v = 0; v < VW; ++v) {
                        if (VW > 1) break;                 // scalar fall-back path only
                        if (VFL) vfl_bg_elem(cur[u].get(v), kc, vp, acc, g[v]);
                        else bce_bg_elem(cur[u].get(v), kc, acc, g[v]);
                    }
                    if (VW > 1) {
#pragma unroll
                        for (int v = 0; v + 1 < VW; v += 2) {
                            if (VFL) vfl_bg_pair(cur[u].get(v), cur[u].get(v + 1), k2, vp, acc2, g[v], g[v + 1]);
                            else bce_bg_pair(cur[u].get(v), cur[u].get(v + 1), k2, acc2, g[v], g[v + 1]);
                        }
                    }
                    if (WRITE_GRAD) {
                        Group<T, VW>::store(grad + base + (size_t)(c + u) * n_anchors, g);
                        if (b_next < b_hi) box_row(b_next++);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) cur[u] = nxt[u];
        }
        if (WRITE_GRAD)
            for (; b_next < b_hi; ++b_next) box_row(b_next);       // fewer class rows than box rows in this slice
    }
    {
        float lo, hi;
        unpack2(acc2, lo, hi);
        acc += lo + hi;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kTalThreads / 32; ++w) s += s_red[w];
        part[(size_t)n * gridDim.x + blockIdx.x] = s;
    }
}

// (Measured and not kept: tal_cls_kernel correcting the positive class cell of its own foreground anchors and adding their
// box / DFL terms after its row loop, partials into per-image fixed-point accumulators, this kernel reduced to one CTA:
// the dense kernel goes from 64 to 72 registers and gains a divergent tail -- loss phase 169 -> 193 us.)
// fixed-order reduction in two levels: CTA b sums its slice of every array (tree of fixed shape) and
// publishes 4 partials; the last CTA to finish adds the partials up in index order.
constexpr int kTalFinThreads = 256;
template <typename T>
__global__ void __launch_bounds__(kTalFinThreads)
tal_finalize_kernel(T *__restrict__ grad, const long long *fcell_off, const float4 *fterm, const float *tsc, int n_part,
                    int n_slots, const float *part, unsigned long long *stat_acc, const float *tss_dev, float lambda_box,
                    float lambda_cls, float lambda_dfl, int vfl, VflParams vp, double *__restrict__ cta_sums,
                    unsigned int *__restrict__ ticket, int wipe, unsigned int *__restrict__ img_done, int n_images,
                    float *__restrict__ out_loss) {
    __shared__ double s[3][kTalFinThreads / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if YB_TAL_PDL
    pdl_launch_dependents();                               // the next step's tal_decode_kernel may move in behind this grid
    pdl_wait();                                            // tal_cls_kernel's gradient cells and partial sums
#endif
    const size_t i = (size_t)blockIdx.x * kTalFinThreads + threadIdx.x;
    float f_cls = 0.f, f_box = 0.f, f_dfl = 0.f;
    // every load of the slot is issued at once (one round trip, not a chain of three behind the t >= 0 test)
    const bool slot = i < (size_t)n_slots;
    const float t = slot ? __ldcg(tsc + i) : -1.f;
    const float4 u = slot ? __ldcg(fterm + i) : make_float4(0.f, 0.f, 0.f, 0.f);   // 1 - CIoU, DFL term, class logit, its sigmoid
    const long long cell = slot ? __ldcg(fcell_off + i) : 0ll;
    const float p_i = i < (size_t)n_part ? __ldcg(part + i) : 0.f;
    const float inv_tss = 1.f / fmaxf(__ldcg(tss_dev), 1.f);
    if (t >= 0.f) {                                        // a foreground anchor: the slot's terms, now that t is known
        const float z = u.z, sg = u.w;
        f_box = u.x * t;
        f_dfl = u.y * t;
        if (vfl) {
            // the dense pass counted this cell as background (w_bg * softplus); it is t * BCE(x, t) instead
            const float sp = fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
            f_cls = t * (sp - t * z) - vfl_bg_weight(sg, vp) * sp;
        } else {
            f_cls = -t * z;                                // BCE(x, t) - BCE(x, 0)
        }
        // the anchor's one positive class cell: BCE(x, t) = softplus(x) - t x  ->  (sigmoid(x) - t) / normaliser, over
        // the background value the dense kernel wrote (varifocal: weighted by its own target score, a constant)
        if (grad != nullptr) store_from_float(grad + cell, lambda_cls * (sg - t) * (vfl ? t : 1.f) * inv_tss);
    }
    // fixed-shape trees: butterfly inside each warp, then the warps in index order
    double v[3] = {(double)p_i + (double)f_cls, (double)f_box, (double)f_dfl};
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = warp_sum_d(v[k]);
    if (lane == 0)
        for (int k = 0; k < 3; ++k) s[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) {
            double a = 0.0;
            for (int w = 0; w < kTalFinThreads / 32; ++w) a += s[k][w];
            cta_sums[(size_t)blockIdx.x * 3 + k] = a;
        }
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double acc[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kTalFinThreads)
        for (int k = 0; k < 3; ++k) acc[k] += __ldcg(cta_sums + (size_t)b * 3 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = warp_sum_d(acc[k]);
    if (lane == 0)
        for (int k = 0; k < 3; ++k) s[k][warp] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < 3; ++k)
            for (int w = 1; w < kTalFinThreads / 32; ++w) s[k][0] += s[k][w];
    if (threadIdx.x == 0) {
        const double tss = fmax((double)tss_dev[0], 1.0);
        const float l_cls = (float)(s[0][0] / tss), l_box = (float)(s[1][0] / tss), l_dfl = (float)(s[2][0] / tss);
        // (ticket[10]: a warp of tal_gt_kernel gave up waiting for an image's decode CTAs -- cannot happen, every producer is
        // resident before the first consumer exists; if it ever did, the loss is NaN, not silently wrong)
        out_loss[0] = ticket[10] != 0u ? __int_as_float(0x7fc00000) : lambda_box * l_box + lambda_cls * l_cls + lambda_dfl * l_dfl;
        out_loss[1] = l_box;
        out_loss[2] = l_cls;
        out_loss[3] = l_dfl;
        out_loss[4] = (float)tss;
        out_loss[5] = (float)__ldcg(stat_acc + 2 * kTalStatAcc);   // foreground anchors (counted by tal_fg_kernel)
        out_loss[6] = (float)ticket[3];                    // 1: yb_tal_assign was given a grid hint that does not describe the anchors
        out_loss[7] = (float)ticket[2];                    // GT rows whose class id lies outside [0, nc) (counted by yb_tal_assign)
    }
    // ticket[0] is re-armed for a second yb_tal_loss on the same assignment.  YB_TAL_WS_CLEAN: the step's last reader of the
    // counters puts ALL their zeros back, so that a workspace that started out zeroed stays valid for the next
    // yb_tal_assign without a memset node.
    __syncthreads();
    if (!wipe) {
        if (threadIdx.x == 0) *ticket = 0u;
        return;
    }
    for (int w = threadIdx.x; w < 8 + 2 * kTalStatAcc + 1; w += kTalFinThreads)
        (w < 8 ? reinterpret_cast<unsigned long long *>(ticket) + w : stat_acc + (w - 8))[0] = 0ull;
    for (int b = threadIdx.x; b < n_images; b += kTalFinThreads) img_done[b] = 0u;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <typename T>
static bool tal_vec_ok(const void *preds, const void *grad, int n_anchors) {
    constexpr int VW = ElemsPer16<T>::value;
    return n_anchors % VW == 0 && aligned16(preds) && (grad == nullptr || aligned16(grad));
}

template <typename T, int VW>
static int launch_tal_assign(const T *preds, int n_images, int nc, int n_anchors, const float *anchors,
                             const float *strides, const float *gt, const int32_t *gt_off, int gt_total,
                             const yb_tal_params &p, const TalGrid &grid, const yb_peer_exchange &px, float *out_stats,
                             int32_t *out_assigned, float *out_tscore, const TalWorkspace &w, cudaStream_t st) {
    const int n_ch = 4 * kRegMax + nc;
    // the counters (tickets, statistics, per-GT unit counts) here; the per-anchor arrays by tal_decode_kernel
    // (YB_TAL_WS_CLEAN: the caller vouches that the counters are zero, as the previous step's tal_finalize_kernel left them)
    if (gt_total == 0) YB_CUDA(cudaMemsetAsync(w.ticket, 0, w.zero_bytes, st));
    else if (!(p.flags & YB_TAL_WS_CLEAN)) YB_CUDA(cudaMemsetAsync(w.ticket, 0, w.small_zero_bytes, st));
    if (out_assigned) YB_CUDA(cudaMemsetAsync(out_assigned, 0xff, sizeof(int32_t) * (size_t)n_images * n_anchors, st));
    if (out_tscore) YB_CUDA(cudaMemsetAsync(out_tscore, 0, sizeof(float) * (size_t)n_images * n_anchors, st));
    if (gt_total > 0) {
        constexpr int TILE = kTalThreads * VW;
        const int n_tiles = (n_anchors + TILE - 1) / TILE;
        TalGtArgs<T> A;
        A.preds = preds; A.n_images = n_images; A.n_ch = n_ch; A.n_anchors = n_anchors;
        A.anchors = anchors; A.strides = strides; A.gt = gt; A.gt_off = gt_off; A.gt_total = gt_total; A.topk = p.topk;
        A.alpha = p.alpha; A.beta = p.beta; A.lambda_box = p.lambda_box; A.lambda_dfl = p.lambda_dfl;
        A.dbox = w.dbox; A.gext = w.gext; A.ctr = w.ctr; A.grid = grid;
        A.sel = w.sel; A.sel_count = w.sel_count; A.akey = w.akey; A.fterm = w.fterm; A.fgrad = w.fgrad; A.fcell_off = w.fcell_off;
        A.bad_cls = w.ticket + 2;
#if YB_TAL_PDL
        YB_CUDA(launch_pdl(tal_decode_kernel<T, VW>, dim3(n_tiles, n_images), dim3(kTalThreads), 0, st, preds, n_ch, n_anchors, anchors,
                           strides, gt_off, w.dbox, w.gext, w.ctr, w.akey, w.aslot, w.sel_count, grid, w.ticket + 3, w.img_done));
#else
        tal_decode_kernel<T, VW><<<dim3(n_tiles, n_images), kTalThreads, 0, st>>>(
            preds, n_ch, n_anchors, anchors, strides, gt_off, w.dbox, w.gext, w.ctr, w.akey, w.aslot, w.sel_count, grid, w.ticket + 3,
            w.img_done);
#endif
        YB_LAUNCH_CHECK();
        // warps draw GTs from a counter
        const int gt_ctas = (int)std::min<long long>(((long long)gt_total + kTopkWarps - 1) / kTopkWarps, 148 * YB_TOPK_MINBLOCKS);
        TalResolveArgs R;
        R.tsc = w.tsc; R.aslot = w.aslot; R.out_assigned = out_assigned; R.out_tscore = out_tscore; R.stat_acc = w.stat_acc;
#if YB_TAL_PDL
        YB_CUDA(launch_pdl(tal_gt_kernel<T>, dim3(gt_ctas), dim3(32 * kTopkWarps), 0, st, A, R, w.ticket + 3, w.ticket + 1,
                           (int)(grid.n_levels > 0), out_stats, px, w.img_done, n_tiles));
#else
        tal_gt_kernel<T><<<gt_ctas, 32 * kTopkWarps, 0, st>>>(A, R, w.ticket + 3, w.ticket + 1, grid.n_levels > 0, out_stats, px,
                                                              w.img_done, n_tiles);
#endif
        YB_LAUNCH_CHECK();
    } else {
        YB_CUDA(cudaMemsetAsync(out_stats, 0, sizeof(float) * 8, st));
        if (px.world > 0) {                                // a rank without boxes still owes its peers an entry
            tal_stats_kernel<<<1, 32, 0, st>>>(w.stat_acc, w.ticket + 3, 0, out_stats, px);
            YB_LAUNCH_CHECK();
        }
    }
    return YB_OK;
}

template <typename T, int VW>
static int launch_tal_loss(const T *preds, int n_images, int nc, int n_anchors, int gt_total, const yb_tal_params &p,
                           const float *tss_dev, T *grad, float *out_loss, const TalWorkspace &w, cudaStream_t st) {
    const int n_ch = 4 * kRegMax + nc;
    const VflParams vp{p.vfl_alpha, p.vfl_gamma};
    {
        constexpr int TILE = kTalThreads * VW;
        dim3 grid(((n_anchors + TILE - 1) / TILE) * tal_cls_split(TILE), n_images);
#if YB_TAL_PDL
#define YB_TAL_CLS(WG, VF)                                                                                           \
    YB_CUDA(launch_pdl(tal_cls_kernel<T, VW, WG, VF>, grid, dim3(kTalThreads), 0, st, preds, n_ch, n_anchors, nc, tss_dev, \
                       p.lambda_cls, vp, w.aslot, w.fgrad, w.tsc, grad, w.part))
#else
#define YB_TAL_CLS(WG, VF)                                                                                           \
    tal_cls_kernel<T, VW, WG, VF><<<grid, kTalThreads, 0, st>>>(preds, n_ch, n_anchors, nc, tss_dev, p.lambda_cls, vp, w.aslot, \
                                                                w.fgrad, w.tsc, grad, w.part)
#endif
        if (grad != nullptr) { if (p.vfl) YB_TAL_CLS(true, true); else YB_TAL_CLS(true, false); }
        else { if (p.vfl) YB_TAL_CLS(false, true); else YB_TAL_CLS(false, false); }
#undef YB_TAL_CLS
        YB_LAUNCH_CHECK();
    }
    {
        const int n_part = n_images * w.cls_tiles, n_slots = gt_total * p.topk;
        const int n_max = max(max(n_part, n_slots), 1);
        const int blocks = (n_max + kTalFinThreads - 1) / kTalFinThreads;
#if YB_TAL_PDL
        YB_CUDA(launch_pdl(tal_finalize_kernel<T>, dim3(blocks), dim3(kTalFinThreads), 0, st, grad, w.fcell_off, w.fterm, w.tsc,
                           n_part, n_slots, w.part, w.stat_acc, tss_dev, p.lambda_box, p.lambda_cls, p.lambda_dfl, (int)p.vfl, vp,
                           w.cta_sums, w.ticket, (int)((p.flags & YB_TAL_WS_CLEAN) != 0), w.img_done, n_images, out_loss));
#else
        tal_finalize_kernel<T><<<blocks, kTalFinThreads, 0, st>>>(grad, w.fcell_off, w.fterm, w.tsc, n_part, n_slots, w.part,
                                                                  w.stat_acc, tss_dev, p.lambda_box, p.lambda_cls, p.lambda_dfl,
                                                                  p.vfl, vp, w.cta_sums, w.ticket,
                                                                  (int)((p.flags & YB_TAL_WS_CLEAN) != 0), w.img_done, n_images, out_loss);
#endif
    }
    YB_LAUNCH_CHECK();
    return YB_OK;
}

static int tal_check(const void *preds, void *workspace, const yb_tal_params *p, int dtype, int n_images, int nc, int reg_max,
                     int n_anchors, int gt_total, const char *who) {
    YB_REQUIRE(preds && workspace && p, "%s: null pointer", who);
    YB_REQUIRE(n_images > 0 && n_images <= 65535 && nc > 0 && n_anchors > 0 && gt_total >= 0, "%s: bad sizes", who);
    YB_REQUIRE(reg_max == kRegMax, "%s: reg_max must be %d (got %d)", who, kRegMax, reg_max);
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "%s: dtype must be YB_F32 or YB_BF16", who);
    YB_REQUIRE(p->topk >= 1 && p->topk <= kTalMaxK, "%s: topk must be in [1, %d]", who, kTalMaxK);
    YB_REQUIRE(!p->vfl || (p->vfl_alpha >= 0.f && p->vfl_gamma >= 0.f), "%s: vfl_alpha and vfl_gamma must be non-negative", who);
    YB_REQUIRE(n_anchors <= 65535 * 8, "%s: too many anchors", who);
    if (!aligned16(workspace)) {
        set_error("%s: workspace must be 16-byte aligned", who);
        return YB_ERR_ALIGN;
    }
    return YB_OK;
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_tal_workspace_bytes(int n_images, int n_anchors, int gt_total, int dtype, int topk) {
    if (n_images <= 0 || n_anchors <= 0 || gt_total < 0 || topk < 1 || topk > kTalMaxK) return 0;
    // the larger of the vector tile's and the scalar fall-back's needs (which one runs depends on pointer alignment)
    return std::max(carve_tal(nullptr, n_images, n_anchors, gt_total, topk, tal_tile(dtype, false)).total_bytes,
                    carve_tal(nullptr, n_images, n_anchors, gt_total, topk, tal_tile(dtype, true)).total_bytes);
}

extern "C" int yb_tal_assign(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                             const float *anchors, const float *strides, const float *gt, const int32_t *gt_offsets,
                             int gt_total, const yb_tal_params *params, const yb_tal_grid *grid_hint,
                             const yb_peer_exchange *peers, float *out_stats, int32_t *out_assigned_gt,
                             float *out_target_score, void *workspace, size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_tal_assign");
    if (int rc = tal_check(preds, workspace, params, dtype, n_images, nc, reg_max, n_anchors, gt_total, "yb_tal_assign"))
        return rc;
    YB_REQUIRE(anchors && strides && gt_offsets && out_stats, "yb_tal_assign: null pointer");
    YB_REQUIRE(gt_total == 0 || gt != nullptr, "yb_tal_assign: gt is null but gt_total > 0");
    if (workspace_bytes < yb_tal_workspace_bytes(n_images, n_anchors, gt_total, dtype, params->topk)) {
        set_error("yb_tal_assign: workspace too small");
        return YB_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    yb_peer_exchange px;
    memset(&px, 0, sizeof(px));                            // world = 0: no exchange here (the caller all-reduces out_stats)
    if (peers != nullptr) {
        if (int rc = check_peer(peers, "yb_tal_assign")) return rc;
        px = *peers;
    }
    TalGrid grid;
    memset(&grid, 0, sizeof(grid));                        // n_levels = 0: no hint, generic scan
    if (grid_hint != nullptr && grid_hint->n_levels > 0) {
        // a hint must at least be a partition of [0, A) into h x w grids; its VALUES are verified on the device
        YB_REQUIRE(grid_hint->n_levels <= YB_TAL_MAX_LEVELS, "yb_tal_assign: grid hint has more than %d levels", YB_TAL_MAX_LEVELS);
        long long next = 0;
        for (int l = 0; l < grid_hint->n_levels; ++l) {
            YB_REQUIRE(grid_hint->start[l] == next && grid_hint->w[l] > 0 && grid_hint->h[l] > 0 && grid_hint->stride[l] > 0.f,
                       "yb_tal_assign: grid hint level %d is inconsistent", l);
            next += (long long)grid_hint->w[l] * grid_hint->h[l];
        }
        YB_REQUIRE(next == n_anchors, "yb_tal_assign: grid hint covers %lld anchors, not %d", next, n_anchors);
        grid = *grid_hint;
    }
    if (dtype == YB_F32) {
        const bool vec = tal_vec_ok<float>(preds, nullptr, n_anchors);
        const TalWorkspace w = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec));
        if (vec)
            return launch_tal_assign<float, 4>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt, gt_offsets,
                                               gt_total, *params, grid, px, out_stats, out_assigned_gt, out_target_score, w, st);
        return launch_tal_assign<float, 1>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt, gt_offsets,
                                           gt_total, *params, grid, px, out_stats, out_assigned_gt, out_target_score, w, st);
    }
    const bool vec = tal_vec_ok<__nv_bfloat16>(preds, nullptr, n_anchors);
    const TalWorkspace w = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec));
    if (vec)
        return launch_tal_assign<__nv_bfloat16, 8>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                                   gt_offsets, gt_total, *params, grid, px, out_stats, out_assigned_gt, out_target_score,
                                                   w, st);
    return launch_tal_assign<__nv_bfloat16, 1>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                               gt_offsets, gt_total, *params, grid, px, out_stats, out_assigned_gt, out_target_score, w,
                                               st);
}

extern "C" int yb_tal_loss(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors, int gt_total,
                           const yb_tal_params *params, const float *tss_dev, const yb_peer_exchange *peers,
                           void *grad_preds, float *out_loss, void *workspace, size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_tal_loss");
    if (int rc = tal_check(preds, workspace, params, dtype, n_images, nc, reg_max, n_anchors, gt_total, "yb_tal_loss"))
        return rc;
    YB_REQUIRE((tss_dev || peers) && out_loss, "yb_tal_loss: null pointer");
    if (peers != nullptr)
        if (int rc = check_peer(peers, "yb_tal_loss")) return rc;
    if (workspace_bytes < yb_tal_workspace_bytes(n_images, n_anchors, gt_total, dtype, params->topk)) {
        set_error("yb_tal_loss: workspace too small");
        return YB_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the tile (hence the workspace carving) must be the one yb_tal_assign used: decided by preds only
    if (peers != nullptr) {
        // the exchange, consumer side: wait for every rank's entry of this step, average in rank order -> workspace
        const bool vec_p = dtype == YB_F32 ? tal_vec_ok<float>(preds, nullptr, n_anchors) : tal_vec_ok<__nv_bfloat16>(preds, nullptr, n_anchors);
        const TalWorkspace wp = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec_p));
        if (int rc = launch_peer_wait(*peers, wp.peer_tss, wp.ticket + 4, st)) return rc;
        tss_dev = wp.peer_tss;
    }
    if (dtype == YB_F32) {
        const bool vec_a = tal_vec_ok<float>(preds, nullptr, n_anchors);
        const TalWorkspace w = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec_a));
        if (vec_a && tal_vec_ok<float>(preds, grad_preds, n_anchors))
            return launch_tal_loss<float, 4>((const float *)preds, n_images, nc, n_anchors, gt_total, *params, tss_dev,
                                             (float *)grad_preds, out_loss, w, st);
        YB_REQUIRE(!vec_a, "yb_tal_loss: grad_preds must be 16-byte aligned when preds is");
        return launch_tal_loss<float, 1>((const float *)preds, n_images, nc, n_anchors, gt_total, *params, tss_dev,
                                         (float *)grad_preds, out_loss, w, st);
    }
    const bool vec_a = tal_vec_ok<__nv_bfloat16>(preds, nullptr, n_anchors);
    const TalWorkspace w = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec_a));
    if (vec_a && tal_vec_ok<__nv_bfloat16>(preds, grad_preds, n_anchors))
        return launch_tal_loss<__nv_bfloat16, 8>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, gt_total, *params,
                                                 tss_dev, (__nv_bfloat16 *)grad_preds, out_loss, w, st);
    YB_REQUIRE(!vec_a, "yb_tal_loss: grad_preds must be 16-byte aligned when preds is");
    return launch_tal_loss<__nv_bfloat16, 1>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, gt_total, *params, tss_dev,
                                             (__nv_bfloat16 *)grad_preds, out_loss, w, st);
}
