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
//   yb_tal_assign   tal_candidates_kernel  per (image, anchor tile): decode the tile into shared memory; the
//                                          GTs that meet the tile are found by all threads at once and then
//                                          handed to the warps: group-extent skip + ballot compaction of the
//                                          anchors whose centre lies inside the GT, alignment metric of the
//                                          queue (approximate-reciprocal arithmetic: it only RANKS), and
//                                          either the whole queue (<= 64 entries) or its k best (REDUX
//                                          rounds) are appended to the GT's candidate list
//                   tal_select_kernel      one warp per GT: global top-k (metric desc, anchor asc) with the
//                                          list in registers, 64-bit atomicMax (overlap, ~gt) per anchor
//                                          resolves conflicts
//                   tal_fg_kernel          one half-warp per (GT, selected anchor): did the GT keep the anchor,
//                                          the GT's max metric / overlap -> target score; CIoU and DFL loss and
//                                          the gradient of the anchor's 64 box logits (NOT yet divided by the
//                                          normaliser) into a compact buffer; anchor -> slot map; per-CTA sums of
//                                          the target scores, added as fixed point; the last CTA writes
//                                          [sum of target scores, #foreground]
//   yb_tal_loss     tal_cls_kernel         dense BCE-with-logits at target 0 + gradient; writes the box rows of the
//                                          gradient too: zero, or the foreground anchor's 64 values / normaliser
//                                          (predicated loads through the anchor -> slot map, no scattered stores)
//                   tal_finalize_kernel    patches the one positive class cell of every foreground anchor, then
//                                          fixed-order reduction -> loss scalars
// The anchors x GT overlap / metric matrices never exist.
#include <algorithm>

#include "common.cuh"

namespace yb {

constexpr int kTalThreads = 128;
constexpr int kTalMaxK = 16;
constexpr int kTalAppendAll = 64;          // a (GT, tile) queue of at most this many anchors is appended unselected
constexpr float kEpsCiou = 1e-7f;
constexpr float kEpsIn = 1e-9f;
constexpr float kEpsNorm = 1e-9f;
constexpr float kFourOverPi2 = 0.40528473456935109f;
constexpr int kTalFinThreadsDecl = 256;
constexpr int kTalStatAcc = 16;            // sub-accumulators of the target-score sum (spreads same-address atomics)
constexpr double kTalFix = 4294967296.0;   // 2^32: target scores are summed in fixed point (order-independent, exact)
// CTAs per anchor tile in the dense pass (each takes 1/SPLIT of the class and box rows): measured best 8 for
// 512-anchor (fp32) tiles, 4 for 1024-anchor (bf16) tiles; the scalar fall-back (128-anchor tiles) uses 8 too
__host__ __device__ constexpr int tal_cls_split(int tile) { return tile >= 1024 ? 4 : 8; }
#ifndef YB_TAL_CLS_UNROLL
#define YB_TAL_CLS_UNROLL 4
#endif
#ifndef YB_TAL_CAND_MINBLOCKS
#define YB_TAL_CAND_MINBLOCKS 8
#endif
constexpr int kTalCandMinBlocks = YB_TAL_CAND_MINBLOCKS;
// shared memory of tal_candidates_kernel: box 16 B + centre 8 B per anchor, 16 B per group of 32 anchors,
// and per warp a queue of (metric 4 B, anchor 2 B) per anchor
static size_t tal_cand_smem(int tile) { return (size_t)tile * (16 + 8 + (kTalThreads / 32) * (4 + 2)) + (size_t)(tile / 32) * 16; }

struct TalWorkspace {
    // zeroed by yb_tal_assign
    unsigned int *ticket;               // [0] finalize ticket, [1] fg ticket, [2] GT rows with a class id outside [0, nc)
    unsigned long long *stat_acc;       // [kTalStatAcc + 1] fixed-point sums of the target scores, then #foreground
    int *cand_count;                    // [gt_total]
    unsigned long long *akey;           // [N * A]  (overlap bits << 32) | ~gt_local   (0 = nobody)
    int *aslot;                         // [N * A]  1 + (g * topk + r) of the GT slot that owns the anchor (0 = background)
    // plain scratch
    float4 *cand;                       // [gt_total * cand_cap]  metric, overlap, anchor (as int bits), -
    float4 *sel;                        // [gt_total * kTalMaxK]  anchor bits, metric, overlap, -
    int *sel_count;                     // [gt_total]
    float *fg_box, *fg_dfl, *fg_cls;    // [gt_total * kTalMaxK] per-foreground loss terms (not yet divided by the normaliser)
    float *fgrad;                       // [gt_total * kTalMaxK * 64] box-logit gradient of every foreground anchor (ditto)
    long long *fcell_off;               // [gt_total * kTalMaxK] element offset of the anchor's positive class cell (-1 = none)
    float *fcell_val;                   // [gt_total * kTalMaxK] its gradient (ditto)
    float *part;                        // [N * cls_tiles]
    double *cta_sums;                   // [4 * finalize CTAs]
    int cand_cap, cls_tiles;
    size_t zero_bytes, total_bytes;
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
    off += round_up(sizeof(unsigned long long) * (kTalStatAcc + 1), 64);
    w.cand_count = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * g, 64);
    w.akey = reinterpret_cast<unsigned long long *>(p + off);
    off += round_up(sizeof(unsigned long long) * (size_t)n_images * n_anchors, 64);
    w.aslot = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images * n_anchors, 64);
    w.zero_bytes = off;
    w.cand_cap = std::max(topk, std::min(tile, kTalAppendAll)) * tiles;   // what one tile can append, times the tiles
    w.cls_tiles = tiles * tal_cls_split(tile);             // partial sums per image
    w.cand = reinterpret_cast<float4 *>(p + off);
    off += round_up(sizeof(float4) * g * (size_t)w.cand_cap, 64);
    w.sel = reinterpret_cast<float4 *>(p + off);
    off += round_up(sizeof(float4) * g * kTalMaxK, 64);
    w.sel_count = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * g, 64);
    w.fg_box = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * g * kTalMaxK, 64);
    w.fg_dfl = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * g * kTalMaxK, 64);
    w.fg_cls = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * g * kTalMaxK, 64);
    w.fgrad = reinterpret_cast<float *>(p + off);
    off += round_up(sizeof(float) * g * kTalMaxK * 4 * kRegMax, 64);
    w.fcell_off = reinterpret_cast<long long *>(p + off);
    off += round_up(sizeof(long long) * g * kTalMaxK, 64);
    w.fcell_val = reinterpret_cast<float *>(p + off);
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

// overlap = max(CIoU, 0) of predicted box p and GT box g
__device__ __forceinline__ float overlap_fast(const float4 &p, const float4 &g, float atan_g, float area_g) {
    const float w1 = p.z - p.x, h1 = p.w - p.y + kEpsCiou;
    const float iw = fmaxf(fminf(p.z, g.z) - fmaxf(p.x, g.x), 0.f), ih = fmaxf(fminf(p.w, g.w) - fmaxf(p.y, g.y), 0.f);
    const float inter = iw * ih;
    if (inter <= 0.f) return 0.f;                          // IoU 0: the penalties can only push the CIoU below zero
    const float uni = w1 * h1 + area_g - inter + kEpsCiou;
    const float iou = fast_div(inter, uni);
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

// ------------------------------------------------------------------------------------------
// tal_candidates_kernel
// ------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(kTalThreads, VW == 8 ? 4 : kTalCandMinBlocks)   // bf16 tiles: 4 CTAs of shared memory per SM
tal_candidates_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, const float *__restrict__ anchors,
                      const float *__restrict__ strides, const float *__restrict__ gt, const int *__restrict__ gt_off,
                      int topk, float alpha, float beta, int *__restrict__ cand_count, float4 *__restrict__ cand,
                      int cand_cap) {
    constexpr int TILE = kTalThreads * VW;
    constexpr int NW = kTalThreads / 32;
    constexpr int NG = TILE / 32;                          // groups of 32 consecutive anchors (<= 32)
    constexpr int GL = 32 / VW;                            // lanes that share one group
    constexpr int NE = VW == 8 ? 8 : 4;                    // queue entries per lane the register top-k holds
    // dynamic shared memory: 24 B per anchor of the tile + 6 B per anchor and warp (see tal_cand_smem)
    extern __shared__ __align__(16) unsigned char tal_smem[];
    float4 *s_box = reinterpret_cast<float4 *>(tal_smem);                         // decoded xyxy (pixels)
    float2 *s_ctr = reinterpret_cast<float2 *>(s_box + TILE);                     // anchor centres (pixels)
    float4 *s_grp = reinterpret_cast<float4 *>(s_ctr + TILE);                     // centre extent of each group
    float (*s_qm)[TILE] = reinterpret_cast<float (*)[TILE]>(s_grp + NG);          // per-warp queue: metric
    unsigned short (*s_q)[TILE] = reinterpret_cast<unsigned short (*)[TILE]>(&s_qm[NW][0]);   //     anchor
    __shared__ float s_bb[4][NW];                        // tile extent of the anchor centres
    __shared__ int s_list[kTalThreads];                  // GTs of the current chunk that meet the tile
    __shared__ int s_cnt[NW];
    __shared__ int s_next;

    // Launch order: image fastest, LAST tile first.  The coarse levels sit at the end of the anchor axis and meet
    // nearly every GT of their image, so their CTAs are the long ones: they go out first and the grid drains on short ones.
    const int n = blockIdx.x;
    const int tile0 = ((int)gridDim.y - 1 - (int)blockIdx.y) * TILE;
    const int a0 = tile0 + threadIdx.x * VW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t img = (size_t)n * n_ch * n_anchors;
    const int g_begin = gt_off[n];
    const int m_img = gt_off[n + 1] - g_begin;
    if (m_img == 0) return;                               // uniform per CTA

    float lo_x = __int_as_float(0x7f800000), lo_y = lo_x, hi_x = -lo_x, hi_y = -lo_x;
    if (a0 < n_anchors) {
        float dist[4][VW];
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
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            const float ax = __ldg(anchors + a0 + v), ay = __ldg(anchors + n_anchors + a0 + v), s = __ldg(strides + a0 + v);
            const PredBox b = decode_box(ax, ay, s, dist[0][v], dist[1][v], dist[2][v], dist[3][v]);
            s_box[threadIdx.x * VW + v] = make_float4(b.x1, b.y1, b.x2, b.y2);
            const float cx = ax * s, cy = ay * s;
            s_ctr[threadIdx.x * VW + v] = make_float2(cx, cy);
            lo_x = fminf(lo_x, cx); hi_x = fmaxf(hi_x, cx);
            lo_y = fminf(lo_y, cy); hi_y = fmaxf(hi_y, cy);
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        if (o == GL && (lane & (GL - 1)) == 0)             // extent of this lane's group of 32 anchors
            s_grp[(threadIdx.x * VW) >> 5] = make_float4(lo_x, lo_y, hi_x, hi_y);
        lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if (GL == 32 && lane == 0) s_grp[warp] = make_float4(lo_x, lo_y, hi_x, hi_y);
    if (lane == 0) { s_bb[0][warp] = lo_x; s_bb[1][warp] = lo_y; s_bb[2][warp] = hi_x; s_bb[3][warp] = hi_y; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        lo_x = fminf(lo_x, s_bb[0][w]); lo_y = fminf(lo_y, s_bb[1][w]);
        hi_x = fmaxf(hi_x, s_bb[2][w]); hi_y = fmaxf(hi_y, s_bb[3][w]);
    }
    const int tile_n = min(TILE, n_anchors - tile0);
    float4 my_grp = make_float4(0.f, 0.f, -1.f, -1.f);    // an empty extent never intersects
    if (lane < NG) my_grp = s_grp[lane];
    const int n_cls = n_ch - 4 * kRegMax;

    // GTs are tested against the tile's centre extent 128 at a time, one GT per thread; the hits are compacted into
    // s_list and handed to the warps from a shared counter (a GT that meets the tile costs a warp a lot, one that
    // misses it costs one thread a few compares)
    for (int g0 = 0; g0 < m_img; g0 += kTalThreads) {
        const int gi = g0 + threadIdx.x;
        bool hit = false;
        if (gi < m_img) {
            const float *g5 = gt + (size_t)(g_begin + gi) * 5;
            const float gcx = __ldg(g5), gcy = __ldg(g5 + 1), hw = __ldg(g5 + 2) * 0.5f, hh = __ldg(g5 + 3) * 0.5f;
            hit = gcx - hw < hi_x && gcx + hw > lo_x && gcy - hh < hi_y && gcy + hh > lo_y;
        }
        const unsigned hit_mask = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_cnt[warp] = __popc(hit_mask);
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        int before = 0, n_hit = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            before += w < warp ? s_cnt[w] : 0;
            n_hit += s_cnt[w];
        }
        if (hit) s_list[before + __popc(hit_mask & ((1u << lane) - 1u))] = gi;
        __syncthreads();
        for (;;) {
            int li = 0;
            if (lane == 0) li = atomicAdd(&s_next, 1);
            li = __shfl_sync(0xffffffffu, li, 0);
            if (li >= n_hit) break;
            const int g = s_list[li];
            const float *g5 = gt + (size_t)(g_begin + g) * 5;
            const float gcx = __ldg(g5), gcy = __ldg(g5 + 1), gw = __ldg(g5 + 2), gh = __ldg(g5 + 3);
            const float4 gb = make_float4(gcx - gw * 0.5f, gcy - gh * 0.5f, gcx + gw * 0.5f, gcy + gh * 0.5f);
            // 1. queue the anchors whose centre is strictly inside the GT (ascending anchor order); only the
            //    groups of 32 anchors whose centre extent meets the GT are looked at
            unsigned groups = __ballot_sync(0xffffffffu, gb.x < my_grp.z && gb.z > my_grp.x && gb.y < my_grp.w && gb.w > my_grp.y);
            int nq = 0;
            while (groups) {
                const int a = ((__ffs(groups) - 1) << 5) + lane;
                groups &= groups - 1;
                bool in = false;
                if (a < tile_n) {
                    const float2 c = s_ctr[a];
                    in = fminf(fminf(c.x - gb.x, c.y - gb.y), fminf(gb.z - c.x, gb.w - c.y)) > kEpsIn;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, in);
                if (in) s_q[warp][nq + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)a;
                nq += __popc(mask);
            }
            if (nq == 0) continue;
            const int cls = min(max((int)__ldg(g5 + 4), 0), n_cls - 1);
            const float at_g = gt_atan_fast(gb);
            const float area_g = (gb.z - gb.x) * (gb.w - gb.y + kEpsCiou);
            const T *cls_row = preds + img + (size_t)(4 * kRegMax + cls) * n_anchors + tile0;
            __syncwarp();
            float4 *out = cand + (size_t)(g_begin + g) * cand_cap;
            if (nq <= kTalAppendAll || nq <= topk) {
                // 2a. a short queue goes to the GT's list as it is: tal_select_kernel ranks the whole list anyway
                int slot = 0;
                if (lane == 0) slot = atomicAdd(cand_count + g_begin + g, nq);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                for (int q = lane; q < nq; q += 32) {
                    const int a = s_q[warp][q];
                    const float ov = overlap_fast(s_box[a], gb, at_g, area_g);
                    const float m = metric_fast(load_as_float(cls_row + a), ov, alpha, beta);
                    if (slot + q < cand_cap) out[slot + q] = make_float4(m, ov, __int_as_float(tile0 + a), 0.f);
                }
                __syncwarp();                                     // the queue is reused by the warp's next GT
                continue;
            }
            // 2b. a long queue: its metrics ...
            for (int q = lane; q < nq; q += 32) {
                const int a = s_q[warp][q];
                const float ov = overlap_fast(s_box[a], gb, at_g, area_g);
                s_qm[warp][q] = metric_fast(load_as_float(cls_row + a), ov, alpha, beta);
            }
            __syncwarp();
            // 3. ... and the tile's k best (metric descending, ties -> lowest anchor = lowest queue position).
            //    Metrics are >= 0, so their bit patterns order like signed integers; lane r ends up holding the
            //    r-th best (queue position, metric) and re-evaluates its overlap once, all lanes in parallel.
            const int n_sel = topk;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(cand_count + g_begin + g, n_sel);
            int my_q = lane, my_m = 0;
            if (nq <= 32 * NE) {
                // the usual case: the lane keeps its (at most NE) queue entries q = lane + 32 i in registers
                int v[NE];
#pragma unroll
                for (int i = 0; i < NE; ++i) v[i] = lane + 32 * i < nq ? __float_as_int(s_qm[warp][lane + 32 * i]) : (int)0x80000000;
                for (int r = 0; r < n_sel; ++r) {
                    int bm = v[0], bi = 0;                     // strict: the lowest q of the lane wins its ties
#pragma unroll
                    for (int i = 1; i < NE; ++i)
                        if (v[i] > bm) { bm = v[i]; bi = i; }
                    const int wm = __reduce_max_sync(0xffffffffu, bm);
                    const int wq = __reduce_min_sync(0xffffffffu, bm == wm ? lane + 32 * bi : 0x7fffffff);
                    if (lane == r) { my_q = wq; my_m = wm; }
                    if (lane == (wq & 31)) {
#pragma unroll
                        for (int i = 0; i < NE; ++i)
                            if (i == (wq >> 5)) v[i] = (int)0x80000000;          // taken
                    }
                }
            } else {
                for (int r = 0; r < n_sel; ++r) {
                    int bm = -1, bq = 0x7fffffff;
                    for (int q = lane; q < nq; q += 32) {
                        const int m = __float_as_int(s_qm[warp][q]);
                        if (m > bm) { bm = m; bq = q; }               // strict: first (lowest q) maximum per lane
                    }
                    const int wm = __reduce_max_sync(0xffffffffu, bm);
                    const int wq = __reduce_min_sync(0xffffffffu, bm == wm ? bq : 0x7fffffff);
                    if (lane == r) { my_q = wq; my_m = wm; }
                    if (lane == (wq & 31)) s_qm[warp][wq] = -2.f;     // taken (negative as an integer too)
                    __syncwarp();
                }
            }
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (lane < n_sel && slot + lane < cand_cap) {
                const int a = s_q[warp][my_q];
                const float ov = overlap_fast(s_box[a], gb, at_g, area_g);
                out[slot + lane] = make_float4(__int_as_float(my_m), ov, __int_as_float(tile0 + a), 0.f);
            }
            __syncwarp();                                             // the queue is reused by the warp's next GT
        }
        __syncthreads();                                              // s_list / s_next are rewritten by the next chunk
    }
}

// ------------------------------------------------------------------------------------------
// tal_select_kernel: one warp per GT, global top-k of its candidates
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int gt_image(const int *__restrict__ gt_off, int n_images, int g) {
    int lo = 0, hi = n_images;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(gt_off + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int kSelNE = 12;                                  // list entries a lane keeps in registers: lists up to 384

__global__ void __launch_bounds__(128)
tal_select_kernel(int n_images, int n_anchors, const int *__restrict__ gt_off, int gt_total, int topk,
                  const int *__restrict__ cand_count, float4 *__restrict__ cand, int cand_cap,
                  float4 *__restrict__ sel, int *__restrict__ sel_count, unsigned long long *__restrict__ akey,
                  const float *__restrict__ gt, int n_classes, unsigned int *__restrict__ bad_cls) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (g >= gt_total) return;
    if (lane == 0) {                                       // class ids outside [0, nc) are clamped by the kernels and counted here
        const int c = (int)__ldg(gt + (size_t)g * 5 + 4);
        if (c < 0 || c >= n_classes) atomicAdd(bad_cls, 1u);
    }
    const int n = gt_image(gt_off, n_images, g);
    const int g_local = g - __ldg(gt_off + n);
    const int nc = min(cand_count[g], cand_cap);
    float4 *c = cand + (size_t)g * cand_cap;
    const int n_sel = min(nc, topk);
    auto publish = [&](int r, int wa, float wm, float wo) {            // lane 0
        sel[(size_t)g * kTalMaxK + r] = make_float4(__int_as_float(wa), wm, wo, 0.f);
        // conflict resolution: the anchor goes to the GT with the largest overlap, ties -> lowest GT
        atomicMax(akey + (size_t)n * n_anchors + wa,
                  ((unsigned long long)__float_as_uint(wo) << 32) | (unsigned int)(~(unsigned int)g_local));
    };
    if (nc <= 32 * kSelNE) {
        // the list in registers: (metric bits, anchor) of entries q = lane + 32 i; metrics are >= 0, so their bit
        // patterns order like signed integers; ties -> lowest anchor
        int vm[kSelNE], va[kSelNE];
#pragma unroll
        for (int i = 0; i < kSelNE; ++i) {
            vm[i] = (int)0x80000000; va[i] = 0x7fffffff;
            if (lane + 32 * i < nc) {
                const float4 e = c[lane + 32 * i];
                vm[i] = __float_as_int(e.x); va[i] = __float_as_int(e.z);
            }
        }
        for (int r = 0; r < n_sel; ++r) {
            int bm = vm[0], ba = va[0], bi = 0;
#pragma unroll
            for (int i = 1; i < kSelNE; ++i)
                if (vm[i] > bm || (vm[i] == bm && va[i] < ba)) { bm = vm[i]; ba = va[i]; bi = i; }
            const int wm = __reduce_max_sync(0xffffffffu, bm);
            const int wa = __reduce_min_sync(0xffffffffu, bm == wm ? ba : 0x7fffffff);
            const bool mine = bm == wm && ba == wa;        // exactly one lane: an anchor is in a GT's list once
            const int wl = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            float wo = 0.f;
            if (mine) {
                wo = c[lane + 32 * bi].y;
#pragma unroll
                for (int i = 0; i < kSelNE; ++i)
                    if (i == bi) vm[i] = (int)0x80000000;  // taken
            }
            wo = __shfl_sync(0xffffffffu, wo, wl);
            if (lane == 0) publish(r, wa, __int_as_float(wm), wo);
        }
    } else {
        // long list: scanned in global memory, a selected entry gets metric -2
        for (int r = 0; r < n_sel; ++r) {
            float bm = -1.f;
            int ba = 0x7fffffff, bt = -1;
            float bo = 0.f;
            for (int q = lane; q < nc; q += 32) {
                const float4 e = c[q];
                const int a = __float_as_int(e.z);
                if (e.x > bm || (e.x == bm && a < ba)) { bm = e.x; ba = a; bo = e.y; bt = q; }
            }
            const int wmi = __reduce_max_sync(0xffffffffu, __float_as_int(bm));
            const int wa = __reduce_min_sync(0xffffffffu, __float_as_int(bm) == wmi ? ba : 0x7fffffff);
            const int wl = __ffs(__ballot_sync(0xffffffffu, __float_as_int(bm) == wmi && ba == wa)) - 1;
            const float wo = __shfl_sync(0xffffffffu, bo, wl);
            if (lane == wl && bt >= 0) c[bt].x = -2.f;         // taken (metrics are >= 0)
            __syncwarp();
            if (lane == 0) publish(r, wa, __int_as_float(wmi), wo);
        }
    }
    if (lane == 0) sel_count[g] = n_sel;
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
tal_cls_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, int nc, const float *__restrict__ tss_dev,
               float lambda_cls, VflParams vp, const int *__restrict__ aslot, const float *__restrict__ fgrad,
               T *__restrict__ grad, float *__restrict__ part) {
    __shared__ float s_red[kTalThreads / 32];
    // kTalClsSplit CTAs share an anchor tile: each takes a slice of the class rows and of the box rows, so the CTAs
    // are short and the grid has several waves (one CTA per tile left a 1.6-wave grid with a long tail)
    constexpr int kTalClsSplit = tal_cls_split(kTalThreads * VW);
    const int n = blockIdx.y;
    const int tile = blockIdx.x / kTalClsSplit, split = blockIdx.x - tile * kTalClsSplit;
    const int a0 = (tile * kTalThreads + threadIdx.x) * VW;
    const int c_per = (nc + kTalClsSplit - 1) / kTalClsSplit, c_lo = min(split * c_per, nc), c_hi = min(c_lo + c_per, nc);
    constexpr int B_PER = (4 * kRegMax + kTalClsSplit - 1) / kTalClsSplit;
    const int b_lo = min(split * B_PER, 4 * kRegMax), b_hi = min(b_lo + B_PER, 4 * kRegMax);
    const float inv_tss = 1.f / fmaxf(__ldg(tss_dev), 1.f);
    const float kc = lambda_cls * inv_tss;
    const f32x2 k2 = pack2(kc, kc);
    float acc = 0.f;
    f32x2 acc2 = pack2(0.f, 0.f);                          // packed BCE path: two running sums of softplus
    if (a0 < n_anchors) {
        const size_t img = (size_t)n * n_ch * n_anchors + a0;
        // box rows of the gradient: zero, except the foreground anchors, whose 64 values tal_fg_kernel left in fgrad
        // (predicated loads, no divergence), here divided by the normaliser.  One box row is written per class row of the loop below, so that the
        // kernel's reads and writes stay interleaved instead of opening with a write-only burst.
        int fo[VW];                                        // offset of the anchor's 64 values in fgrad, < 0 = background
#pragma unroll
        for (int v = 0; v < VW; ++v)
            fo[v] = WRITE_GRAD ? (__ldg(aslot + (size_t)n * n_anchors + a0 + v) - 1) * (4 * kRegMax) : -1;
        auto box_row = [&](int c) {
            float vals[VW];
#pragma unroll
            for (int v = 0; v < VW; ++v) vals[v] = fo[v] >= 0 ? __ldg(fgrad + fo[v] + c) * inv_tss : 0.f;
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

// One HALF-warp per (GT, selected anchor) slot: lane & 15 is the DFL bin, and the lane holds that bin of all four
// sides, so the scalar part (CIoU, its gradient, the class cell) is issued once for two slots.  The half-warp first
// works out what tal_select_kernel left open: whether its GT kept the anchor (conflicts went to the larger overlap)
// and the GT's largest metric / overlap over the anchors it kept -> the slot's target score t.  Everything written
// here is NOT yet divided by the normaliser (the sum of all t, possibly over several ranks): tal_cls_kernel and
// tal_finalize_kernel apply 1 / normaliser.  The per-CTA sums of t are added in fixed point, so the statistics the
// last CTA writes do not depend on the order in which CTAs finish.
template <typename T>
__global__ void __launch_bounds__(128)
tal_fg_kernel(const T *__restrict__ preds, int n_images, int n_ch, int n_anchors, int nc,
              const float *__restrict__ anchors, const float *__restrict__ strides, const float *__restrict__ gt,
              const int *__restrict__ gt_off, int gt_total, int topk, const float4 *__restrict__ sel,
              const int *__restrict__ sel_count, const unsigned long long *__restrict__ akey, float lambda_box,
              float lambda_cls, float lambda_dfl, int vfl, VflParams vp, float *__restrict__ fgrad,
              long long *__restrict__ fcell_off, float *__restrict__ fcell_val, float *__restrict__ fg_box,
              float *__restrict__ fg_dfl, float *__restrict__ fg_cls, int *__restrict__ aslot,
              int *__restrict__ out_assigned, float *__restrict__ out_tscore, unsigned long long *__restrict__ stat_acc,
              unsigned int *__restrict__ ticket, float *__restrict__ out_stats) {
    __shared__ float s_t[8];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, bin = lane & 15, base = lane & 16;
    const int slot = blockIdx.x * 8 + (threadIdx.x >> 4);          // slot = g * topk + r
    const int g_raw = slot / topk, r = slot - g_raw * topk;
    const bool in_range = g_raw < gt_total;
    const int g = in_range ? g_raw : 0;
    const int n = gt_image(gt_off, n_images, g);
    const int g_local = g - __ldg(gt_off + n);
    // lane `bin` of the half looks at the GT's bin-th selected anchor
    const int ns = in_range ? sel_count[g] : 0;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    bool pos = false;
    if (bin < ns) {
        e = sel[(size_t)g * kTalMaxK + bin];
        const unsigned long long k = akey[(size_t)n * n_anchors + __float_as_int(e.x)];
        pos = (unsigned int)(k & 0xffffffffull) == (unsigned int)(~(unsigned int)g_local);
    }
    float mm = pos ? e.y : 0.f, mo = pos ? e.z : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    }
    const float my_t = pos ? e.y * (mo / (mm + kEpsNorm)) : 0.f;
    // the half's own slot is entry r
    const int idx_r = __shfl_sync(0xffffffffu, __float_as_int(e.x), base + min(r, 15));
    const float t_r = __shfl_sync(0xffffffffu, my_t, base + min(r, 15));
    const bool live = __shfl_sync(0xffffffffu, (int)pos, base + min(r, 15)) != 0 && r < ns;
    const int idx = live ? idx_r : 0;
    const float t = live ? t_r : 0.f;
    if (bin == 0) s_t[threadIdx.x >> 4] = t;
    if (bin == 0 && in_range) {
        if (!live) { fg_box[slot] = 0.f; fg_dfl[slot] = 0.f; fg_cls[slot] = 0.f; fcell_off[slot] = -1; }
        else {
            aslot[(size_t)n * n_anchors + idx] = slot + 1;
            if (out_assigned) out_assigned[(size_t)n * n_anchors + idx] = g_local;
            if (out_tscore) out_tscore[(size_t)n * n_anchors + idx] = t;
        }
    }
    if (__any_sync(0xffffffffu, live)) {                   // else both halves idle
        // an idle half walks through the same shuffles on harmless stand-in data and writes nothing
        const T *img = preds + (size_t)n * n_ch * n_anchors;
        float z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = load_as_float(img + (size_t)(k * kRegMax + bin) * n_anchors + idx);
        const float *g5 = gt + (size_t)g * 5;
        const float gcx = __ldg(g5), gcy = __ldg(g5 + 1), gw = __ldg(g5 + 2), gh = __ldg(g5 + 3);
        int cls = (int)__ldg(g5 + 4);
        cls = min(max(cls, 0), nc - 1);
        const float z_cls = load_as_float(img + (size_t)(4 * kRegMax + cls) * n_anchors + idx);
        const float ax = __ldg(anchors + idx), ay = __ldg(anchors + n_anchors + idx), s = __ldg(strides + idx);

        // softmax and expectation of each side over the 16 lanes of the half (xor offsets <= 8 stay inside it)
        float mx[4], sm[4], pr[4], ds[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float m = z[k];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float ex = expf(z[k] - m);
            float sum = ex;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float p = __fdiv_rn(ex, sum);
            float d = p * (float)bin;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            mx[k] = m; sm[k] = sum; pr[k] = p; ds[k] = d;
        }
        const PredBox b = decode_box(ax, ay, s, ds[0], ds[1], ds[2], ds[3]);
        const float4 pb = make_float4(b.x1, b.y1, b.x2, b.y2);
        const float4 gb = make_float4(gcx - gw * 0.5f, gcy - gh * 0.5f, gcx + gw * 0.5f, gcy + gh * 0.5f);

        // ---- CIoU and its gradient w.r.t. the predicted corners (alpha_v constant) --------------------
        const Ciou c = ciou_eval(pb, gb, gt_atan(gb));
        auto w_gt = [](float a, float o) { return a > o ? 1.f : (a == o ? 0.5f : 0.f); };     // d max(a,o)/da
        auto w_lt = [](float a, float o) { return a < o ? 1.f : (a == o ? 0.5f : 0.f); };     // d min(a,o)/da
        const float iw = fmaxf(c.iw_raw, 0.f), ih = fmaxf(c.ih_raw, 0.f);
        const float d_inter = (c.uni + c.inter) / (c.uni * c.uni);       // d iou / d inter (union contains -inter)
        const float d_area1 = -c.inter / (c.uni * c.uni);
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
        const float hyp = c.h1 * c.h1 + c.w1 * c.w1;
        const float dv_dw1 = dv_dA1 * (c.h1 / hyp), dv_dh1 = dv_dA1 * (-c.w1 / hyp);
        gx1 -= c.alpha * (-dv_dw1);
        gx2 -= c.alpha * dv_dw1;
        gy1 -= c.alpha * (-dv_dh1);
        gy2 -= c.alpha * dv_dh1;
        // L_box = (1 - ciou) * t / tss * lambda_box   (1 / tss applied later)
        const float kb = -lambda_box * t;
        const float dd[4] = {kb * gx1 * (-s), kb * gy1 * (-s), kb * gx2 * s, kb * gy2 * s};   // d / d (dl, dt, dr, db)

        // ---- DFL rows (same target rule as the reference, src/model/losses.py:226-246, :63-78) -------
        const float tgt[4] = {ax - gb.x / s, ay - gb.y / s, gb.z / s - ax, gb.w / s - ay};
        const float hi_clamp = (float)(kRegMax - 1 - 0.01);
        const float kd = lambda_dfl * t * 0.25f;
        float dfl4 = 0.f, gk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float tk = fminf(fmaxf(tgt[k], 0.f), hi_clamp);
            const int bl = (int)tk;
            const float wl = (float)(bl + 1) - tk, wr = tk - (float)bl;
            const float lpk = (z[k] - mx[k]) - logf(sm[k]);               // log-softmax of this lane's bin
            dfl4 -= __shfl_sync(0xffffffffu, lpk, base + bl) * wl + __shfl_sync(0xffffffffu, lpk, base + bl + 1) * wr;
            const float oh = (bin == bl ? wl : 0.f) + (bin == bl + 1 ? wr : 0.f);
            gk[k] = kd * ((wl + wr) * pr[k] - oh) + dd[k] * pr[k] * ((float)bin - ds[k]);
        }
        if (live) {                                            // no shuffles below
            // compact, coalesced: the dense kernel merges these 64 values (times 1 / normaliser) into the anchor's box rows
#pragma unroll
            for (int k = 0; k < 4; ++k) fgrad[(size_t)slot * (4 * kRegMax) + k * kRegMax + bin] = gk[k];
            if (bin == 0) {
                // the anchor's one positive class cell: BCE(x, t) = softplus(x) - t x  ->  (sigmoid(x) - t) / tss;
                // patched in by tal_finalize_kernel after the dense kernel has written the background value
                const float sg = __fdiv_rn(1.f, 1.f + expf(-z_cls));
                fcell_off[slot] = (long long)((size_t)n * n_ch * n_anchors + (size_t)(4 * kRegMax + cls) * n_anchors + idx);
                // varifocal: the positive cell is weighted by its own target score, a constant
                fcell_val[slot] = lambda_cls * (sg - t) * (vfl ? t : 1.f);
                fg_box[slot] = (1.f - c.value) * t;
                fg_dfl[slot] = dfl4 * 0.25f * t;
                if (vfl) {
                    // the dense pass counted this cell as background (w_bg * softplus); it is t * BCE(x, t) instead
                    const float sp = fmaxf(z_cls, 0.f) + log1pf(expf(-fabsf(z_cls)));
                    fg_cls[slot] = t * (sp - t * z_cls) - vfl_bg_weight(sg, vp) * sp;
                } else {
                    fg_cls[slot] = -t * z_cls;                       // BCE(x, t) - BCE(x, 0)
                }
            }
        }
    }
    // ---- statistics: sum of the target scores and number of foreground anchors ----------------------
    {   // foreground count: exact integer atomics, one per warp, fenced before the CTA takes its ticket
        const unsigned live_mask = __ballot_sync(0xffffffffu, live && bin == 0);
        if (lane == 0 && live_mask) {
            atomicAdd(stat_acc + kTalStatAcc, (unsigned long long)__popc(live_mask));
            __threadfence();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += s_t[k];                // fixed order within the CTA
        atomicAdd(stat_acc + (blockIdx.x & (kTalStatAcc - 1)), (unsigned long long)__double2ll_rn((double)sum * kTalFix));
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int k = 0; k < kTalStatAcc; ++k) acc += (long long)__ldcg(stat_acc + k);
        out_stats[0] = (float)((double)acc / kTalFix);            // local sum of target scores (un-clamped)
        out_stats[1] = (float)__ldcg(stat_acc + kTalStatAcc);      // foreground anchors
#pragma unroll
        for (int i = 2; i < 8; ++i) out_stats[i] = 0.f;
    }
}

// fixed-order reduction in two levels: CTA b sums its slice of every array (tree of fixed shape) and
// publishes 4 partials; the last CTA to finish adds the partials up in index order.
constexpr int kTalFinThreads = 256;
template <typename T>
__global__ void __launch_bounds__(kTalFinThreads)
tal_finalize_kernel(T *__restrict__ grad, const long long *__restrict__ fcell_off, const float *__restrict__ fcell_val,
                    int n_part, int n_slots, int gt_total, const float *__restrict__ part,
                    const float *__restrict__ fg_box, const float *__restrict__ fg_dfl, const float *__restrict__ fg_cls,
                    const unsigned long long *__restrict__ stat_acc, const float *__restrict__ tss_dev, float lambda_box,
                    float lambda_cls, float lambda_dfl, double *__restrict__ cta_sums, unsigned int *__restrict__ ticket,
                    float *__restrict__ out_loss) {
    __shared__ double s[3][kTalFinThreads];
    __shared__ bool s_last;
    (void)gt_total;
    const size_t i = (size_t)blockIdx.x * kTalFinThreads + threadIdx.x;
    if (grad != nullptr && i < (size_t)n_slots) {          // positive class cell of every foreground anchor
        const long long cell = fcell_off[i];
        if (cell >= 0) store_from_float(grad + cell, fcell_val[i] * (1.f / fmaxf(__ldg(tss_dev), 1.f)));
    }
    s[0][threadIdx.x] = (i < (size_t)n_part ? (double)part[i] : 0.0) + (i < (size_t)n_slots ? (double)fg_cls[i] : 0.0);
    s[1][threadIdx.x] = i < (size_t)n_slots ? (double)fg_box[i] : 0.0;
    s[2][threadIdx.x] = i < (size_t)n_slots ? (double)fg_dfl[i] : 0.0;
    __syncthreads();
    for (int o = kTalFinThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) cta_sums[(size_t)blockIdx.x * 3 + k] = s[k][0];
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double acc[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kTalFinThreads)
        for (int k = 0; k < 3; ++k) acc[k] += __ldcg(cta_sums + (size_t)b * 3 + k);
    for (int k = 0; k < 3; ++k) s[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int o = kTalFinThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double tss = fmax((double)tss_dev[0], 1.0);
        const float l_cls = (float)(s[0][0] / tss), l_box = (float)(s[1][0] / tss), l_dfl = (float)(s[2][0] / tss);
        out_loss[0] = lambda_box * l_box + lambda_cls * l_cls + lambda_dfl * l_dfl;
        out_loss[1] = l_box;
        out_loss[2] = l_cls;
        out_loss[3] = l_dfl;
        out_loss[4] = (float)tss;
        out_loss[5] = (float)__ldcg(stat_acc + kTalStatAcc);      // foreground anchors (counted by tal_fg_kernel)
        out_loss[6] = 0.f;
        out_loss[7] = (float)ticket[2];                    // GT rows whose class id lies outside [0, nc) (counted by yb_tal_assign)
        *ticket = 0u;                                      // re-armed for the next yb_tal_loss on this workspace
    }
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
                             const yb_tal_params &p, float *out_stats, int32_t *out_assigned, float *out_tscore,
                             const TalWorkspace &w, cudaStream_t st) {
    const int n_ch = 4 * kRegMax + nc;
    YB_CUDA(cudaMemsetAsync(w.ticket, 0, w.zero_bytes, st));
    if (out_assigned) YB_CUDA(cudaMemsetAsync(out_assigned, 0xff, sizeof(int32_t) * (size_t)n_images * n_anchors, st));
    if (out_tscore) YB_CUDA(cudaMemsetAsync(out_tscore, 0, sizeof(float) * (size_t)n_images * n_anchors, st));
    if (gt_total > 0) {
        constexpr int TILE = kTalThreads * VW;
        dim3 grid(n_images, (n_anchors + TILE - 1) / TILE);
        const size_t smem = tal_cand_smem(TILE);
        YB_CUDA(cudaFuncSetAttribute(tal_candidates_kernel<T, VW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tal_candidates_kernel<T, VW><<<grid, kTalThreads, smem, st>>>(preds, n_ch, n_anchors, anchors, strides, gt, gt_off,
                                                                     p.topk, p.alpha, p.beta, w.cand_count, w.cand, w.cand_cap);
        YB_LAUNCH_CHECK();
        tal_select_kernel<<<(gt_total + 3) / 4, 128, 0, st>>>(n_images, n_anchors, gt_off, gt_total, p.topk, w.cand_count, w.cand,
                                                              w.cand_cap, w.sel, w.sel_count, w.akey, gt, nc, w.ticket + 2);
        YB_LAUNCH_CHECK();
        const int slots = gt_total * p.topk;
        tal_fg_kernel<T><<<(slots + 7) / 8, 128, 0, st>>>(preds, n_images, n_ch, n_anchors, nc, anchors, strides, gt, gt_off,
                                                         gt_total, p.topk, w.sel, w.sel_count, w.akey, p.lambda_box, p.lambda_cls,
                                                         p.lambda_dfl, p.vfl, VflParams{p.vfl_alpha, p.vfl_gamma}, w.fgrad,
                                                         w.fcell_off, w.fcell_val, w.fg_box, w.fg_dfl, w.fg_cls, w.aslot,
                                                         out_assigned, out_tscore, w.stat_acc, w.ticket + 1, out_stats);
        YB_LAUNCH_CHECK();
    } else {
        YB_CUDA(cudaMemsetAsync(out_stats, 0, sizeof(float) * 8, st));
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
#define YB_TAL_CLS(WG, VF)                                                                                           \
    tal_cls_kernel<T, VW, WG, VF><<<grid, kTalThreads, 0, st>>>(preds, n_ch, n_anchors, nc, tss_dev, p.lambda_cls, vp, w.aslot, \
                                                                w.fgrad, grad, w.part)
        if (grad != nullptr) { if (p.vfl) YB_TAL_CLS(true, true); else YB_TAL_CLS(true, false); }
        else { if (p.vfl) YB_TAL_CLS(false, true); else YB_TAL_CLS(false, false); }
#undef YB_TAL_CLS
        YB_LAUNCH_CHECK();
    }
    {
        const int n_part = n_images * w.cls_tiles, n_slots = gt_total * p.topk;
        const int n_max = max(max(n_part, n_slots), 1);
        const int blocks = (n_max + kTalFinThreads - 1) / kTalFinThreads;
        tal_finalize_kernel<T><<<blocks, kTalFinThreads, 0, st>>>(grad, w.fcell_off, w.fcell_val, n_part, n_slots, gt_total,
                                                                  w.part, w.fg_box, w.fg_dfl, w.fg_cls, w.stat_acc, tss_dev,
                                                                  p.lambda_box, p.lambda_cls, p.lambda_dfl, w.cta_sums, w.ticket,
                                                                  out_loss);
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
                             int gt_total, const yb_tal_params *params, float *out_stats, int32_t *out_assigned_gt,
                             float *out_target_score, void *workspace, size_t workspace_bytes, void *stream) {
    if (int rc = tal_check(preds, workspace, params, dtype, n_images, nc, reg_max, n_anchors, gt_total, "yb_tal_assign"))
        return rc;
    YB_REQUIRE(anchors && strides && gt_offsets && out_stats, "yb_tal_assign: null pointer");
    YB_REQUIRE(gt_total == 0 || gt != nullptr, "yb_tal_assign: gt is null but gt_total > 0");
    if (workspace_bytes < yb_tal_workspace_bytes(n_images, n_anchors, gt_total, dtype, params->topk)) {
        set_error("yb_tal_assign: workspace too small");
        return YB_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == YB_F32) {
        const bool vec = tal_vec_ok<float>(preds, nullptr, n_anchors);
        const TalWorkspace w = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec));
        if (vec)
            return launch_tal_assign<float, 4>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt, gt_offsets,
                                               gt_total, *params, out_stats, out_assigned_gt, out_target_score, w, st);
        return launch_tal_assign<float, 1>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt, gt_offsets,
                                           gt_total, *params, out_stats, out_assigned_gt, out_target_score, w, st);
    }
    const bool vec = tal_vec_ok<__nv_bfloat16>(preds, nullptr, n_anchors);
    const TalWorkspace w = carve_tal(workspace, n_images, n_anchors, gt_total, params->topk, tal_tile(dtype, vec));
    if (vec)
        return launch_tal_assign<__nv_bfloat16, 8>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                                   gt_offsets, gt_total, *params, out_stats, out_assigned_gt, out_target_score,
                                                   w, st);
    return launch_tal_assign<__nv_bfloat16, 1>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                               gt_offsets, gt_total, *params, out_stats, out_assigned_gt, out_target_score, w,
                                               st);
}

extern "C" int yb_tal_loss(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors, int gt_total,
                           const yb_tal_params *params, const float *tss_dev, void *grad_preds, float *out_loss,
                           void *workspace, size_t workspace_bytes, void *stream) {
    if (int rc = tal_check(preds, workspace, params, dtype, n_images, nc, reg_max, n_anchors, gt_total, "yb_tal_loss"))
        return rc;
    YB_REQUIRE(tss_dev && out_loss, "yb_tal_loss: null pointer");
    if (workspace_bytes < yb_tal_workspace_bytes(n_images, n_anchors, gt_total, dtype, params->topk)) {
        set_error("yb_tal_loss: workspace too small");
        return YB_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the tile (hence the workspace carving) must be the one yb_tal_assign used: decided by preds only
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
