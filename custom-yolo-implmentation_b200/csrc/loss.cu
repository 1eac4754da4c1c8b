// Fused decode + nearest-centre assignment + DFL/QFL loss + backward (sm_100a).
//
// Replaces YoloDFLQFLoss.forward and its autograd backward (src/model/losses.py:93-281).
// ONE launch per step, every head-output byte read once and every gradient byte written once; the (M x A)
// distance matrix, the decoded boxes and the dense (A x nc) QFL target never exist:
//
//   fused_main_kernel  CTAs of four roles (YB_LOSS_SPLIT_LAUNCH runs them as the separate kernels assign_kernel /
//                      cls_loss_kernel / match_kernel):
//     box role    reads the 4*16 box channels (128-bit loads), decodes each anchor's predicted centre,
//                 zero-fills the box-channel gradient, drops the GTs that cannot find their nearest centre
//                 in this tile (extent of the tile's centres vs the distance already published), scans the
//                 rest (one GT per thread, anchors streamed from shared memory) and merges the per-CTA
//                 winners with a 64-bit atomicMax on (distance, anchor) keys.
//     class role  reads the nc class channels: QFL loss + gradient of every cell for target 0 (all but
//                 <= one cell per GT), per-CTA partial sums.  Independent of the matching.
//     match role  one half-warp per GT, after every tile of the GT's image has counted itself off (in-kernel
//                 dependency, common.cuh): gathers the 64 logits of the matched anchor, DFL loss and its
//                 gradient, IoU soft target (reference formula, slip included) and the gradient that flows
//                 through it, duplicate-anchor resolution; corrects the one positive QFL cell of each matched
//                 anchor (loss delta + gradient); the anchor's 64 box-gradient values go out as whole 32-byte sectors.
//     reducer     the last CTA: waits for every image, turns the per-image accumulators into the loss scalars.
// Partial sums (one per class-role CTA, three per GT) are added into per-image 64-bit FIXED-POINT
// accumulators (2^-32 resolution) with integer atomics: integer addition is associative, so the loss is
// run-to-run bit-identical whatever the order in which CTAs finish, and nothing has to be re-read.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace yb {

#ifndef YB_ASSIGN_THREADS
#define YB_ASSIGN_THREADS 128
#endif
#ifndef YB_CLS_THREADS
#define YB_CLS_THREADS 128
#endif
#ifndef YB_CLS_UNROLL                 // class rows in flight per thread, bf16 rows (measured at cfg5: 4 -> 157.3 us, 5 -> 161.0, 8 -> 182.8)
#define YB_CLS_UNROLL 4
#endif
#ifndef YB_CLS_UNROLL_F32             // ... fp32 rows: a class CTA's 20 rows as 4 batches of 5 instead of 5 of 4 (cfg2: 234.9 -> 232.9 us;
#define YB_CLS_UNROLL_F32 5           // 8 spills at 80 registers: 248.5)
#endif
#ifndef YB_CLS_CSPLIT
#define YB_CLS_CSPLIT 4
#endif
#ifndef YB_CLS_MINBLOCKS
#define YB_CLS_MINBLOCKS 1
#endif
#ifndef YB_ASSIGN_MINBLOCKS          // resident CTAs/SM the register allocator must allow (measured: 4 best
#define YB_ASSIGN_MINBLOCKS 0       // for fp32 rows, 6 for bf16 rows; 0 = pick by row type)
#endif
#ifndef YB_SCAN_UNROLL
#define YB_SCAN_UNROLL 4
#endif
#ifndef YB_ASSIGN_PRUNE
#define YB_ASSIGN_PRUNE 1
#endif
#ifndef YB_BOX_SKEW                   // tile groups by which a box CTA precedes its tile's class CTAs in launch order
#define YB_BOX_SKEW 256
#endif
#ifndef YB_COARSE_FRAC256
#define YB_COARSE_FRAC256 61
#endif
#ifndef YB_COARSE_FIRST
#define YB_COARSE_FIRST 1
#endif
#ifndef YB_MATCH_MAX_CTAS
#define YB_MATCH_MAX_CTAS 64
#endif
#ifndef YB_MATCH_SEPARATE             // 1: match_kernel as a launch of its own after fused_main_kernel (measurement aid)
#define YB_MATCH_SEPARATE 0
#endif
#ifndef YB_MATCH_WHOLE_SECTORS
#define YB_MATCH_WHOLE_SECTORS 1
#endif
#ifndef YB_BF16_SHORT_LOG             // bf16 head outputs: degree-2 log polynomial in the class role
#define YB_BF16_SHORT_LOG 1
#endif
#ifndef YB_LOSS_PROBE                 // 0: ignore the grid hint (measurement aid)
#define YB_LOSS_PROBE 1
#endif
#ifndef YB_LOSS_PDL                   // 1: the launch may become resident behind its predecessor's last wave (programmatic dependent launch)
#define YB_LOSS_PDL 1
#endif
#ifndef YB_PROBE_LEVELS               // the probe role looks at the GT's home cell on the k finest levels (0: on every level)
#define YB_PROBE_LEVELS 1
#endif
#ifndef YB_PROBE_MIN_GT               // the probe role runs when an image can hold more GTs than this (gmax)
#define YB_PROBE_MIN_GT 128
#endif
#ifndef YB_LOSS_POLL_NS               // longest sleep between two polls of a waiting match CTA / the reducer: the launch ENDS on these
#define YB_LOSS_POLL_NS 256           // waits (cfg2 233.2 -> 232.1 us, cfg5 158.5 -> 156.5 against 1024 ns)
#endif
constexpr unsigned int kLossPollNs = YB_LOSS_POLL_NS;
constexpr int kScanUnroll = YB_SCAN_UNROLL;
constexpr int kAssignThreads = YB_ASSIGN_THREADS;
constexpr int kClsThreads = YB_CLS_THREADS;
constexpr unsigned long long kNoKey = ~0ull;

constexpr int kAccCls = 16;           // class-role sub-accumulators per image (spreads same-address atomics)
constexpr int kAccDfl = 16, kAccDcls = 17, kAccWin = 18, kAccPerImage = 20;
constexpr double kFixScale = 4294967296.0;   // 2^32

// The marked arrays must be zero when a call starts: memset by the entry point, or -- YB_LOSS_WS_CLEAN -- left zero by the
// previous call, whose last CTAs wipe every word the launch used (the image's last match CTA the image's keys and bounds,
// the reducer the rest).
struct LossWorkspace {
    unsigned int *ticket;          // [16] [0] match_kernel ticket, [1] non-finite partial seen, [2] class ids out of range, } zero
                                   //      [3] a consumer CTA timed out (common.cuh::dep_wait), [4] probe CTAs done          } at the
    unsigned int *done;            // [N] CTAs of fused_main_kernel that have finished with the image                        } start
    unsigned long long *acc;       // [N * kAccPerImage] fixed-point sums: 16 x class part, DFL, QFL cell correction, winners }
    unsigned long long *best;      // [gt_total] inverted (distance, anchor) keys                                           }
    unsigned int *bound;           // [gt_total] bits of an upper bound of the GT's nearest-centre distance (0 = none yet)   }
    int *gt_img;                   // [gt_total] image of each GT (written by the box role, read by match_kernel)
    size_t zero_bytes;
    size_t total_bytes;
};

static LossWorkspace carve(void *base, int n_images, int gt_total) {
    LossWorkspace w;
    char *p = static_cast<char *>(base);
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned int *>(p + off);
    off += 64;
    w.done = reinterpret_cast<unsigned int *>(p + off);
    off += round_up(sizeof(unsigned int) * (size_t)n_images, 64);
    w.acc = reinterpret_cast<unsigned long long *>(p + off);
    off += round_up(sizeof(unsigned long long) * (size_t)n_images * kAccPerImage, 64);
    w.best = reinterpret_cast<unsigned long long *>(p + off);
    off += round_up(sizeof(unsigned long long) * (size_t)(gt_total > 0 ? gt_total : 1), 64);
    w.bound = reinterpret_cast<unsigned int *>(p + off);
    off += round_up(sizeof(unsigned int) * (size_t)(gt_total > 0 ? gt_total : 1), 64);
    w.zero_bytes = off;
    w.gt_img = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)(gt_total > 0 ? gt_total : 1), 64);
    w.total_bytes = off;
    return w;
}

// float partial -> fixed point, added with an integer atomic (two's complement: negative values work)
__device__ __forceinline__ void acc_add(unsigned long long *slot, float v, unsigned int *flags) {
    if (!isfinite(v)) {                                   // NaN / Inf must reach the loss, not vanish in the conversion
        atomicOr(flags + 1, 1u);
        return;
    }
    atomicAdd(slot, (unsigned long long)__double2ll_rn((double)v * kFixScale));
}
__device__ __forceinline__ double acc_get(const unsigned long long *slot) {
    return (double)(long long)__ldcg(slot) * (1.0 / kFixScale);
}

// ------------------------------------------------------------------------------------------
// assign_kernel
// ------------------------------------------------------------------------------------------
template <typename T, int VW>
__device__ __forceinline__ void assign_body(int n, int tile, const T *__restrict__ preds, int n_ch, int n_anchors,
                                            const float *__restrict__ anchors, const float *__restrict__ strides,
                                            const float *__restrict__ gt, const int *__restrict__ gt_off,
                                            unsigned long long *__restrict__ best, const unsigned int *bound,
                                            int *__restrict__ gt_img, T *__restrict__ grad, bool prune) {
    constexpr int TILE = kAssignThreads * VW;
    constexpr int TILE4 = (TILE + 3) & ~3;
    // predicted centres of the tile's anchors, structure-of-arrays so that four anchors are one LDS.128
    __shared__ __align__(16) float s_x[TILE4], s_y[TILE4], s_p[TILE4];     // cx, cy, cx^2 + cy^2
    __shared__ unsigned long long s_key[kAssignThreads];
    __shared__ float s_ext[4][kAssignThreads / 32];        // per warp: extent of the predicted centres
    __shared__ int s_list[2 * kAssignThreads];             // GTs (image-local index) that still need this tile
    __shared__ int s_cnt[kAssignThreads / 32];
    float lo_x = __int_as_float(0x7f800000), lo_y = lo_x, hi_x = -lo_x, hi_y = -lo_x;

    const int tile0 = tile * TILE;
    const int a0 = tile0 + threadIdx.x * VW;
    const size_t img = (size_t)n * n_ch * n_anchors;
    const int g_begin = gt_off[n];
    const int m_img = gt_off[n + 1] - g_begin;
    if (tile == 0 && gt_img != nullptr)                    // GT -> image table for match_kernel (not needed by the match role)
        for (int m = threadIdx.x; m < m_img; m += kAssignThreads) gt_img[g_begin + m] = n;

    if (a0 < n_anchors) {
        float dist[4][VW];
#pragma unroll
        for (int side = 0; side < 4; ++side) {
            DflPartial part[VW];
#pragma unroll
            for (int h = 0; h < 2; ++h) {                  // 8 rows (128 B per thread) in flight at a time
                Group<T, VW> row[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    row[j].load(preds + img + (size_t)(side * kRegMax + h * 8 + j) * n_anchors + a0);
                if (grad != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        Group<T, VW>::store_zero(grad + img + (size_t)(side * kRegMax + h * 8 + j) * n_anchors + a0);
                }
                if (m_img > 0) {
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
        if (m_img > 0) {
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                const float ax = __ldg(anchors + a0 + v), ay = __ldg(anchors + n_anchors + a0 + v);
                const float s = __ldg(strides + a0 + v);
                const PredBox b = decode_box(ax, ay, s, dist[0][v], dist[1][v], dist[2][v], dist[3][v]);
                s_x[threadIdx.x * VW + v] = b.cx;
                s_y[threadIdx.x * VW + v] = b.cy;
                s_p[threadIdx.x * VW + v] = __fadd_rn(__fmul_rn(b.cx, b.cx), __fmul_rn(b.cy, b.cy));
                lo_x = fminf(lo_x, b.cx); hi_x = fmaxf(hi_x, b.cx);
                lo_y = fminf(lo_y, b.cy); hi_y = fmaxf(hi_y, b.cy);
            }
        }
    } else if (m_img > 0) {
        // past the last anchor: a centre at infinite distance never wins
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            s_x[threadIdx.x * VW + v] = 0.f;
            s_y[threadIdx.x * VW + v] = 0.f;
            s_p[threadIdx.x * VW + v] = __int_as_float(0x7f800000);
        }
    }
    if (m_img == 0) return;                                // uniform per CTA
    if (TILE4 != TILE && threadIdx.x < TILE4 - TILE) {
        s_x[TILE + threadIdx.x] = 0.f;
        s_y[TILE + threadIdx.x] = 0.f;
        s_p[TILE + threadIdx.x] = __int_as_float(0x7f800000);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_ext[0][threadIdx.x >> 5] = lo_x; s_ext[1][threadIdx.x >> 5] = lo_y;
        s_ext[2][threadIdx.x >> 5] = hi_x; s_ext[3][threadIdx.x >> 5] = hi_y;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kAssignThreads / 32; ++w) {
        lo_x = fminf(lo_x, s_ext[0][w]); lo_y = fminf(lo_y, s_ext[1][w]);
        hi_x = fmaxf(hi_x, s_ext[2][w]); hi_y = fmaxf(hi_y, s_ext[3][w]);
    }
    // largest |p|^2 of the tile, for the rounding bound of the matmul-form distance
    const float pn_max = fmaxf(lo_x * lo_x, hi_x * hi_x) + fmaxf(lo_y * lo_y, hi_y * hi_y);

    const int tile_n = min(TILE, n_anchors - tile0);
    const int tile_n4 = (tile_n + 3) & ~3;                 // entries in [tile_n, tile_n4) are at infinity
    // GTs are tested 128 at a time and the survivors appended to s_list; whenever 128 survivors are waiting
    // (or the GTs are exhausted) they are scanned as one chunk: thread <-> one GT; when the chunk is small
    // the anchors of the tile are split into slices so that all warps have work.  Lanes of a warp share
    // the slice, so every shared-memory read below is a broadcast.  Four anchors per step, two packed pairs.
    int count = 0, g0 = 0;                                  // uniform
    while (g0 < m_img || count > 0) {
        while (g0 < m_img && count < kAssignThreads) {
            // Which GTs can still find (or tie) their nearest centre in this tile?  `best` holds what the
            // CTAs that already finished found; a GT whose distance to the extent of this tile's centres
            // exceeds that (with the rounding of the matmul-form d^2 and of the sqrt on its side) cannot.
            // Exact whatever the timing: a stale or empty entry only means less pruning.
            const int gi = g0 + threadIdx.x;
            bool need = gi < m_img;
#if YB_ASSIGN_PRUNE
            if (need && prune) {
                const unsigned long long kinv = __ldcg(best + g_begin + gi);
                // ... or the probe role's bound (the distance to SOME predicted centre, inflated: see probe_body)
                const unsigned int pb = bound != nullptr ? __ldcg(bound + g_begin + gi) : 0u;
                if (kinv != 0ull || pb != 0u) {
                    const float gx = __ldg(gt + (size_t)(g_begin + gi) * 5 + 0);
                    const float gy = __ldg(gt + (size_t)(g_begin + gi) * 5 + 1);
                    float sb = __int_as_float(0x7f800000);
                    if (kinv != 0ull) sb = __uint_as_float((unsigned int)((~kinv) >> 32)) * 1.000001f;
                    if (pb != 0u) sb = fminf(sb, __uint_as_float(pb));
                    const float dx = fmaxf(fmaxf(lo_x - gx, gx - hi_x), 0.f), dy = fmaxf(fmaxf(lo_y - gy, gy - hi_y), 0.f);
                    const float lb2 = (dx * dx + dy * dy) * 0.999999f;
                    need = !(lb2 > fmaf(sb * sb, 1.000001f, 1e-6f * (gx * gx + gy * gy + pn_max)));
                }
            }
#endif
            const unsigned need_mask = __ballot_sync(0xffffffffu, need);
            if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = __popc(need_mask);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kAssignThreads / 32; ++w) {
                before += w < (int)(threadIdx.x >> 5) ? s_cnt[w] : 0;
                total += s_cnt[w];
            }
            if (need) s_list[count + before + __popc(need_mask & ((1u << (threadIdx.x & 31)) - 1u))] = gi;
            count += total;
            g0 += kAssignThreads;
            __syncthreads();
        }
        if (count == 0) break;                              // uniform: nobody (else) needs this tile
        const int m_chunk = min(count, kAssignThreads);
        const int m_pad = (m_chunk + 31) & ~31;
        const int n_slice = kAssignThreads / m_pad;
        const int g = threadIdx.x % m_pad;
        const int slice = threadIdx.x / m_pad;
        unsigned long long key = kNoKey;
        if (g < m_chunk && slice < n_slice) {
            const int gl = s_list[g];
            const float gx = __ldg(gt + (size_t)(g_begin + gl) * 5 + 0);
            const float gy = __ldg(gt + (size_t)(g_begin + gl) * 5 + 1);
            // ATen _euclidean_dist: [-2gx, -2gy, |g|^2, 1] . [px, py, 1, |p|^2], K = 4, FMA chain k = 0..3
            const float c0 = __fmul_rn(-2.f, gx), c1 = __fmul_rn(-2.f, gy);
            const float gn = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
            const f32x2 c0p = pack2(c0, c0), c1p = pack2(c1, c1), gnp = pack2(gn, gn);
            const int per = (((tile_n4 >> 2) + n_slice - 1) / n_slice) << 2;
            const int j_begin = slice * per, j_end = min(tile_n4, j_begin + per);
            // The reference takes argmin of sqrt(clamp_min(d^2, 0)) with the first index winning ties.
            // Both maps are monotone, so the winner is the FIRST anchor whose clamped d^2 is at most
            // hi(m) = the largest float with the same rounded sqrt as the minimum m.  Two cheap passes
            // over the slice instead of index bookkeeping in a divergent loop:
            //   pass 1: m = min d^2 (packed FMA chain + FMNMX only);
            //   pass 2: first j with d^2 <= hi(m) (one compare per four anchors, branch taken once).
            auto d2_of4 = [&](int j, float (&d)[4]) {
                const ulonglong2 xs = *reinterpret_cast<const ulonglong2 *>(s_x + j);
                const ulonglong2 ys = *reinterpret_cast<const ulonglong2 *>(s_y + j);
                const ulonglong2 ps = *reinterpret_cast<const ulonglong2 *>(s_p + j);
                unpack2(add2(add2(fma2(c1p, ys.x, mul2(c0p, xs.x)), gnp), ps.x), d[0], d[1]);
                unpack2(add2(add2(fma2(c1p, ys.y, mul2(c0p, xs.y)), gnp), ps.y), d[2], d[3]);
            };
            float m_raw = __int_as_float(0x7f800000);
#pragma unroll kScanUnroll
            for (int j = j_begin; j < j_end; j += 4) {
                float d[4];
                d2_of4(j, d);
                m_raw = fminf(m_raw, fminf(fminf(d[0], d[1]), fminf(d[2], d[3])));
            }
            int best_j = -1;
            float best_s = __int_as_float(0x7f800000);
            if (m_raw < __int_as_float(0x7f800000)) {
                const float m = fmaxf(m_raw, 0.f);
                best_s = __fsqrt_rn(m);
                float hi = m;                              // at most three floats share a rounded sqrt
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float nx = __uint_as_float(__float_as_uint(hi) + 1u);
                    if (__fsqrt_rn(nx) == best_s) hi = nx;
                }
                for (int j = j_begin; j < j_end && best_j < 0; j += 4) {
                    float d[4];
                    d2_of4(j, d);
                    if (fminf(fminf(d[0], d[1]), fminf(d[2], d[3])) <= hi) {
#pragma unroll
                        for (int v = 3; v >= 0; --v)
                            if (d[v] <= hi) best_j = j + v;  // descending v: the lowest index sticks
                    }
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
            if (k != kNoKey) atomicMax(best + g_begin + s_list[threadIdx.x], ~k);
        }
        const int rest = count - m_chunk;                   // < kAssignThreads survivors still waiting
        const int carry = (int)threadIdx.x < rest ? s_list[m_chunk + threadIdx.x] : 0;
        __syncthreads();
        if ((int)threadIdx.x < rest) s_list[threadIdx.x] = carry;
        count = rest;
        __syncthreads();
    }
}

template <typename T, int VW>
__global__ void __launch_bounds__(kAssignThreads, YB_ASSIGN_MINBLOCKS ? YB_ASSIGN_MINBLOCKS : (VW == 8 ? 6 : 4))
assign_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, const float *__restrict__ anchors,
              const float *__restrict__ strides, const float *__restrict__ gt, const int *__restrict__ gt_off,
              unsigned long long *__restrict__ best, int *__restrict__ gt_img, T *__restrict__ grad) {
    assign_body<T, VW>(blockIdx.y, blockIdx.x, preds, n_ch, n_anchors, anchors, strides, gt, gt_off, best, nullptr, gt_img, grad,
                       true);
}

// ------------------------------------------------------------------------------------------
// match_kernel: one half-warp per GT; its last CTA writes the loss scalars
// ------------------------------------------------------------------------------------------
// Sixteen lanes per GT: lane l of a half-warp holds bin l of all four box sides of the matched anchor.
// What depends on the anchor only (shared by all GTs matched to it):
struct AnchorTerms {
    float z[4];         // this lane's bin of the four sides (raw logits)
    float mx[4], sm[4]; // softmax max / sum of each side
    float pr[4];        // softmax probability of this lane's bin
    float ds[4];        // expected distance of each side
    float ax, ay, s;
    PredBox b;
};
struct GtTerms {
    float g[4];         // gradient of this lane's bin on sides l, t, r, b
    float dfl;          // sum over the four sides of the DFL term (all lanes)
    float iou;          // soft target (all lanes)
    float cell_delta;   // QFL loss change of the (anchor, class) cell per unit of target: q^2 log p - p^2 log q
    float cell_grad0;   // d total / d logit of that cell at target t is  cell_grad0 + t * cell_grad1
    float cell_grad1;
};

// All 32 lanes of the warp must call these together (full-mask shuffles; xor offsets <= 8 stay in a half).
__device__ __forceinline__ void anchor_softmax(AnchorTerms &a) {
    const int bin = threadIdx.x & 15;
    // softmax + expectation of each side over the 16 lanes of the half
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float m = a.z[k];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float e = expf(a.z[k] - m);
        float sum = e;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float p = __fdiv_rn(e, sum);
        float d = p * (float)bin;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        a.mx[k] = m; a.sm[k] = sum; a.pr[k] = p; a.ds[k] = d;
    }
    a.b = decode_box(a.ax, a.ay, a.s, a.ds[0], a.ds[1], a.ds[2], a.ds[3]);
}

// ------------------------------------------------------------------------------------------
// probe role: an upper bound of every GT's nearest-centre distance before any tile has been scanned
// ------------------------------------------------------------------------------------------
// One half-warp per GT.  When the caller describes the anchors as a pyramid of regular grids (yb_anchor_grid, the same
// hint the task-aligned path takes), the cell of each level that holds the GT's centre is a very good guess of where
// the nearest predicted centre will be; the half-warp decodes those few anchors and publishes the smallest distance,
// inflated far beyond any rounding, as the GT's bound.  The box role then drops (GT, tile) pairs from its very first
// tile on -- the tiles of the coarse levels, which used to scan every GT of the image, included.  Nothing depends on
// the hint being right: ANY anchor's distance bounds the minimum from above, so a wrong hint only costs pruning.
template <typename T>
__device__ __forceinline__ void probe_gts(const T *__restrict__ preds, const float *__restrict__ gt, int g0, int gt_total,
                                          const int *__restrict__ gt_off, int n_images, int n_ch, int n_anchors,
                                          const float *__restrict__ anchors, const float *__restrict__ strides,
                                          const yb_anchor_grid &grid, unsigned int *bound) {
    const int lane = threadIdx.x & 31, bin = lane & 15;
    const int g_raw = g0 + (threadIdx.x >> 4);
    if (g0 + ((threadIdx.x & ~31) >> 4) >= gt_total) return;          // warp-uniform
    const bool live = g_raw < gt_total;
    const int g = live ? g_raw : gt_total - 1;
    int lo = 0, hi = n_images;                                         // image of the GT
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(gt_off + mid) <= g) lo = mid; else hi = mid;
    }
    const int n = lo;
    const float gx = __ldg(gt + (size_t)g * 5), gy = __ldg(gt + (size_t)g * 5 + 1);
    float best_d = __int_as_float(0x7f800000);
    // Only the YB_PROBE_LEVELS finest levels are probed (0: all).  Each probed anchor costs 64 scattered 32-byte sectors, every
    // GT probes at the head of the launch -- a burst of random DRAM accesses during which nothing streams -- and the finest
    // level's home cell nearly always holds the smallest distance (cfg5: 163.4 -> 157.5 us with one level, 158.5 with two)
    const int n_levels = YB_PROBE_LEVELS > 0 ? min(grid.n_levels, YB_PROBE_LEVELS) : grid.n_levels;
    for (int l0 = 0; l0 < n_levels; l0 += 4) {             // the gathers of four levels fly together
        float z[4][4];
        int a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int l = min(l0 + u, n_levels - 1);       // (the tail repeats the last level's index and loads nothing)
            const float s = grid.stride[l];
            const int col = min(max((int)floorf(gx / s - grid.x0[l] + 0.5f), 0), grid.w[l] - 1);
            const int row = min(max((int)floorf(gy / s - grid.y0[l] + 0.5f), 0), grid.h[l] - 1);
            a[u] = min(max(grid.start[l] + row * grid.w[l] + col, 0), n_anchors - 1);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                z[u][k] = l0 + u < n_levels ? load_as_float(preds + ((size_t)n * n_ch + k * kRegMax + bin) * n_anchors + a[u]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (l0 + u >= n_levels) break;                 // warp-uniform
            AnchorTerms at;
#pragma unroll
            for (int k = 0; k < 4; ++k) at.z[k] = z[u][k];
            at.ax = __ldg(anchors + a[u]); at.ay = __ldg(anchors + n_anchors + a[u]); at.s = __ldg(strides + a[u]);
            anchor_softmax(at);
            const float dx = at.b.cx - gx, dy = at.b.cy - gy;
            best_d = fminf(best_d, sqrtf(dx * dx + dy * dy));
        }
    }
    // the decode here and the box role's differ by ~1e-6 relative, the matmul-form d^2 by ~1e-6 (|g|^2 + |p|^2):
    // 0.1 % + 0.05 px is orders of magnitude more; a non-finite distance publishes nothing
    const float b = fmaf(best_d, 1.001f, 0.05f);
    if (live && bin == 0 && b < __int_as_float(0x7f800000)) bound[g] = __float_as_uint(b);
}

__device__ __forceinline__ GtTerms gt_terms(const AnchorTerms &a, float z_cls, float gcx, float gcy, float gw, float gh,
                                            float k_dfl, float k_cls) {
    const int lane = threadIdx.x & 31;
    const int bin = lane & 15;
    const int base = lane & 16;
    const PredBox &b = a.b;
    const float ax = a.ax, ay = a.ay, s = a.s;

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

    // ---- the (anchor, class) QFL cell (src/model/losses.py:51-56) ------------------------------
    // loss(t) = t q^2 log(p+e) + (1-t) p^2 log(q+e)   (sign and 1/A applied later); linear in t, so
    // d total / d iou does not depend on the target value (and reaches over-written GTs too)
    const float pc = __fdiv_rn(1.f, 1.f + expf(-z_cls));
    const float qc = 1.f - pc;
    const float lp = logf(pc + kEpsLog), lq = logf(qc + kEpsLog);
    const float cell_delta = qc * qc * lp - pc * pc * lq;
    const float g_iou = -k_cls * cell_delta;
    const float dpos = -2.f * qc * lp + qc * qc / (pc + kEpsLog);     // d/dp of q^2 log(p+e)
    const float dneg = 2.f * pc * lq - pc * pc / (qc + kEpsLog);      // d/dp of p^2 log(q+e)

    // backward of the IoU in autograd's conventions: clamp passes where the input >= 0, max/min
    // route to the larger/smaller argument and split evenly on ties.
    const float d_inter = g_iou * (1.f / den + inter / (den * den));   // incl. the -inter inside the union
    const float d_area1 = -g_iou * inter / (den * den);
    const float d_iw = (iw_raw >= 0.f) ? d_inter * ih : 0.f;
    const float d_ih = (ih_raw >= 0.f) ? d_inter * iw : 0.f;
    auto pick_max = [](float x, float o) { return x > o ? 1.f : (x == o ? 0.5f : 0.f); };   // weight on `x`
    auto pick_min = [](float x, float o) { return x < o ? 1.f : (x == o ? 0.5f : 0.f); };
    const float d_ax1 = -d_iw * pick_max(ax1, bx1) - d_area1 * ah;
    const float d_ax2 = d_iw * pick_min(ax2, bx2) + d_area1 * ah;
    const float d_ay1 = -d_ih * pick_max(ay1, by1) - d_area1 * aw;
    const float d_ay2 = d_ih * pick_min(ay2, by2) + d_area1 * aw;
    // ax1 = cx - w/2, ax2 = cx + w/2, ay1 = cy - h/2, ay2 = h + cy/2
    const float d_cx = d_ax1 + d_ax2;
    const float d_w = 0.5f * (d_ax2 - d_ax1);
    const float d_cy = d_ay1 + 0.5f * d_ay2;
    const float d_h = d_ay2 - 0.5f * d_ay1;
    // cx = (x1+x2)/2, w = x2-x1 ;  x1 = (ax-dl)*s, x2 = (ax+dr)*s
    const float d_x1 = 0.5f * d_cx - d_w, d_x2 = 0.5f * d_cx + d_w;
    const float d_y1 = 0.5f * d_cy - d_h, d_y2 = 0.5f * d_cy + d_h;
    const float dd[4] = {-d_x1 * s, -d_y1 * s, d_x2 * s, d_y2 * s};   // d total / d (dl, dt, dr, db) through the IoU

    // ---- DFL target bins (src/model/losses.py:226-246) and loss rows (:63-78) -----------------
    const float tgt[4] = {ax - bx1 / s, ay - by1 / s, bx2 / s - ax, by2 / s - ay};
    const float hi_clamp = (float)(kRegMax - 1 - 0.01);
    GtTerms r;
    float dfl = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float t = fminf(fmaxf(tgt[k], 0.f), hi_clamp);
        const int bl = (int)t;
        const float wl = (float)(bl + 1) - t, wr = t - (float)bl;
        const float lpk = (a.z[k] - a.mx[k]) - logf(a.sm[k]);         // log-softmax of this lane's bin
        dfl -= __shfl_sync(0xffffffffu, lpk, base + bl) * wl + __shfl_sync(0xffffffffu, lpk, base + bl + 1) * wr;
        const float oh = (bin == bl ? wl : 0.f) + (bin == bl + 1 ? wr : 0.f);
        r.g[k] = k_dfl * ((wl + wr) * a.pr[k] - oh) + dd[k] * a.pr[k] * ((float)bin - a.ds[k]);
    }
    r.dfl = dfl;
    r.iou = iou;
    r.cell_delta = cell_delta;
    r.cell_grad0 = -k_cls * dneg * pc * qc;
    r.cell_grad1 = -k_cls * (dpos - dneg) * pc * qc;
    return r;
}

constexpr int kMatchThreads = 128;

// One HALF-warp per GT; all 32 lanes of a warp call this together (the two halves may serve different images).
// `m` is the GT's index within image n (of m_img > 0 GTs starting at row g_begin); an idle half (live == false) walks
// through the same shuffles on the image's last GT and writes nothing.
template <typename T>
__device__ __forceinline__ void match_gt(const T *__restrict__ preds, int n, int g_begin, int m_img, int m, bool live, int n_ch,
                                         int n_anchors, int nc, const float *__restrict__ anchors,
                                         const float *__restrict__ strides, const float *__restrict__ gt,
                                         const unsigned long long *best, float k_dfl_num, float k_cls, T *grad,
                                         bool whole_sectors, unsigned long long *acc, unsigned int *ticket,
                                         int *__restrict__ out_idx, float *__restrict__ out_iou) {
    constexpr int SEC = 32 / (int)sizeof(T);               // anchors per 32-byte sector of a gradient row
    const int bin = threadIdx.x & 15;
    const int g = g_begin + m;
    const T *img = preds + (size_t)n * n_ch * n_anchors;
    auto idx_of = [&](int mm) {
        // a plain (L1-cached) load: every CTA of the image scans the same few hundred keys.  Final by now, and ordered
        // after the producers' atomics by the acquire of dep_wait (or by the kernel boundary); never the .nc path,
        // which is only defined for data nobody writes during the launch
        const unsigned long long inv = best[g_begin + mm];
        return inv == 0ull ? 0 : (int)(unsigned int)(~inv & 0xffffffffull);   // no finite distance -> anchor 0
    };
    const int idx = idx_of(m);
    const float k_dfl = k_dfl_num / (float)m_img;          // lambda_dfl / (N * 4 * M)

    // the gathers (DRAM latency) are issued first; the duplicate search below runs in their shadow
    AnchorTerms at;
#pragma unroll
    for (int k = 0; k < 4; ++k) at.z[k] = load_as_float(img + (size_t)(k * kRegMax + bin) * n_anchors + idx);
    const float *g5 = gt + (size_t)g * 5;
    const float gcx = __ldg(g5 + 0), gcy = __ldg(g5 + 1), gw = __ldg(g5 + 2), gh = __ldg(g5 + 3);
    const int cls_raw = (int)__ldg(g5 + 4);                                        // .long(): truncation
    const int cls = min(max(cls_raw, 0), nc - 1);          // memory-safe; the violation itself is reported below
    const float z_cls = load_as_float(img + (size_t)(4 * kRegMax + cls) * n_anchors + idx);
    at.ax = __ldg(anchors + idx); at.ay = __ldg(anchors + n_anchors + idx); at.s = __ldg(strides + idx);

    // Which GTs of this image share my anchor?  owner = lowest such m (writes the summed gradient),
    // winner = highest (its class row is the anchor's QFL target: "last write wins", losses.py:261).
    // ... and does another GT's anchor share a 32-byte sector of the gradient rows with mine (a sector mate)?
    int first = m, last = m, n_later = 0, mate = 0;
    for (int mm = bin; mm < m_img; mm += 16) {
        const int o = idx_of(mm);
        if (o == idx) {
            first = min(first, mm);
            last = max(last, mm);
            n_later += mm > m ? 1 : 0;
        } else if (o / SEC == idx / SEC) {
            mate = 1;
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        n_later += __shfl_xor_sync(0xffffffffu, n_later, o);
        mate |= __shfl_xor_sync(0xffffffffu, mate, o);
    }

    anchor_softmax(at);
    const GtTerms mine = gt_terms(at, z_cls, gcx, gcy, gw, gh, k_dfl, k_cls);
    const bool winner = (last == m);
    T *gimg = grad ? grad + (size_t)n * n_ch * n_anchors : nullptr;
    if (live && bin == 0) {
        unsigned long long *a_img = acc + (size_t)n * kAccPerImage;
        acc_add(a_img + kAccDfl, mine.dfl, ticket);
        if (winner) {
            acc_add(a_img + kAccDcls, mine.iou * mine.cell_delta, ticket);
            atomicAdd(a_img + kAccWin, 1ull);
        }
        if (cls_raw != cls) atomicAdd(ticket + 2, 1u);     // the reference's scatter_ raises here (losses.py:260)
        if (out_idx) out_idx[g] = idx;
        if (out_iou) out_iou[g] = mine.iou;
        if (winner && gimg) {
            // the anchor's one positive QFL cell: overwrite the target-0 gradient the class pass wrote
            store_from_float(gimg + (size_t)(4 * kRegMax + cls) * n_anchors + idx,
                             mine.cell_grad0 + mine.iou * mine.cell_grad1);
        }
    }
    // the owner adds the terms of the later GTs that share its anchor (rare); the trip count is made
    // warp-uniform because gt_terms shuffles across the whole warp
    float sum[4] = {mine.g[0], mine.g[1], mine.g[2], mine.g[3]};
    const bool owner = live && first == m && gimg != nullptr;
    int todo = owner ? n_later : 0;
    const int trips = max(todo, __shfl_xor_sync(0xffffffffu, todo, 16));
    int cur = m;
    for (int it = 0; it < trips; ++it) {
        int nxt = 0x7fffffff;
        if (it < todo)
            for (int mm = bin; mm < m_img; mm += 16)
                if (mm > cur && idx_of(mm) == idx) nxt = min(nxt, mm);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) nxt = min(nxt, __shfl_xor_sync(0xffffffffu, nxt, o));
        const bool has = nxt != 0x7fffffff;
        const float *o5 = gt + (size_t)(g_begin + (has ? nxt : m)) * 5;
        const int ocls = min(max((int)__ldg(o5 + 4), 0), nc - 1);
        const float oz = load_as_float(img + (size_t)(4 * kRegMax + ocls) * n_anchors + idx);
        const GtTerms o = gt_terms(at, oz, __ldg(o5 + 0), __ldg(o5 + 1), __ldg(o5 + 2), __ldg(o5 + 3), k_dfl, k_cls);
        if (has) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sum[k] += o.g[k];
            cur = nxt;
        }
    }
    if (owner) {
        // The box role left zeros in these rows, so every neighbour of the anchor inside its 32-byte sector is known to be
        // zero unless a sector mate exists: the whole sector is written with one 256-bit store and DRAM never has to be
        // read to merge a 4-byte write into it (the read-modify-write was half of this step's scattered DRAM traffic).
        if (whole_sectors && !mate) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                const int pos = idx % SEC;
                if (sizeof(T) == 4) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = i == pos ? __float_as_uint(sum[k]) : 0u;
                } else {
                    const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(sum[k])) << ((pos & 1) * 16);
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = i == (pos >> 1) ? h : 0u;
                }
                T *p = gimg + (size_t)(k * kRegMax + bin) * n_anchors + (idx - pos);
                asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                             "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                             : "memory");
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) store_from_float(gimg + (size_t)(k * kRegMax + bin) * n_anchors + idx, sum[k]);
        }
    }
}

// per-image terms, then the batch means, in a fixed order -> 8 loss scalars.  One whole CTA of kMatchThreads threads,
// after every accumulator of the call is final.
__device__ __forceinline__ void final_reduce(int n_images, int n_anchors, const int *__restrict__ gt_off,
                                             const unsigned long long *acc, const unsigned int *ticket, float lambda_cls,
                                             float lambda_dfl, float *__restrict__ out_loss,
                                             float *__restrict__ out_per_image) {
    __shared__ double s_d[kMatchThreads / 32], s_c[kMatchThreads / 32], s_f[kMatchThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double d = 0.0, c = 0.0, f = 0.0;
    for (int b = threadIdx.x; b < n_images; b += kMatchThreads) {
        const unsigned long long *a_img = acc + (size_t)b * kAccPerImage;
        double c_img = acc_get(a_img + kAccDcls);
#pragma unroll
        for (int k = 0; k < kAccCls; ++k) c_img += acc_get(a_img + k);
        const double d_img = acc_get(a_img + kAccDfl);
        const int mb = __ldg(gt_off + b + 1) - __ldg(gt_off + b);
        const float cls_b = (float)(-c_img / (double)n_anchors);                     // losses.py:56
        const float dfl_b = mb > 0 ? (float)(d_img / (4.0 * (double)mb)) : 0.f;      // losses.py:78, :252
        if (out_per_image) {
            out_per_image[b] = dfl_b;
            out_per_image[n_images + b] = cls_b;
        }
        d += (double)dfl_b;
        c += (double)cls_b;
        f += (double)__ldcg(a_img + kAccWin);
    }
    d = warp_sum_d(d); c = warp_sum_d(c); f = warp_sum_d(f);
    if (lane == 0) { s_d[warp] = d; s_c[warp] = c; s_f[warp] = f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double dd = 0.0, cc = 0.0, ff = 0.0;
        for (int w = 0; w < kMatchThreads / 32; ++w) { dd += s_d[w]; cc += s_c[w]; ff += s_f[w]; }
        float mean_dfl = (float)(dd / (double)n_images);              // N counts images without GT (losses.py:271)
        float mean_cls = (float)(cc / (double)n_images);
        if (__ldcg(ticket + 1) != 0u) mean_cls = __int_as_float(0x7fc00000);   // a non-finite partial: NaN, as the reference
        const unsigned int stalled = __ldcg(ticket + 3);              // a consumer CTA gave up waiting (common.cuh::dep_wait)
        out_loss[0] = stalled ? __int_as_float(0x7fc00000) : lambda_dfl * mean_dfl + lambda_cls * mean_cls;  // losses.py:275
        out_loss[1] = mean_dfl;
        out_loss[2] = mean_cls;
        out_loss[3] = (float)ff;
        out_loss[4] = out_loss[5] = 0.f;
        out_loss[6] = (float)stalled;
        out_loss[7] = (float)__ldcg(ticket + 2);                      // GT rows whose class id is outside [0, nc)
    }
}

// the split launch's third kernel (YB_LOSS_SPLIT_LAUNCH; the fused launch runs the same code as a role of
// fused_main_kernel): 8 GTs per CTA; the grid has at least one CTA even without GTs, because the last CTA to finish
// (ticket) writes the loss scalars
template <typename T>
__global__ void __launch_bounds__(kMatchThreads, 6)
match_kernel(const T *__restrict__ preds, int n_images, int n_ch, int n_anchors, int nc,
             const float *__restrict__ anchors, const float *__restrict__ strides, const float *__restrict__ gt,
             const int *__restrict__ gt_off, const int *__restrict__ gt_img, int gt_total,
             const unsigned long long *best, float k_dfl_num, float k_cls, float lambda_cls, float lambda_dfl,
             T *__restrict__ grad, int whole_sectors, unsigned long long *__restrict__ acc, unsigned int *__restrict__ ticket,
             int *__restrict__ out_idx, float *__restrict__ out_iou, float *__restrict__ out_loss,
             float *__restrict__ out_per_image) {
    const int g_raw = (blockIdx.x * kMatchThreads + threadIdx.x) >> 4;
    const bool warp_has_work = ((blockIdx.x * kMatchThreads + (threadIdx.x & ~31)) >> 4) < gt_total;
    if (warp_has_work) {                                       // warp-uniform
        const bool live = g_raw < gt_total;                    // the second half of the last warp may be idle
        const int g = live ? g_raw : gt_total - 1;             // ... but must walk through the same shuffles
        const int n = __ldg(gt_img + g);
        const int g_begin = __ldg(gt_off + n);
        const int m_img = __ldg(gt_off + n + 1) - g_begin;
        match_gt<T>(preds, n, g_begin, m_img, g - g_begin, live, n_ch, n_anchors, nc, anchors, strides, gt, best, k_dfl_num,
                    k_cls, grad, whole_sectors != 0, acc, ticket, out_idx, out_iou);
    }
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    final_reduce(n_images, n_anchors, gt_off, acc, ticket, lambda_cls, lambda_dfl, out_loss, out_per_image);
}

// ------------------------------------------------------------------------------------------
// cls_loss_kernel
// ------------------------------------------------------------------------------------------
// One element of the quality focal loss at target 0 and its gradient w.r.t. the logit
// (src/model/losses.py:51-56 with target_scores == 0):
//     loss term  p^2 log(1 - p + 1e-12)          (sign and 1/A applied by the reduction)
//     gradient   k p^2 (p q / (q + 1e-12) - 2 q log(q + 1e-12)),   k = lambda_cls / (N A),  q = fl(1 - p)
// One formula for ANY logit, two elements at a time on the packed fp32 pipe (FMUL2 / FFMA2 / FADD2):
//   p        MUFU.EX2 + MUFU.RCP; q is formed as fl(1 - p) exactly as the reference does, so its rounding near 1 (and its
//            collapse to 0 for logits above ~16.6, where the reference's gradient vanishes) is reproduced, not avoided;
//   log q    q >= 0.7071 (logit <= -0.88, the background): degree-5 minimax polynomial of log1p on [-0.293, 0] (max
//            relative error 9.4e-8, fitted offline) -- MUFU.LG2's 2^-22 ABSOLUTE error would be a 1e-5 RELATIVE error
//            on these, the cells that carry the loss; the +1e-12 is below half an ulp of q there and drops out;
//            q < 0.7071: MUFU.LG2 of q + 1e-12, whose absolute error is harmless next to |log q| >= 0.35;
//   q / (q + 1e-12) is 1 to within 1.7e-5 for every non-zero float q (>= 2^-24) and is taken as 1; q == 0 zeroes the gradient.
// (Round 1 took a non-inlined scalar expf / logf path for every group of four with a logit above -0.88: 0.47 ms per step
// instead of 0.26 on class logits ~ N(0, 2).)
constexpr float kFastQ = 0.70710678f;

__device__ __forceinline__ float fast_lg2(float x) {         // MUFU.LG2
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// SHORT: bf16 head outputs carry 8 bits of mantissa and are held to 1e-2 relative; their log q on the polynomial branch
// uses a degree-2 R(f) (max relative error 1.2e-4, fitted offline) instead of the degree-5 one: three packed FMAs per
// two cells less in a kernel that is bound by instruction issue at bf16.
struct QflPair {                                              // two cells between the two halves of qfl_bg_group
    f32x2 p, q, lq;
};

template <bool SHORT>
__device__ __forceinline__ QflPair qfl_bg_front(float x0, float x1) {
    const f32x2 one = pack2(1.f, 1.f), mone = pack2(-1.f, -1.f);
    float a0, a1;
    unpack2(mul2(pack2(x0, x1), pack2(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
    const f32x2 u = add2(pack2(fast_ex2(a0), fast_ex2(a1)), one);            // 1 + exp(-x)
    float u0, u1;
    unpack2(u, u0, u1);
    QflPair c;
    c.p = pack2(fast_rcp(u0), fast_rcp(u1));
    c.q = fma2(c.p, mone, one);                                               // fl(1 - p), as the reference
    // log q, polynomial branch: f = q - 1 is exact for q in [0.5, 1]
    const f32x2 f = add2(c.q, mone);
    f32x2 r;
    if (SHORT) {
        r = pack2(-0.37099915742874146f, -0.37099915742874146f);
        r = fma2(r, f, pack2(0.3176640272140503f, 0.3176640272140503f));
        r = fma2(r, f, pack2(-0.5004022717475891f, -0.5004022717475891f));
    } else {
        r = pack2(0.3410167098045349f, 0.3410167098045349f);
        r = fma2(r, f, pack2(-0.08926734328269958f, -0.08926734328269958f));
        r = fma2(r, f, pack2(0.21280372142791748f, 0.21280372142791748f));
        r = fma2(r, f, pack2(-0.249073788523674f, -0.249073788523674f));
        r = fma2(r, f, pack2(0.33335742354393005f, 0.33335742354393005f));
        r = fma2(r, f, pack2(-0.49999991059303284f, -0.49999991059303284f));
    }
    c.lq = fma2(mul2(f, f), r, f);                                            // log(q) = f + f^2 R(f)
    return c;
}

// the cells the polynomial does not cover (q < 0.7071: logit above -0.88): MUFU.LG2; and q == 0 (logit above ~16.6), where
// sigmoid'(x) = p q = 0: the reference's gradient vanishes while its loss term p^2 log(1e-12) stays -- that term is added
// here and the cell's p zeroed, which zeroes both of its later products (NaN stays NaN: no comparison below holds for it)
__device__ __forceinline__ void qfl_bg_fix(QflPair &c, f32x2 &acc2) {
    float q0, q1, l0, l1, p0, p1;
    unpack2(c.q, q0, q1);
    unpack2(c.lq, l0, l1);
    unpack2(c.p, p0, p1);
    const float lm0 = fast_lg2(q0 + kEpsLog) * 0.6931471805599453f, lm1 = fast_lg2(q1 + kEpsLog) * 0.6931471805599453f;
    l0 = q0 >= kFastQ ? l0 : lm0;
    l1 = q1 >= kFastQ ? l1 : lm1;
    const float t0 = q0 == 0.f ? p0 * p0 * lm0 : 0.f, t1 = q1 == 0.f ? p1 * p1 * lm1 : 0.f;
    acc2 = add2(acc2, pack2(t0, t1));
    p0 = q0 == 0.f ? 0.f : p0;
    p1 = q1 == 0.f ? 0.f : p1;
    c.lq = pack2(l0, l1);
    c.p = pack2(p0, p1);
}

__device__ __forceinline__ void qfl_bg_back(const QflPair &c, f32x2 k2, f32x2 &acc2, float &g0, float &g1) {
    const f32x2 mtwo = pack2(-2.f, -2.f);
    const f32x2 p2 = mul2(c.p, c.p);
    acc2 = fma2(p2, c.lq, acc2);
    unpack2(mul2(mul2(p2, k2), fma2(mul2(c.q, mtwo), c.lq, c.p)), g0, g1);    // k p^2 (p - 2 q log q)
}

// One row of a thread's anchors.  The polynomial serves every cell first; ONE test per row (the smallest q of the row, a
// per-lane branch the background never takes) sends the row through qfl_bg_fix -- per pair of cells that test cost a
// compare, a branch with its convergence barrier and two selects for q == 0, a fifth of the class role's instructions.
template <typename T, int VW>
__device__ __forceinline__ void qfl_bg_group(const Group<T, VW> &row, f32x2 k2, f32x2 &acc2, float (&g)[VW]) {
    constexpr bool SHORT = YB_BF16_SHORT_LOG && sizeof(T) == 2;
    if constexpr (VW == 1) {
        float g1;
        f32x2 pair = pack2(0.f, 0.f);
        QflPair c = qfl_bg_front<SHORT>(row.get(0), row.get(0));
        float q0, q1;
        unpack2(c.q, q0, q1);
        if (!(q0 >= kFastQ)) qfl_bg_fix(c, pair);
        qfl_bg_back(c, k2, pair, g[0], g1);
        float lo, hi;
        unpack2(pair, lo, hi);
        acc2 = add2(acc2, pack2(lo, 0.f));                 // the second lane is a copy: count it once
    } else {
        QflPair c[VW / 2];
        float qmin = __int_as_float(0x7f800000);
#pragma unroll
        for (int v = 0; v < VW; v += 2) {
            c[v / 2] = qfl_bg_front<SHORT>(row.get(v), row.get(v + 1));
            float q0, q1;
            unpack2(c[v / 2].q, q0, q1);
            qmin = fminf(qmin, fminf(q0, q1));             // (a NaN q drops out of the minimum: its cell stays NaN either way)
        }
        if (qmin < kFastQ) {
#pragma unroll
            for (int v = 0; v < VW; v += 2) qfl_bg_fix(c[v / 2], acc2);
        }
#pragma unroll
        for (int v = 0; v < VW; v += 2) qfl_bg_back(c[v / 2], k2, acc2, g[v], g[v + 1]);
    }
}

// the class channels of a tile are split over n_split CTAs: short-lived CTAs, small tail
template <typename T, int VW, bool WRITE_GRAD>
__device__ __forceinline__ void cls_body(int n, int tile, int n_tiles, int split, int n_split, const T *__restrict__ preds,
                                         int n_ch, int n_anchors, int nc, float k_cls, T *__restrict__ grad,
                                         unsigned long long *__restrict__ acc, unsigned int *__restrict__ flags) {
    __shared__ float s_red[kClsThreads / 32];
    const int a0 = (tile * kClsThreads + threadIdx.x) * VW;
    const int c_per = (nc + n_split - 1) / n_split;
    const int c_lo = split * c_per;
    nc = min(nc, c_lo + c_per) - c_lo;
    f32x2 acc2 = pack2(0.f, 0.f);
    if (a0 < n_anchors && nc > 0) {
        const size_t base = ((size_t)n * n_ch + 4 * kRegMax + c_lo) * n_anchors + a0;
        const f32x2 k2 = pack2(k_cls, k_cls);
        constexpr int U = sizeof(T) == 4 ? YB_CLS_UNROLL_F32 : YB_CLS_UNROLL;
        // software pipeline: the next U rows are in flight while the current U are evaluated
        Group<T, VW> cur[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (u < nc) cur[u].load(preds + base + (size_t)u * n_anchors);
        for (int c = 0; c < nc; c += U) {
            Group<T, VW> nxt[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c + U + u < nc) nxt[u].load(preds + base + (size_t)(c + U + u) * n_anchors);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (c + u < nc) {
                    float g[VW];
                    qfl_bg_group<T, VW>(cur[u], k2, acc2, g);
                    if (WRITE_GRAD) Group<T, VW>::store(grad + base + (size_t)(c + u) * n_anchors, g);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) cur[u] = nxt[u];
        }
    }
    float lo, hi;
    unpack2(acc2, lo, hi);
    const float wsum = warp_sum(lo + hi);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kClsThreads / 32; ++w) s += s_red[w];
        // per-image fixed-point accumulator; consecutive CTAs use different sub-accumulators
        acc_add(acc + (size_t)n * kAccPerImage + ((tile * n_split + split) & (kAccCls - 1)), s, flags);
    }
}

template <typename T, int VW, bool WRITE_GRAD>
__global__ void __launch_bounds__(kClsThreads, YB_CLS_MINBLOCKS)
cls_loss_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, int nc, float k_cls, T *__restrict__ grad,
                unsigned long long *__restrict__ acc, unsigned int *__restrict__ flags) {
    cls_body<T, VW, WRITE_GRAD>(blockIdx.y, blockIdx.x, gridDim.x, blockIdx.z, gridDim.z, preds, n_ch, n_anchors, nc, k_cls,
                                grad, acc, flags);
}

// One launch for the whole step.  Its CTAs take one of four roles:
//   box role    (assign_body)  } consecutive CTAs alternate between the two (YB_CLS_CSPLIT class CTAs per box CTA), so every
//   class role  (cls_body)     } SM holds a mix of the latency-bound decode/scan CTAs and the streaming class CTAs
//   match role  (match_gt)     the per-GT terms of ONE image, once every box and class CTA of that image has counted itself
//                              off (common.cuh: in-kernel dependencies)
//   reducer     (final_reduce) the grid's last CTA: waits for every image, then writes the loss scalars.
// Block order: [probe CTAs] [skew box CTAs] [coarse tiles, image-major] [fine tiles, image-major] [match CTAs, image-major]
// [reducer].  (probe role: probe_gts above; its bounds are read as they come, like `best`: exact whatever the timing.)
// The match CTAs become resident while the last tiles are still streaming and start on the early images at once: the
// step no longer pays a kernel boundary, a second launch and a reduction launch for them (36 us -> 21 us of a 255 us
// step at cfg2).  Measured and not kept: match CTAs of image i placed right behind the tiles of image i + lag, so that
// they run UNDER the streaming roles -- 250 / 244 / 240 us per step for a lag of 1 / 4 / all machine-loads of CTAs: the
// scattered sector reads and writes of the match role cost the same DRAM time wherever they run, and more when they
// break into the streams' open rows.
// zeros over `n` 64-bit words by one CTA of kAssignThreads threads (the arrays of the workspace are 64-byte aligned and padded)
__device__ __forceinline__ void wipe_words(unsigned long long *p, size_t n) {
    for (size_t i = threadIdx.x; i < n; i += kAssignThreads) p[i] = 0ull;
}

struct FusedPlan {
    int n_tiles, coarse, skew;     // coarse < 0: pruning (and the coarse-first order) off
    int match_ctas;                // match CTAs per image (0: no GT in the whole batch); < 0: match_kernel is launched separately
                                   // (no match / reducer CTAs here and nobody counts itself off)
    int whole_sectors;             // gradient rows may be patched with whole 32-byte sectors (match_gt)
    int probe_ctas;                // CTAs of the probe role at the head of the grid (0: no grid hint)
    int gt_total;
    int self_clean;                // YB_LOSS_WS_CLEAN: the reducer leaves every workspace word the launch touched at zero again
};
#ifdef YB_LOSS_TRACE                 // measurement aid (scratch/trace_loss.py): when and where every CTA of the launch ran
__device__ ulonglong4 g_trace[1 << 16];
struct TraceScope {
    unsigned long long t0;
    int role;
    __device__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ TraceScope() : t0(now()), role(-1) {}
    __device__ ~TraceScope() {
        unsigned int sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        if (threadIdx.x == 0 && blockIdx.x < (1u << 16)) g_trace[blockIdx.x] = make_ulonglong4(t0, now(), sm, (unsigned long long)(role + 1));
    }
};
extern "C" int yb_trace_dump(void *dst) { return (int)cudaMemcpyFromSymbol(dst, g_trace, sizeof(g_trace)); }
#define YB_TRACE_ROLE(r) trace_scope.role = (r)
#else
#define YB_TRACE_ROLE(r)
#endif
#ifndef YB_FUSED_MINBLOCKS           // measured on B200: 6 resident CTAs/SM is best for fp32 rows, 5 for bf16 rows
#define YB_FUSED_MINBLOCKS 0
#endif
template <typename T, int VW, bool WRITE_GRAD>
__global__ void __launch_bounds__(kAssignThreads, YB_FUSED_MINBLOCKS ? YB_FUSED_MINBLOCKS : (VW == 8 ? 5 : 6))
fused_main_kernel(const T *__restrict__ preds, int n_images, int n_ch, int n_anchors, int nc, const FusedPlan plan,
                  const yb_anchor_grid grid, const float *__restrict__ anchors, const float *__restrict__ strides,
                  const float *__restrict__ gt, const int *__restrict__ gt_off, unsigned long long *best,
                  unsigned int *bound, int *__restrict__ gt_img, float k_cls,
                  float k_dfl_num, float lambda_cls, float lambda_dfl, T *grad, unsigned long long *acc,
                  unsigned int *flags, unsigned int *done, int *__restrict__ out_idx, float *__restrict__ out_iou,
                  float *__restrict__ out_loss, float *__restrict__ out_per_image) {
    static_assert(kAssignThreads == kClsThreads && kAssignThreads == kMatchThreads, "roles share one block shape");
#if YB_LOSS_PDL
    pdl_launch_dependents();
    pdl_wait();
#endif
#ifdef YB_LOSS_TRACE
    TraceScope trace_scope;
#endif
    constexpr int ROLES = 1 + YB_CLS_CSPLIT;
    const int n_tiles = plan.n_tiles, skew = plan.skew;
    const bool prune = plan.coarse >= 0;                   // coarse < 0: YB_LOSS_NO_PRUNE (exactness tests)
    const int coarse = max(plan.coarse, 0);
    const int n_groups = n_tiles * n_images;
    const unsigned int tile_ctas = (unsigned int)(n_tiles * ROLES);      // box + class CTAs of one image
    const bool chained = plan.match_ctas >= 0;             // the match role and the reducer are part of this launch
    // A tile's box CTA lives about three times as long as one of its class CTAs, so it is launched `skew` tile
    // groups AHEAD of them: the first `skew` blocks are the box CTAs of the first groups, and the last groups of
    // the grid carry class CTAs only.
    int group = -1, role = 0, match_image = -1, match_j = 0;
    bool reducer = false;
    {
        int id = blockIdx.x;
        if (id < plan.probe_ctas) {                        // probe role: the grid's first CTAs, 8 GTs each
            YB_TRACE_ROLE(10);
            probe_gts<T>(preds, gt, id * (kAssignThreads / 16), plan.gt_total, gt_off, n_images, n_ch, n_anchors, anchors, strides,
                         grid, bound);
            if (plan.self_clean) {                         // the reducer may only wipe `bound` once nobody writes it any more
                __syncthreads();
                if (threadIdx.x == 0) dep_signal(flags + 4);
            }
            return;
        }
        id -= plan.probe_ctas;
        if (id < skew) {
            group = id;
        } else {
            id -= skew;
            if (id < n_groups * ROLES) {
                group = id / ROLES;
                role = id - group * ROLES;
                if (role == 0) {
                    group += skew;
                    if (group >= n_groups) return;         // the box CTAs of the last groups went out earlier
                }
            } else {
                id -= n_groups * ROLES;                    // only reached when chained
                if (id < n_images * plan.match_ctas) {
                    match_image = id / plan.match_ctas;
                    match_j = id - match_image * plan.match_ctas;
                } else {
                    reducer = true;
                }
            }
        }
    }
    if (reducer) {
        YB_TRACE_ROLE(12);
        for (int b = threadIdx.x; b < n_images; b += kAssignThreads)
            dep_wait(done + b, tile_ctas + (unsigned int)plan.match_ctas, flags + 3, kLossPollNs);
        // (self-cleaning) the probe CTAs' counter may only be wiped once they have all counted themselves off
        if (plan.self_clean && plan.probe_ctas > 0 && threadIdx.x == 0) dep_wait(flags + 4, (unsigned int)plan.probe_ctas, flags + 3, kLossPollNs);
        __syncthreads();
#ifdef YB_LOSS_TRACE
        const unsigned long long t_waited = TraceScope::now();
#endif
        final_reduce(n_images, n_anchors, gt_off, acc, flags, lambda_cls, lambda_dfl, out_loss, out_per_image);
#ifdef YB_LOSS_TRACE
        if (threadIdx.x == 0) g_trace[(1 << 16) - 1] = make_ulonglong4(t_waited, TraceScope::now(), 0ull, 99ull);
#endif
        if (plan.self_clean) {
            // Every other CTA of the launch that touches the workspace has counted itself off: put back the zeros the next
            // call expects (the words this launch used, no more), so that the step needs no memset node in front of it.
            __syncthreads();                               // ... and final_reduce has read the sums and the flags
            wipe_words(reinterpret_cast<unsigned long long *>(flags), 8);
            wipe_words(reinterpret_cast<unsigned long long *>(done), ((size_t)n_images + 1) / 2);
            wipe_words(acc, (size_t)n_images * kAccPerImage);
        }
        return;
    }
    int image;
    if (match_image >= 0) {
        YB_TRACE_ROLE(11);
        image = match_image;
        const int g_begin = __ldg(gt_off + image), m_img = __ldg(gt_off + image + 1) - g_begin;
        if (m_img > 0) {                                    // uniform per CTA
            if (threadIdx.x == 0) dep_wait(done + image, tile_ctas, flags + 3, kLossPollNs);
            __syncthreads();
            // 8 GTs (half-warps) per pass; CTA j of the image takes passes j, j + match_ctas, ...
            for (int base = match_j * (kMatchThreads / 16); base < m_img; base += plan.match_ctas * (kMatchThreads / 16)) {
                const int m_raw = base + (threadIdx.x >> 4);
                if (base + ((threadIdx.x & ~31) >> 4) < m_img)                       // warp-uniform
                    match_gt<T>(preds, image, g_begin, m_img, min(m_raw, m_img - 1), m_raw < m_img, n_ch, n_anchors, nc, anchors,
                                strides, gt, best, k_dfl_num, k_cls, WRITE_GRAD ? grad : nullptr, plan.whole_sectors != 0, acc, flags,
                                out_idx, out_iou);
            }
        }
    } else {
        int tile;
        const int groups_c = coarse * n_images;
        if (group < groups_c) {
            image = group / coarse;
            tile = n_tiles - coarse + (group - image * coarse);
        } else {
            const int fine = n_tiles - coarse, gf = group - groups_c;
            image = gf / fine;
            tile = gf - image * fine;
        }
        YB_TRACE_ROLE(role == 0 ? (tile >= n_tiles - coarse ? 1 : 0) : 2);
        if (role == 0)
            assign_body<T, VW>(image, tile, preds, n_ch, n_anchors, anchors, strides, gt, gt_off, best,
                               plan.probe_ctas > 0 ? bound : nullptr, chained ? nullptr : gt_img, WRITE_GRAD ? grad : nullptr, prune);
        else
            cls_body<T, VW, WRITE_GRAD>(image, tile, n_tiles, role - 1, YB_CLS_CSPLIT, preds, n_ch, n_anchors, nc, k_cls, grad,
                                        acc, flags);
    }
    if (!chained) return;
    __syncthreads();
    if (match_image >= 0 && plan.self_clean) {
        // (self-cleaning) the image's last match CTA to finish wipes the image's keys and bounds: every other CTA that reads
        // or writes them has counted itself off (the probe CTAs on their own counter); the reducer no longer looks at them
        __shared__ bool s_last_of_image;
        if (threadIdx.x == 0) {
            // (before this CTA counts itself off: afterwards the reducer may wipe the probes' counter at any time)
            if (plan.probe_ctas > 0) dep_wait(flags + 4, (unsigned int)plan.probe_ctas, flags + 3, kLossPollNs);
            __threadfence();
            s_last_of_image = atomicAdd(done + image, 1u) + 1u == tile_ctas + (unsigned int)plan.match_ctas;
        }
        __syncthreads();
        if (s_last_of_image) {
            const int g_begin = __ldg(gt_off + image), m_img = __ldg(gt_off + image + 1) - g_begin;
            for (int m = threadIdx.x; m < m_img; m += kAssignThreads) {
                best[g_begin + m] = 0ull;
                if (plan.probe_ctas > 0) bound[g_begin + m] = 0u;
            }
        }
        return;
    }
    if (threadIdx.x == 0) dep_signal(done + image);
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

// Optional per-kernel timing (bench.py's roofline leg): the caller may pass three of ITS OWN cudaEvent_t handles,
// recorded on the launching stream before the main launch, between the two launches and after the second.
static int stage_mark(void *const *events, int i, cudaStream_t st) {
    if (events == nullptr || events[i] == nullptr) return YB_OK;
    YB_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(events[i]), st));
    return YB_OK;
}

template <typename T, int VW>
static int launch_loss(const T *preds, int n_images, int nc, int n_anchors, const float *anchors, const float *strides,
                       const float *gt, const int32_t *gt_off, int gt_total, int gmax, float lambda_cls,
                       float lambda_dfl, T *grad, float *out_loss, int32_t *out_idx, float *out_iou,
                       float *out_per_image, const LossWorkspace &w, unsigned flags, const yb_anchor_grid &grid,
                       void *const *stage_events, cudaStream_t st) {
    const int n_ch = 4 * kRegMax + nc;
    const float k_cls = lambda_cls / ((float)n_images * (float)n_anchors);
    const float k_dfl_num = lambda_dfl / ((float)n_images * 4.f);
    constexpr int TILE_A = kAssignThreads * VW, TILE_C = kClsThreads * VW;
    static_assert(TILE_A == TILE_C, "both roles tile the anchors identically");
    const int cls_split = YB_CLS_CSPLIT;
    const int n_tiles = (n_anchors + TILE_C - 1) / TILE_C;
    // whole-sector patches of the gradient rows need 32-byte aligned rows (match_gt)
    const int whole_sectors = YB_MATCH_WHOLE_SECTORS && grad != nullptr && n_anchors % (32 / (int)sizeof(T)) == 0 &&
                              (reinterpret_cast<uintptr_t>(grad) & 31u) == 0;
    // YB_LOSS_WS_CLEAN: the caller vouches that the workspace is all zero -- as this entry point leaves it when it is called
    // with the flag (fused launch only): no memset node in front of the step
    const bool ws_clean = (flags & YB_LOSS_WS_CLEAN) && !(flags & YB_LOSS_SPLIT_LAUNCH) && !YB_MATCH_SEPARATE;
    if (!ws_clean) YB_CUDA(cudaMemsetAsync(w.ticket, 0, w.zero_bytes, st));
    if (int rc = stage_mark(stage_events, 0, st)) return rc;
    if (!(flags & YB_LOSS_SPLIT_LAUNCH)) {
        // Block order: first the last quarter of the tiles of every image -- the coarse pyramid levels (P4 + P5
        // hold 23.8 % of a three-level grid's anchors), whose anchors lie within a cell or two of every GT --
        // then the rest, where assign_body prunes every (GT, tile) pair that cannot beat the distance already
        // published in `best`.  The result never depends on what has been published (see assign_body).
        FusedPlan plan;
        plan.n_tiles = n_tiles;
        plan.coarse = 0;
#if YB_COARSE_FIRST && YB_ASSIGN_PRUNE
        if (gt_total > 0 && n_tiles >= 4) plan.coarse = (n_tiles * YB_COARSE_FRAC256 + 255) / 256;   // ceil(0.238 n_tiles)
#endif
        if (flags & YB_LOSS_NO_PRUNE) plan.coarse = -1;      // pruning (and the coarse-first order) off: the exactness tests compare both
        plan.skew = (int)std::min<long long>(YB_BOX_SKEW, (long long)n_tiles * n_images);
        // match CTAs per image: enough for the largest image in one pass when the caller's gmax is right (8 GTs per
        // CTA); an image with more GTs than that takes further passes, so the result never depends on gmax
        const int g_hint = gmax > 0 ? gmax : (gt_total + n_images - 1) / n_images;
        plan.match_ctas = gt_total > 0 ? std::min(std::max((g_hint + 7) / 8, 1), YB_MATCH_MAX_CTAS) : 0;
        if (YB_MATCH_SEPARATE) plan.match_ctas = -1;
        plan.whole_sectors = whole_sectors;
        plan.gt_total = gt_total;
        plan.self_clean = ws_clean ? 1 : 0;
        // the probe role pays once the coarse tiles would otherwise scan more than one chunk of GTs per image (measured:
        // cfg5, <= 300 GT per image, 188 -> 170 us; cfg2, <= 100 GT per image, 241 -> 252 us with it)
        plan.probe_ctas = (YB_LOSS_PROBE && grid.n_levels > 0 && gt_total > 0 && !(flags & YB_LOSS_NO_PRUNE) &&
                           (g_hint > YB_PROBE_MIN_GT || (flags & YB_LOSS_FORCE_PROBE)))
                              ? (gt_total + kAssignThreads / 16 - 1) / (kAssignThreads / 16) : 0;
        const long long blocks = (long long)n_tiles * (1 + cls_split) * n_images + plan.skew + plan.probe_ctas +
                                 (plan.match_ctas < 0 ? 0 : (long long)plan.match_ctas * n_images + 1);
        YB_REQUIRE(blocks < (1ll << 31), "yb_loss_fwd_bwd: too many tiles for one launch");
        // a programmatic dependent launch: the grid may become resident while its predecessor in the stream drains (it waits
        // at its top for that kernel's completion); YB_LOSS_NO_PDL: a plain launch
        auto launch = [&](auto kernel) -> cudaError_t {
            if (YB_LOSS_PDL && !(flags & YB_LOSS_NO_PDL))
                return launch_pdl(kernel, dim3((unsigned)blocks), dim3(kAssignThreads), 0, st, preds, n_images, n_ch, n_anchors, nc, plan,
                                  grid, anchors, strides, gt, gt_off, w.best, w.bound, w.gt_img, k_cls, k_dfl_num, lambda_cls, lambda_dfl,
                                  grad, w.acc, w.ticket, w.done, out_idx, out_iou, out_loss, out_per_image);
            kernel<<<(unsigned)blocks, kAssignThreads, 0, st>>>(preds, n_images, n_ch, n_anchors, nc, plan, grid, anchors, strides, gt,
                                                                 gt_off, w.best, w.bound, w.gt_img, k_cls, k_dfl_num, lambda_cls,
                                                                 lambda_dfl, grad, w.acc, w.ticket, w.done, out_idx, out_iou, out_loss,
                                                                 out_per_image);
            return cudaSuccess;
        };
        if (grad != nullptr) YB_CUDA(launch(fused_main_kernel<T, VW, true>));
        else YB_CUDA(launch(fused_main_kernel<T, VW, false>));
        YB_LAUNCH_CHECK();
        if (int rc = stage_mark(stage_events, 1, st)) return rc;
        if (plan.match_ctas < 0) {
            const int per_cta = kMatchThreads / 16;
            const int mblocks = std::max(1, (gt_total + per_cta - 1) / per_cta);
            match_kernel<T><<<mblocks, kMatchThreads, 0, st>>>(
                preds, n_images, n_ch, n_anchors, nc, anchors, strides, gt, gt_off, w.gt_img, gt_total, w.best, k_dfl_num, k_cls,
                lambda_cls, lambda_dfl, grad, whole_sectors, w.acc, w.ticket, out_idx, out_iou, out_loss, out_per_image);
            YB_LAUNCH_CHECK();
        }
    } else {
        // two launches (profiling the roles separately).  Stream-level overlap of the two was measured and does
        // not help: the first kernel's CTAs fill every SM, so the second only starts as the first drains.
        assign_kernel<T, VW><<<dim3(n_tiles, n_images), kAssignThreads, 0, st>>>(preds, n_ch, n_anchors, anchors, strides, gt,
                                                                                 gt_off, w.best, w.gt_img, grad);
        YB_LAUNCH_CHECK();
        dim3 grid(n_tiles, n_images, cls_split);
        if (grad != nullptr)
            cls_loss_kernel<T, VW, true><<<grid, kClsThreads, 0, st>>>(preds, n_ch, n_anchors, nc, k_cls, grad, w.acc, w.ticket);
        else
            cls_loss_kernel<T, VW, false><<<grid, kClsThreads, 0, st>>>(preds, n_ch, n_anchors, nc, k_cls, grad, w.acc, w.ticket);
        YB_LAUNCH_CHECK();
        if (int rc = stage_mark(stage_events, 1, st)) return rc;
        const int per_cta = kMatchThreads / 16;                                  // one half-warp per GT
        const int blocks = std::max(1, (gt_total + per_cta - 1) / per_cta);       // >= 1: its last CTA writes the loss
        match_kernel<T><<<blocks, kMatchThreads, 0, st>>>(
            preds, n_images, n_ch, n_anchors, nc, anchors, strides, gt, gt_off, w.gt_img, gt_total, w.best, k_dfl_num, k_cls,
            lambda_cls, lambda_dfl, grad, whole_sectors, w.acc, w.ticket, out_idx, out_iou, out_loss, out_per_image);
        YB_LAUNCH_CHECK();
    }
    if (int rc = stage_mark(stage_events, 2, st)) return rc;
    return YB_OK;
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_loss_workspace_bytes(int n_images, int n_anchors, int gt_total, int dtype) {
    if (n_images <= 0 || n_anchors <= 0 || gt_total < 0) return 0;
    (void)n_anchors; (void)dtype;
    return carve(nullptr, n_images, gt_total).total_bytes;
}

extern "C" int yb_loss_fwd_bwd(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                               const float *anchors, const float *strides, const float *gt,
                               const int32_t *gt_offsets, int gt_total, int gmax, float lambda_cls, float lambda_dfl,
                               void *grad_preds, float *out_loss, int32_t *out_idx, float *out_iou,
                               float *out_per_image, void *workspace, size_t workspace_bytes, unsigned flags,
                               const yb_anchor_grid *grid_hint, void *const *stage_events, void *stream) {
    YB_NVTX("yb_loss_fwd_bwd");
    YB_REQUIRE(preds && anchors && strides && gt_offsets && out_loss && workspace, "yb_loss_fwd_bwd: null pointer");
    YB_REQUIRE(gt_total == 0 || gt != nullptr, "yb_loss_fwd_bwd: gt is null but gt_total > 0");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0 && gt_total >= 0 && gmax >= 0, "yb_loss_fwd_bwd: bad sizes");
    YB_REQUIRE(reg_max == kRegMax, "yb_loss_fwd_bwd: reg_max must be %d (got %d)", kRegMax, reg_max);
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_loss_fwd_bwd: dtype must be YB_F32 or YB_BF16");
    YB_REQUIRE(n_images <= 65535, "yb_loss_fwd_bwd: at most 65535 images per call");
    // the per-image class sums are kept in 2^-32 fixed point: |sum| <= 27.7 * A * nc must stay below 2^31
    YB_REQUIRE((double)n_anchors * (double)nc * 27.7 < 2147483648.0, "yb_loss_fwd_bwd: A * nc too large for the fixed-point sums");
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
    yb_anchor_grid grid;
    memset(&grid, 0, sizeof(grid));                        // n_levels = 0: no hint, no probe role
    if (grid_hint != nullptr && grid_hint->n_levels > 0 && grid_hint->n_levels <= YB_TAL_MAX_LEVELS) {
        bool sane = true;                                  // shapes only; the VALUES cannot hurt (probe_gts clamps)
        for (int l = 0; l < grid_hint->n_levels; ++l)
            sane = sane && grid_hint->w[l] > 0 && grid_hint->h[l] > 0 && grid_hint->stride[l] > 0.f;
        if (sane) grid = *grid_hint;
    }
    if (dtype == YB_F32) {
        const bool vec = vector_ok<float>(preds, grad_preds, n_anchors);
        const LossWorkspace w = carve(workspace, n_images, gt_total);
        if (vec)
            return launch_loss<float, 4>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                         gt_offsets, gt_total, gmax, lambda_cls, lambda_dfl, (float *)grad_preds,
                                         out_loss, out_idx, out_iou, out_per_image, w, flags, grid, stage_events, st);
        return launch_loss<float, 1>((const float *)preds, n_images, nc, n_anchors, anchors, strides, gt, gt_offsets,
                                     gt_total, gmax, lambda_cls, lambda_dfl, (float *)grad_preds, out_loss, out_idx,
                                     out_iou, out_per_image, w, flags, grid, stage_events, st);
    }
    const bool vec = vector_ok<__nv_bfloat16>(preds, grad_preds, n_anchors);
    const LossWorkspace w = carve(workspace, n_images, gt_total);
    if (vec)
        return launch_loss<__nv_bfloat16, 8>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides,
                                             gt, gt_offsets, gt_total, gmax, lambda_cls, lambda_dfl,
                                             (__nv_bfloat16 *)grad_preds, out_loss, out_idx, out_iou, out_per_image, w,
                                             flags, grid, stage_events, st);
    return launch_loss<__nv_bfloat16, 1>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides, gt,
                                         gt_offsets, gt_total, gmax, lambda_cls, lambda_dfl,
                                         (__nv_bfloat16 *)grad_preds, out_loss, out_idx, out_iou, out_per_image, w, flags, grid, stage_events, st);
}

extern "C" int yb_scale_grad(void *grad, int dtype, size_t n_elements, const float *scale, void *stream) {
    YB_NVTX("yb_scale_grad");
    YB_REQUIRE(grad && scale, "yb_scale_grad: null pointer");
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_scale_grad: bad dtype");
    if (n_elements == 0) return YB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = 148 * 8;
    if (dtype == YB_F32)
        scale_kernel<float><<<blocks, 256, 0, st>>>((float *)grad, n_elements, scale);
    else
        scale_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16 *)grad, n_elements, scale);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

extern "C" int yb_loss_fwd_bwd_host(const void *preds_host, int dtype, int n_images, int nc, int reg_max,
                                    int n_anchors, const float *anchors, const float *strides, const float *gt_host,
                                    const int32_t *gt_offsets_host, int gt_total, int gmax, float lambda_cls,
                                    float lambda_dfl, void *preds_dev, float *gt_dev, int32_t *gt_offsets_dev,
                                    void *grad_dev, float *out_loss_dev, float *out_loss_host, void *grad_host,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_loss_fwd_bwd_host");
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
                                   nullptr, nullptr, nullptr, workspace, workspace_bytes, 0u, nullptr, nullptr, stream);
    if (rc != YB_OK) return rc;
    YB_CUDA(cudaMemcpyAsync(out_loss_host, out_loss_dev, sizeof(float) * 8, cudaMemcpyDeviceToHost, st));
    if (grad_host != nullptr && grad_dev != nullptr)
        YB_CUDA(cudaMemcpyAsync(grad_host, grad_dev, n_el * esz, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    return YB_OK;
}
