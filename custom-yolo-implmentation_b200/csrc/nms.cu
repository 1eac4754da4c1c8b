// Batched class-aware NMS (sm_100a).  Replaces non_max_suppression (src/utils/model_utils.py:174-279)
// and the per-image torchvision.ops.nms call inside it, for all images of a batch in three launches
// (scan, class-parallel kernel, generic kernel for whatever the class-parallel one declined):
//
//   nms_scan_kernel    reads (N, 4+nc, A) once with 128-bit loads: best class (first max), strict
//                      `> conf`, optional class filter; candidates are compacted with one atomic per
//                      warp into 64-bit (score, anchor) keys.
//   nms_class_kernel   see "Class-parallel path" below.
//   nms_sweep_kernel   one CTA per image the class-parallel kernel left: bitonic sort of the keys (score
//                      descending, ties -> lowest anchor), then greedy suppression in score order.  Each thread keeps its
//                      columns' class-offset boxes in registers; the sorted list is walked 32 boxes at
//                      a time: the owning warp resolves the 32x32 block with shuffles + ballots, the
//                      kept rows are broadcast through shared memory and every thread clears the alive
//                      bits of its later columns.  Only KEPT rows are ever compared against the rest,
//                      and the walk stops at max_det keeps, so the n x n mask is never built.
//
// Exactness: IoU = inter / (area_a + area_b - inter) in fp32 with IEEE division and no epsilon, on
// boxes offset by cls * 7680 in fp32 (model_utils.py:262-263), compared as torchvision's CPU kernel
// compares it (float IoU promoted to double against the double threshold).
#include "sort.cuh"

namespace yb {

constexpr int kScanThreads = 128;
constexpr float kMaxWh = 7680.f;      // model_utils.py:210
constexpr int kMaxNms = 30000;        // model_utils.py:211
constexpr int kRegCols = 9;           // columns per thread held in registers -> up to 9216 candidates

struct NmsWorkspace {
    int *count;                   // [N]  (zeroed every call)
    int *cls;                     // [N * A]
    unsigned long long *keys;     // [N * Apad]
    float4 *sbox;                 // [N * A] sorted offset boxes (only used beyond kRegCols * 1024 candidates)
    unsigned long long *keys2;    // [N * Apad] class-segmented sorted keys of the class-parallel path
    int *mode;                    // [N] 1 = image finished by the class-parallel kernel (zeroed every call)
    unsigned *tick;               // [N] CTAs of the image that finished their classes (zeroed every call)
    unsigned *range;              // [N][2] encoded min x1 / max x2 of the image's candidates (zeroed every call)
    unsigned char *alive_g;       // [N * Apad] survivor flag per class-sorted position
    int a_pad;
    size_t zero_bytes, total_bytes;
};

static NmsWorkspace carve_nms(void *base, int n_images, int n_anchors) {
    NmsWorkspace w;
    char *p = static_cast<char *>(base);
    size_t off = 0;
    w.count = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images, 64);
    w.mode = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images, 64);
    w.tick = reinterpret_cast<unsigned *>(p + off);
    off += round_up(sizeof(unsigned) * (size_t)n_images, 64);
    w.range = reinterpret_cast<unsigned *>(p + off);
    off += round_up(sizeof(unsigned) * 2 * (size_t)n_images, 64);
    w.zero_bytes = off;
    w.cls = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images * n_anchors, 64);
    w.a_pad = next_pow2(n_anchors);
    w.keys = reinterpret_cast<unsigned long long *>(p + off);
    off += sizeof(unsigned long long) * (size_t)n_images * w.a_pad;
    w.keys2 = reinterpret_cast<unsigned long long *>(p + off);
    off += sizeof(unsigned long long) * (size_t)n_images * w.a_pad;
    w.alive_g = reinterpret_cast<unsigned char *>(p + off);
    off += round_up((size_t)n_images * w.a_pad, 64);
    w.sbox = reinterpret_cast<float4 *>(p + off);
    off += (n_anchors > kRegCols * kSortThreads) ? sizeof(float4) * (size_t)n_images * n_anchors : 0;
    w.total_bytes = off;
    return w;
}

// Where a launch finds its inputs.  The reference hands NMS one (N, 4 + nc, A) fp32 tensor [xywh | scores]
// (src/utils/model_utils.py:174); the fused post-processing entry point (yb_postprocess) reads the scores straight from
// the head output (N, 64 + nc, A), fp32 or bf16, optionally through a sigmoid, and the boxes from a compact (N, 4, A)
// buffer the decode kernel wrote.
struct NmsInput {
    const float *box;            // image b: box + b * box_stride, rows x, y, w, h of n_anchors floats each
    size_t box_stride;
    const void *score;           // image b: score + b * score_stride (elements), nc rows of n_anchors elements
    size_t score_stride;
};

__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }   // as Tensor.sigmoid()

template <typename T, int VW, bool SIGMOID>
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const T *__restrict__ score, size_t score_stride, int nc, int n_anchors, float conf,
                const int *__restrict__ filter, int n_filter, int *__restrict__ count, int *__restrict__ cls_out,
                unsigned long long *__restrict__ keys, int a_pad) {
    pdl_launch_dependents();                               // the class-parallel kernel's CTAs move in behind this grid's last wave
    const int n = blockIdx.y;
    const int a0 = (blockIdx.x * kScanThreads + threadIdx.x) * VW;
    const int lane = threadIdx.x & 31;
    float best[VW];
    int arg[VW];
    bool pass[VW];
    int n_pass = 0;
#pragma unroll
    for (int v = 0; v < VW; ++v) pass[v] = false;
    if (a0 < n_anchors) {
        const T *pred = score;
        const size_t base = (size_t)n * score_stride + a0;
        auto val = [](float x) { return SIGMOID ? sigmoid_rn(x) : x; };
        {
            Group<T, VW> row;
            row.load(pred + base);
#pragma unroll
            for (int v = 0; v < VW; ++v) { best[v] = val(row.get(v)); arg[v] = 0; }
        }
        constexpr int U = 4;
        int c = 1;
        for (; c + U <= nc; c += U) {
            Group<T, VW> row[U];
#pragma unroll
            for (int u = 0; u < U; ++u) row[u].load(pred + base + (size_t)(c + u) * n_anchors);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    const float s = val(row[u].get(v));
                    if (s > best[v]) { best[v] = s; arg[v] = c + u; }      // first maximum wins
                }
        }
        for (; c < nc; ++c) {
            Group<T, VW> row;
            row.load(pred + base + (size_t)c * n_anchors);
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                const float s = val(row.get(v));
                if (s > best[v]) { best[v] = s; arg[v] = c; }
            }
        }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            bool ok = best[v] > conf;                                       // strict (model_utils.py:206, :245)
            if (ok && n_filter > 0) {
                bool in = false;
                for (int f = 0; f < n_filter; ++f) in |= (__ldg(filter + f) == arg[v]);
                ok = in;
            }
            pass[v] = ok;
            n_pass += ok ? 1 : 0;
            cls_out[(size_t)n * n_anchors + a0 + v] = arg[v];
        }
    }
    int incl = n_pass;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int slot = 0;
    if (lane == 31 && total > 0) slot = atomicAdd(count + n, total);
    slot = __shfl_sync(0xffffffffu, slot, 31) + incl - n_pass;
#pragma unroll
    for (int v = 0; v < VW; ++v)
        if (pass[v]) keys[(size_t)n * a_pad + slot++] = make_score_key(best[v], (unsigned int)(a0 + v));
}

// torchvision's IoU test on two xyxy boxes, bit-exact:  fl(inter / (area_a + area_b - inter)) > thr
// with thr the largest float <= the double threshold.  Boxes that do not intersect give IoU 0 (or
// NaN), never above thr >= 0; otherwise a MUFU.RCP estimate decides unless it lands within 1e-6
// (relative) of the threshold, where the IEEE division is taken.
__device__ __forceinline__ bool iou_exceeds(const float4 &a, float area_a, const float4 &b, float thr) {
    const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (!(w > 0.f && h > 0.f)) return false;
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    // IoU <= min(area) / max(area): boxes of clearly different size cannot exceed the threshold (the 1 %
    // margin is far above the rounding of the exact test below; NaN / non-positive areas fall through)
    if (fminf(area_a, area_b) < 0.99f * thr * fmaxf(area_a, area_b) && fminf(area_a, area_b) > 0.f) return false;
    const float inter = __fmul_rn(w, h);
    const float den = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float est = inter * fast_rcp(den);
    if (fabsf(est - thr) > fmaf(1e-6f, thr, 1e-35f)) return est > thr;
    return __fdiv_rn(inter, den) > thr;
}
__device__ __forceinline__ float box_area(const float4 &b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

// sorted position -> class-offset xyxy box (model_utils.py:239, :262-263)
__device__ __forceinline__ float4 load_offset_box(const float *__restrict__ img, int n_anchors, int a, int cls,
                                                  bool agnostic) {
    const float x = __ldg(img + a), y = __ldg(img + n_anchors + a);
    const float dw = __fmul_rn(__ldg(img + 2 * (size_t)n_anchors + a), 0.5f);
    const float dh = __fmul_rn(__ldg(img + 3 * (size_t)n_anchors + a), 0.5f);
    const float off = agnostic ? 0.f : __fmul_rn((float)cls, kMaxWh);
    return make_float4(__fadd_rn(__fsub_rn(x, dw), off), __fadd_rn(__fsub_rn(y, dh), off),
                       __fadd_rn(__fadd_rn(x, dw), off), __fadd_rn(__fadd_rn(y, dh), off));
}

// Cheap necessary condition for iou_exceeds: the boxes overlap in x and in y (four comparisons).
// (Non-short-circuit on purpose: `&&` compiled into a chain of divergent branches in the pair loops, 17 BRA per
// four rows; as predicate logic the test is four FSETP feeding one PLOP3.)
__device__ __forceinline__ bool boxes_touch(const float4 &a, const float4 &b) {
    return (a.z > b.x) & (b.z > a.x) & (a.w > b.y) & (b.w > a.y);
}

// IoU <= min(area) / max(area), so |log2(area_a) - log2(area_b)| > -log2(0.99 * thr) rules a pair out
// with one subtraction and one comparison (MUFU.LG2 is good to ~2^-22, the 1 % margin is 0.0145 in log2
// units).  NaN (non-positive or infinite areas on both sides) never rules out; the exact test decides.
__device__ __forceinline__ float log_area(const float4 &b) { return __log2f(box_area(b)); }
__device__ __forceinline__ float log_area_bound(float thr) { return -__log2f(0.99f * thr); }
__device__ __forceinline__ bool areas_compatible(float la, float lb, float bound) { return !(fabsf(la - lb) > bound); }

// Resolve one block of 32 consecutive sorted boxes inside a warp.  `bx`/`alive` belong to lane r =
// position c0 + r.  Returns the mask of rows kept (already capped to `room`).
__device__ __forceinline__ unsigned resolve_block(const float4 &bx, bool alive, float thr, int room) {
    const int lane = threadIdx.x & 31;
    unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
    const float my_area = box_area(bx);
    const float my_la = __log2f(my_area), la_bound = log_area_bound(thr);
    // Every unordered pair is tested once: in step k lane l meets lane (l + k) mod 32.  The test is
    // symmetric, so the lane records either "I suppress the later box" (sup) or "the earlier box
    // suppresses me" (by), whichever side of the pair it is on.
    unsigned sup = 0, by = 0;
#pragma unroll 4
    for (int k = 1; k <= 16; ++k) {
        const int partner = (lane + k) & 31;
        float4 o;
        o.x = __shfl_sync(0xffffffffu, bx.x, partner);
        o.y = __shfl_sync(0xffffffffu, bx.y, partner);
        o.z = __shfl_sync(0xffffffffu, bx.z, partner);
        o.w = __shfl_sync(0xffffffffu, bx.w, partner);
        const float o_la = __shfl_sync(0xffffffffu, my_la, partner);
        const bool maybe = boxes_touch(bx, o) & areas_compatible(my_la, o_la, la_bound);
        if (__any_sync(0xffffffffu, maybe)) {
            if (maybe && iou_exceeds(bx, my_area, o, thr)) {
                if (partner > lane) sup |= 1u << partner;
                else by |= 1u << partner;
            }
        }
    }
    unsigned kept = 0;
    if (!__any_sync(0xffffffffu, (sup | by) != 0u)) {
        kept = alive_mask;                                 // no pair of the block exceeds the threshold (the usual case)
    } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const unsigned s_r = __shfl_sync(0xffffffffu, sup, r) | __ballot_sync(0xffffffffu, (by >> r) & 1u);
            if ((alive_mask >> r) & 1u) {
                kept |= 1u << r;
                alive_mask &= ~s_r;
            }
        }
    }
    while (__popc(kept) > room) kept &= ~(0x80000000u >> __clz(kept));   // drop the lowest-ranked keeps
    return kept;
}

__device__ __forceinline__ int cls_n_of(const int *cls, int n, int n_anchors, int a) {
    return cls[(size_t)n * n_anchors + a];
}

__device__ __forceinline__ void emit_row(float *__restrict__ out_rows, int *__restrict__ out_anchor, int n, int max_det,
                                         int pos, const float *__restrict__ img, int n_anchors, int a, float score,
                                         int cls) {
    const float x = __ldg(img + a), y = __ldg(img + n_anchors + a);
    const float dw = __fmul_rn(__ldg(img + 2 * (size_t)n_anchors + a), 0.5f);
    const float dh = __fmul_rn(__ldg(img + 3 * (size_t)n_anchors + a), 0.5f);
    float *row = out_rows + ((size_t)n * max_det + pos) * 6;
    row[0] = __fsub_rn(x, dw);
    row[1] = __fsub_rn(y, dh);
    row[2] = __fadd_rn(x, dw);
    row[3] = __fadd_rn(y, dh);
    row[4] = score;
    row[5] = (float)cls;
    if (out_anchor) out_anchor[(size_t)n * max_det + pos] = a;
}

// REG = true: up to kRegCols*1024 candidates, boxes and alive bits in registers.
// REG = false: any count up to kMaxNms, boxes in a global scratch, alive bits in shared memory.
// MULTI (multi_label): a key's low word is anchor * nc + class instead of the anchor.
template <bool REG, bool MULTI>
__global__ void __launch_bounds__(kSortThreads, 1)
nms_sweep_kernel(const float *__restrict__ pred, size_t box_stride, int nc, int n_anchors, const int *count,
                 const int *mode, const int *cls, unsigned long long *keys, int a_pad,   // written by the predecessor kernels: no
                 // `const __restrict__` (a read-only load may be scheduled above pdl_wait())
                 float4 *__restrict__ sbox, int sbox_stride, float thr, int max_det, int agnostic, float *__restrict__ out_rows,
                 int *__restrict__ out_count, int *__restrict__ out_anchor) {
    __shared__ float4 s_row[32];
    __shared__ float s_area[32];
    __shared__ int s_nrow;
    __shared__ unsigned s_alive[REG ? 1 : (kMaxNms + 31) / 32 + 1];

    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();                                            // launched under programmatic dependent launch behind the kernel that wrote `mode`
    if (!MULTI && mode[n] != 0) return;                    // finished by the class-parallel kernel
    const int n_cand = min(min(count[n], a_pad), kMaxNms);
    auto anchor_of = [&](unsigned id) { return MULTI ? (int)(id / (unsigned)nc) : (int)id; };
    auto class_of = [&](unsigned id, int a) { return MULTI ? (int)(id % (unsigned)nc) : cls_n_of(cls, n, n_anchors, a); };
    if (n_cand == 0 || max_det <= 0) {
        if (tid == 0) out_count[n] = 0;
        return;
    }
    if (REG && n_cand > kRegCols * kSortThreads) return;      // host picks the other variant; never taken
    const float *img = pred + (size_t)n * box_stride;
    const int *cls_n = cls + (size_t)n * n_anchors;
    {   // bitonic sort of all the image's keys (ascending key = descending score, ties -> lowest anchor)
        extern __shared__ unsigned long long s_keys[];
        unsigned long long *ks = keys + (size_t)n * a_pad;
        const int cnt = min(count[n], a_pad);
        if (cnt > 1) {
            const int n_pad = next_pow2(cnt);
            for (int t = cnt + threadIdx.x; t < n_pad; t += blockDim.x) ks[t] = kSentinel;
            __syncthreads();
            cta_bitonic_sort(ks, n_pad, s_keys);
        }
        __syncthreads();
    }
    const unsigned long long *k = keys + (size_t)n * a_pad;

    float4 box[REG ? kRegCols : 1];
    unsigned alive = 0;
    if (REG) {
#pragma unroll
        for (int t = 0; t < kRegCols; ++t) {
            const int j = tid + t * kSortThreads;
            box[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n_cand) {
                const unsigned id = key_anchor(k[j]);
                const int a = anchor_of(id);
                box[t] = load_offset_box(img, n_anchors, a, class_of(id, a), agnostic != 0);
                alive |= 1u << t;
            }
        }
    } else {
        float4 *sb = sbox + (size_t)n * sbox_stride;
        for (int j = tid; j < n_cand; j += kSortThreads) {
            const unsigned id = key_anchor(k[j]);
            const int a = anchor_of(id);
            sb[j] = load_offset_box(img, n_anchors, a, class_of(id, a), agnostic != 0);
        }
        for (int w = tid; w < (n_cand + 31) / 32; w += kSortThreads)
            s_alive[w] = (w * 32 + 32 <= n_cand) ? 0xffffffffu : ((1u << (n_cand - w * 32)) - 1u);
    }
    __syncthreads();
    const float4 *sb = REG ? nullptr : sbox + (size_t)n * sbox_stride;

    int kept_total = 0;
    for (int c0 = 0; c0 < n_cand; c0 += 32) {
        const int own_warp = REG ? ((c0 >> 5) & 31) : 0;
        const int own_slot = c0 >> 10;
        if (warp == own_warp) {
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            bool al = false;
            if (REG) {
#pragma unroll
                for (int t = 0; t < kRegCols; ++t)
                    if (t == own_slot) { bx = box[t]; al = (alive >> t) & 1u; }
            } else if (c0 + lane < n_cand) {
                bx = sb[c0 + lane];
                al = (s_alive[c0 >> 5] >> lane) & 1u;
            }
            const unsigned kept = resolve_block(bx, al, thr, max_det - kept_total);
            if ((kept >> lane) & 1u) {
                const int rank = __popc(kept & ((1u << lane) - 1u));
                s_row[rank] = bx;
                s_area[rank] = box_area(bx);
                const unsigned long long key = k[c0 + lane];
                const int a = anchor_of(key_anchor(key));
                emit_row(out_rows, out_anchor, n, max_det, kept_total + rank, img, n_anchors, a, key_score(key),
                         class_of(key_anchor(key), a));
            }
            if (lane == 0) s_nrow = __popc(kept);
        }
        __syncthreads();
        const int n_row = s_nrow;
        kept_total += n_row;
        if (kept_total >= max_det) break;
        if (n_row > 0) {
            if (REG) {
#pragma unroll
                for (int t = 0; t < kRegCols; ++t) {
                    const int j = tid + t * kSortThreads;
                    if (j >= c0 + 32 && ((alive >> t) & 1u)) {
                        for (int r = 0; r < n_row; ++r)
                            if (iou_exceeds(s_row[r], s_area[r], box[t], thr)) { alive &= ~(1u << t); break; }
                    }
                }
            } else {
                for (int j = c0 + 32 + tid; j < n_cand; j += kSortThreads) {
                    if (!((s_alive[j >> 5] >> (j & 31)) & 1u)) continue;
                    const float4 b = sb[j];
                    for (int r = 0; r < n_row; ++r)
                        if (iou_exceeds(s_row[r], s_area[r], b, thr)) { atomicAnd(&s_alive[j >> 5], ~(1u << (j & 31))); break; }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) out_count[n] = min(kept_total, max_det);
}

// ------------------------------------------------------------------------------------------
// Class-parallel path.  With the reference's class offset (cls * 7680) boxes of different classes
// cannot intersect as long as all candidate boxes of the image span less than 7680 px in x, so greedy
// NMS decomposes into one independent problem per class.  One CTA per image:
//   1. class histogram + exclusive scan (shared-memory atomics)  -> one segment per class
//   2. scatter the (score, anchor) keys into their segments, rank-sort every segment by one warp
//   3. gather the class-offset boxes into shared memory (score order inside each segment)
//   4. greedy NMS per class: one warp per class walks its rows, lanes test the later columns and
//      clear their alive bits with warp ballots
//   5. 5-pass radix select (11-bit digits) of the max_det best survivors, bitonic sort, emit
// Images the path cannot take (too many candidates for shared memory, one class larger than
// kClassSegCap, boxes spanning more than 7680 px) are left to the generic sort + sweep kernels,
// which run afterwards and skip every image whose mode flag is set.  Both paths give the same rows.
// ------------------------------------------------------------------------------------------
#ifdef YB_NMS_PROFILE
__device__ long long g_nms_clk[16];
__device__ long long g_nms_warp[32][4];
#define YB_MARK(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_nms_clk[i] = clock64(); } while (0)
#else
#define YB_MARK(i)
#endif
constexpr int kClassThreads = 1024;
constexpr int kClassCap = 10240;       // candidates per image held in shared memory (16 B each)
constexpr int kClassSegCap = 512;      // largest class segment the warp-per-class sweep accepts
constexpr int kClassMaxNc = 1023;      // 10 class bits in the key
constexpr int kClassMaxDet = 1024;
constexpr int kAnchorBits = 22;
constexpr int kSelBits = 11;


__global__ void __launch_bounds__(kClassThreads, 1)
nms_class_kernel(const float *__restrict__ pred, size_t box_stride, int nc, int n_anchors, const int *count,
                 const int *cls, const unsigned long long *keys, int a_pad,   // nms_scan_kernel's output: no `const __restrict__`
                 // (a read-only load may be scheduled above pdl_wait())
                 unsigned long long *__restrict__ keys2, unsigned char *__restrict__ alive_g,
                 unsigned *__restrict__ tick, unsigned *__restrict__ range, float thr, int max_det,
                 int *__restrict__ mode, float *__restrict__ out_rows, int *__restrict__ out_count,
                 int *__restrict__ out_anchor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_start[kClassMaxNc + 2];
    __shared__ int s_cursor[kClassMaxNc + 1];
    __shared__ unsigned s_alive[kClassCap / 4];           // used as one byte per sorted position
    __shared__ int s_flag, s_total;
    __shared__ unsigned s_minx, s_maxx;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_k, s_exact;
    __shared__ float4 s_rowbox[kClassThreads / 32][32];   // per warp: kept rows of the current 32-box block
    __shared__ float s_rowla[kClassThreads / 32][32];     //           and log2 of their areas
    __shared__ unsigned short s_order[kClassMaxNc + 1];   // this CTA's classes, largest segment first
    __shared__ int s_next[2];                             // work counters of the two per-class phases

    // gridDim.x CTAs share an image: CTA `part` owns the classes c with c % gridDim.x == part
    const int img = blockIdx.y, part = blockIdx.x, n_part = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();                                            // nms_scan_kernel's counts, classes and keys
    const int n = count[img];
    if (n == 0) {
        if (tid == 0 && part == 0) { out_count[img] = 0; mode[img] = 1; }
        return;
    }
    if (n > kClassCap) return;                             // generic path (same decision in every CTA)
    const float *img_pred = pred + (size_t)img * box_stride;
    const unsigned long long *k_in = keys + (size_t)img * a_pad;
    unsigned long long *k2 = keys2 + (size_t)img * a_pad;
    const int *cls_n = cls + (size_t)img * n_anchors;
    unsigned long long *K = reinterpret_cast<unsigned long long *>(smem_raw);

    YB_MARK(0);
    // ---- 1. histogram of classes, exclusive scan ------------------------------------------------
    for (int c = tid; c < nc; c += kClassThreads) s_cursor[c] = 0;
    if (tid == 0) { s_flag = 0; s_minx = 0u; s_maxx = 0u; s_total = 0; s_next[0] = 0; s_next[1] = 0; }
    __syncthreads();
    for (int r = tid; r < n; r += kClassThreads) atomicAdd(&s_cursor[cls_n[key_anchor(k_in[r])]], 1);
    __syncthreads();
    if (warp == 0) {
        int run = 0, big = 0;
        for (int base = 0; base < nc; base += 32) {
            const int c = base + lane;
            const int v = c < nc ? s_cursor[c] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (c < nc) s_start[c] = run + incl - v;
            big |= v > kClassSegCap;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        big = __any_sync(0xffffffffu, big);
        if (lane == 0) { s_start[nc] = run; s_flag = big; }
    }
    __syncthreads();
    if (s_flag) return;                                    // a class too large for one warp: generic path
    // Warps take classes from a shared counter, largest segment first (the per-class work grows with
    // the square of the segment, so index order leaves the warp with two large classes far behind).
    const int n_my = (nc - part + n_part - 1) / n_part;    // classes of this CTA: part, part + n_part, ...
    for (int t = tid; t < n_my; t += kClassThreads) {
        const int c = part + t * n_part;
        const int sz = s_start[c + 1] - s_start[c];
        int rank = 0;
        for (int u = 0; u < n_my; ++u) {
            const int cu = part + u * n_part;
            const int su = s_start[cu + 1] - s_start[cu];
            rank += (su > sz || (su == sz && u < t)) ? 1 : 0;
        }
        s_order[rank] = (unsigned short)c;
    }
    for (int c = tid; c < nc; c += kClassThreads) s_cursor[c] = s_start[c];
    __syncthreads();

    YB_MARK(1);
    // ---- 2. scatter into class segments, rank-sort each segment (score desc, ties -> lowest anchor)
    for (int r = tid; r < n; r += kClassThreads) {
        const unsigned long long key = k_in[r];
        const unsigned a = key_anchor(key);
        const int c = cls_n[a];
        if (c % n_part != part) continue;
        const int slot = atomicAdd(&s_cursor[c], 1);
        K[slot] = ((key >> 32) << kAnchorBits) | a;        // 32-bit inverted score | 22-bit anchor
    }
    __syncthreads();
    YB_MARK(2);
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&s_next[0], 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_my) break;
        const int c = s_order[item];
        const int seg0 = s_start[c], n_c = s_start[c + 1] - seg0;
        for (int e0 = 0; e0 < n_c; e0 += 128) {            // four elements per lane per sweep of the segment
            unsigned long long ke[4];
            int rank[4] = {0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < 4; ++u) ke[u] = (e0 + u * 32 + lane < n_c) ? K[seg0 + e0 + u * 32 + lane] : kSentinel;
#pragma unroll 4
            for (int i = 0; i < n_c; ++i) {
                const unsigned long long ki = K[seg0 + i];  // broadcast
#pragma unroll
                for (int u = 0; u < 4; ++u) rank[u] += (ki < ke[u]) ? 1 : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e0 + u * 32 + lane < n_c) k2[seg0 + rank[u]] = ((unsigned long long)c << 54) | ke[u];
        }
    }
    __syncthreads();

    YB_MARK(3);
    float4 *B = reinterpret_cast<float4 *>(smem_raw) + 0;  // boxes overwrite the unsorted keys: see the barrier above
    float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
    YB_MARK(4);
    // ---- 4. greedy NMS, one warp per class, 32 boxes at a time ---------------------------------
    // Lane l owns column l of every 32-box block of the segment; bit b of `alive_bits` says whether
    // its column in block b is still alive.  Per row block: resolve the 32x32 diagonal block with
    // shuffles (resolve_block), then let the kept rows clear later columns.  All in registers.
    unsigned char *alive_img = alive_g + (size_t)img * a_pad;                // one byte per sorted position
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&s_next[1], 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_my) break;
        const int c = s_order[item];
        const int seg0 = s_start[c], n_c = s_start[c + 1] - seg0;
        if (n_c == 0) break;                               // largest first: every later class is empty too
        const int nb = (n_c + 31) >> 5;                    // <= 16 (kClassSegCap)
        // class-offset boxes of the segment into shared memory, in score order (model_utils.py:239, :262-263)
        for (int m0 = 0; m0 < n_c; m0 += 128) {
            float bx[4], by[4], bw[4], bh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = m0 + u * 32 + lane;
                const int a = m < n_c ? (int)(k2[seg0 + m] & ((1u << kAnchorBits) - 1u)) : 0;
                bx[u] = __ldg(img_pred + a);
                by[u] = __ldg(img_pred + n_anchors + a);
                bw[u] = __ldg(img_pred + 2 * (size_t)n_anchors + a);
                bh[u] = __ldg(img_pred + 3 * (size_t)n_anchors + a);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = m0 + u * 32 + lane;
                if (m >= n_c) continue;
                const float dw = __fmul_rn(bw[u], 0.5f), dh = __fmul_rn(bh[u], 0.5f);
                const float off = __fmul_rn((float)c, kMaxWh);
                const float x1 = __fsub_rn(bx[u], dw), x2 = __fadd_rn(bx[u], dw);
                B[seg0 + m] = make_float4(__fadd_rn(x1, off), __fadd_rn(__fsub_rn(by[u], dh), off), __fadd_rn(x2, off),
                                          __fadd_rn(__fadd_rn(by[u], dh), off));
                lo = fminf(lo, fminf(x1, x2));
                hi = fmaxf(hi, fmaxf(x1, x2));
            }
        }
        __syncwarp();
#ifdef YB_NMS_PROFILE
        long long t_res = 0, t_cross = 0, t0c = clock64();
#endif
        unsigned alive_bits = 0;
        const float la_bound = log_area_bound(thr);
        // (Measured and not kept: for segments of up to 128 boxes, all rows against all later columns first -- ballot words
        // of a block-upper-triangular suppression matrix in registers, no row waiting for the fate of another -- and the
        // greedy order applied afterwards by a sweep over the words: 187 us per 64 images against 153 us for the
        // row-by-row form below, which skips the rows and columns already suppressed.)
        for (int b = 0; b < nb; ++b) alive_bits |= (b * 32 + lane < n_c ? 1u : 0u) << b;
        for (int rb = 0; rb < nb; ++rb) {
            const int my = rb * 32 + lane;
            const float4 rbox = my < n_c ? B[seg0 + my] : make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef YB_NMS_PROFILE
            long long ta = clock64();
#endif
            const unsigned kept = resolve_block(rbox, (alive_bits >> rb) & 1u, thr, 32);
#ifdef YB_NMS_PROFILE
            long long tb = clock64(); t_res += tb - ta;
#endif
            alive_bits = (alive_bits & ~(1u << rb)) | (((kept >> lane) & 1u) << rb);
            if (rb + 1 == nb || kept == 0u) continue;
            // the kept rows of this block, compacted into the warp's scratch list (box + log2 area)
            const int nk = __popc(kept);
            if ((kept >> lane) & 1u) {
                const int rank = __popc(kept & ((1u << lane) - 1u));
                s_rowbox[warp][rank] = rbox;
                s_rowla[warp][rank] = log_area(rbox);
            }
            __syncwarp();
            for (int cb = rb + 1; cb < nb; ++cb) {
                const int col = cb * 32 + lane;
                bool live = (alive_bits >> cb) & 1u;
                const float4 cbox = live ? B[seg0 + col] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float c_area = box_area(cbox), c_la = __log2f(c_area);
                bool sup = false;
                for (int i = 0; i < nk; i += 4) {          // warp-uniform walk over the kept rows, four at a time
                    bool maybe[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = min(i + u, nk - 1);  // the tail repeats the last row: harmless
                        maybe[u] = live & boxes_touch(s_rowbox[warp][r], cbox) &
                                   areas_compatible(s_rowla[warp][r], c_la, la_bound);
                    }
                    // most same-class pairs cannot reach the threshold: one vote skips the IoU arithmetic
                    if (__any_sync(0xffffffffu, maybe[0] | maybe[1] | maybe[2] | maybe[3])) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (maybe[u]) sup |= iou_exceeds(cbox, c_area, s_rowbox[warp][min(i + u, nk - 1)], thr);
                    }
                }
                live = live && !sup;
                if (!live) alive_bits &= ~(1u << cb);
            }
            __syncwarp();                                  // the list is rewritten by the next row block
        }
        for (int b = 0; b < nb; ++b)
            if (b * 32 + lane < n_c) alive_img[seg0 + b * 32 + lane] = (alive_bits >> b) & 1u;
#ifdef YB_NMS_PROFILE
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) { g_nms_warp[warp][0] += t_res; g_nms_warp[warp][1] += clock64() - t0c; g_nms_warp[warp][2] += n_c; g_nms_warp[warp][3] += 1; }
#endif
    }
    // x extent of this CTA's candidates -> global, encoded so that atomicMax on zeroed words works
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    auto enc = [](float x) { const unsigned u = __float_as_uint(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); };
    auto dec = [](unsigned e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); };
    if (lane == 0 && lo <= hi) {
        atomicMax(&s_minx, ~enc(lo));
        atomicMax(&s_maxx, enc(hi));
    }
    __syncthreads();
    if (tid == 0) {
        if (s_maxx != 0u) {
            atomicMax(range + 2 * img, s_minx);
            atomicMax(range + 2 * img + 1, s_maxx);
        }
        __threadfence();                                   // this CTA's keys2 / alive flags / range are published
        s_flag = (atomicAdd(tick + img, 1u) == (unsigned)n_part - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_flag) return;                                   // the last CTA of the image finishes it
    __threadfence();
    {
        const unsigned e_lo = ~__ldcg(range + 2 * img), e_hi = __ldcg(range + 2 * img + 1);
        // classes stay disjoint only while every box fits in one 7680-px slot (2 px of rounding slack);
        // infinite coordinates fail the test and go to the generic path
        if (!(dec(e_hi) - dec(e_lo) <= kMaxWh - 2.f)) return;
    }

    YB_MARK(5);
    // ---- 5. the max_det best survivors (smallest 54-bit keys) -------------------------------------
    // all keys and survivor flags of the image into shared memory, then an 11-bit radix select
    unsigned char *alive8 = reinterpret_cast<unsigned char *>(s_alive);
    unsigned long long *KS = reinterpret_cast<unsigned long long *>(smem_raw);
    int *hist = reinterpret_cast<int *>(smem_raw + sizeof(unsigned long long) * (size_t)kClassCap);
    unsigned long long *L = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(unsigned long long) * (size_t)kClassCap +
                                                                  sizeof(int) * (1 << kSelBits));
    int my_alive = 0;
    for (int q = tid; q < n; q += kClassThreads) {
        KS[q] = __ldcg(k2 + q);
        const unsigned char al = __ldcg(alive_img + q);
        alive8[q] = al;
        my_alive += al;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_alive += __shfl_xor_sync(0xffffffffu, my_alive, o);
    if (lane == 0 && my_alive) atomicAdd(&s_total, my_alive);
    __syncthreads();
    const int n_surv = s_total;
    const int n_sel = min(n_surv, max_det);
    const unsigned long long key_mask = (1ull << 54) - 1ull;
    unsigned long long limit = key_mask;                   // select every survivor whose key54 <= limit
    if (n_surv > max_det) {
        if (tid == 0) { s_prefix = 0ull; s_k = max_det; s_exact = 0; }
        int shift = 54;
        for (int pass = 0; pass < 5; ++pass) {
            const int bits = pass < 4 ? kSelBits : 54 - 4 * kSelBits;
            const int hi_shift = shift;                    // bits above hi_shift are fixed by s_prefix
            shift -= bits;
            for (int b = tid; b < (1 << kSelBits); b += kClassThreads) hist[b] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (int q = tid; q < n; q += kClassThreads) {
                if (!alive8[q]) continue;
                const unsigned long long k54 = KS[q] & key_mask;
                if (hi_shift < 54 && (k54 >> hi_shift) != prefix) continue;
                // lanes with the same digit share one shared-memory atomic
                const int digit = (int)((k54 >> shift) & ((1u << bits) - 1u));
                const unsigned peers = __match_any_sync(__activemask(), digit);
                if ((__ffs(peers) - 1) == lane) atomicAdd(&hist[digit], __popc(peers));
            }
            __syncthreads();
            if (warp == 0) {                               // find the digit where the running count reaches k
                const int per = (1 << kSelBits) / 32;
                int local = 0;
                for (int b = 0; b < per; ++b) local += hist[lane * per + b];
                int incl = local;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int excl = incl - local;
                const int k = s_k;
                if (excl < k && k <= incl) {               // exactly one lane
                    int run = excl;
                    for (int b = 0; b < per; ++b) {
                        const int h = hist[lane * per + b];
                        if (run + h >= k) {
                            s_prefix = (prefix << bits) | (unsigned long long)(lane * per + b);
                            s_k = k - run;
                            s_exact = (run + h == k) ? 1 : 0;      // the whole bucket is wanted: no need to look inside it
                            break;
                        }
                        run += h;
                    }
                }
            }
            __syncthreads();
            if (s_exact) {                                 // every key with this prefix (or a smaller one) is selected
                limit = (s_prefix << shift) | ((1ull << shift) - 1ull);
                break;
            }
            if (pass == 4) limit = s_prefix;               // the max_det-th smallest key
        }
    }
    YB_MARK(6);
    // collect, sort (ascending key = descending score), emit
    const int n_pad = next_pow2(max(n_sel, 1));
    if (tid == 0) s_total = 0;
    for (int t = tid; t < n_pad; t += kClassThreads) L[t] = kSentinel;
    __syncthreads();
    for (int q = tid; q < n; q += kClassThreads) {
        if (!alive8[q]) continue;
        const unsigned long long e = KS[q];
        if ((e & key_mask) <= limit) L[atomicAdd(&s_total, 1)] = ((e & key_mask) << 10) | (e >> 54);
    }
    __syncthreads();
    YB_MARK(7);
    for (int k = 2; k <= n_pad; k <<= 1) bitonic_tile_steps(L, n_pad, 0, k, k >> 1);
    YB_MARK(8);
    for (int r = tid; r < n_sel; r += kClassThreads) {
        const unsigned long long e = L[r];
        const int c = (int)(e & 1023u);
        const unsigned long long k54 = e >> 10;
        const unsigned a = (unsigned)(k54 & ((1u << kAnchorBits) - 1u));
        const unsigned inv = (unsigned)(k54 >> kAnchorBits);
        emit_row(out_rows, out_anchor, img, max_det, r, img_pred, n_anchors, (int)a, __uint_as_float(0xFFFFFFFFu - inv), c);
    }
    YB_MARK(9);
    if (tid == 0) { out_count[img] = n_sel; mode[img] = 1; }
}
// ------------------------------------------------------------------------------------------
// multi_label=True (model_utils.py:240-242): every (anchor, class) pair whose score exceeds conf is a
// candidate, up to A * nc per image, of which the reference keeps the max_nms = 30000 best (:211, :259).
// Candidates are never all materialised: two histogram passes over the scores (12 + 12 bits of the
// inverted score word) find, per image, the 24-bit prefix of the 30000-th best score; a third pass emits
// the keys at or above it (30000 plus the few that share that prefix), and the generic sort + sweep kernel
// finishes, cutting at 30000 in (score, anchor * nc + class) order.
// ------------------------------------------------------------------------------------------
constexpr int kMlBins = 4096;
constexpr int kMlCap = 65536;          // keys per image (>= kMaxNms + the boundary bin)
constexpr int kMlThreads = 128;

struct MlWorkspace {
    int *count;                  // [N] keys emitted                         } zeroed every call
    int *hist1, *hist2;          // [N * kMlBins]                            }
    int *thr;                    // [N * 4]  b1, need1 (still wanted inside bin b1), b2, -
    unsigned long long *keys;    // [N * kMlCap]
    float4 *sbox;                // [N * kMaxNms]
    size_t zero_bytes, total_bytes;
};

static MlWorkspace carve_ml(void *base, int n_images) {
    MlWorkspace w;
    char *p = static_cast<char *>(base);
    size_t off = 0;
    w.count = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images, 64);
    w.hist1 = reinterpret_cast<int *>(p + off);
    off += sizeof(int) * (size_t)n_images * kMlBins;
    w.hist2 = reinterpret_cast<int *>(p + off);
    off += sizeof(int) * (size_t)n_images * kMlBins;
    w.zero_bytes = off;
    w.thr = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * 4 * (size_t)n_images, 64);
    w.keys = reinterpret_cast<unsigned long long *>(p + off);
    off += sizeof(unsigned long long) * (size_t)n_images * kMlCap;
    w.sbox = reinterpret_cast<float4 *>(p + off);
    off += sizeof(float4) * (size_t)n_images * kMaxNms;
    w.total_bytes = off;
    return w;
}

// PASS 0: histogram of the top 12 bits of every candidate's inverted score word
// PASS 1: histogram of the next 12 bits, candidates of the image's boundary bin b1 only
// PASS 2: emit the keys with (bin1 < b1) or (bin1 == b1 and bin2 <= b2)
template <int PASS>
__global__ void __launch_bounds__(kMlThreads)
ml_pass_kernel(const float *__restrict__ pred, int nc, int n_anchors, float conf, const int *__restrict__ filter, int n_filter,
               int *__restrict__ hist1, int *__restrict__ hist2, const int *__restrict__ thr, int *__restrict__ count,
               unsigned long long *__restrict__ keys) {
    __shared__ int s_hist[PASS < 2 ? kMlBins : 1];
    const int n = blockIdx.y;
    const int a = blockIdx.x * kMlThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (PASS < 2) {
        for (int b = threadIdx.x; b < kMlBins; b += kMlThreads) s_hist[b] = 0;
        __syncthreads();
    }
    const int b1 = PASS > 0 ? thr[4 * n] : 0, b2 = PASS > 1 ? thr[4 * n + 2] : 0;
    const float *img = pred + ((size_t)n * (4 + nc) + 4) * n_anchors;
    for (int c = 0; c < nc; ++c) {                         // warp-uniform trip count: the emit pass votes
        bool ok = false;
        unsigned inv = 0;
        if (a < n_anchors) {
            const float sc = __ldg(img + (size_t)c * n_anchors + a);
            ok = sc > conf;                                // strict (model_utils.py:241)
            inv = 0xFFFFFFFFu - __float_as_uint(sc);
        }
        if (ok && n_filter > 0) {
            bool in = false;
            for (int f = 0; f < n_filter; ++f) in |= (__ldg(filter + f) == c);
            ok = in;
        }
        const int bin1 = (int)(inv >> 20), bin2 = (int)((inv >> 8) & 0xFFFu);
        if (PASS == 0) {
            if (ok) atomicAdd(&s_hist[bin1], 1);
        } else if (PASS == 1) {
            if (ok && bin1 == b1) atomicAdd(&s_hist[bin2], 1);
        } else {
            ok = ok && (bin1 < b1 || (bin1 == b1 && bin2 <= b2));
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (mask) {
                int slot = 0;
                if (lane == 0) slot = atomicAdd(count + n, __popc(mask));
                slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(mask & ((1u << lane) - 1u));
                if (ok && slot < kMlCap)
                    keys[(size_t)n * kMlCap + slot] = ((unsigned long long)inv << 32) | ((unsigned)a * (unsigned)nc + (unsigned)c);
            }
        }
    }
    if (PASS < 2) {
        __syncthreads();
        int *h = (PASS == 0 ? hist1 : hist2) + (size_t)n * kMlBins;
        for (int b = threadIdx.x; b < kMlBins; b += kMlThreads)
            if (s_hist[b]) atomicAdd(h + b, s_hist[b]);
    }
}

// one warp per image: the bin in which the running count (best scores first) reaches `want`
template <int LEVEL>
__global__ void __launch_bounds__(32)
ml_threshold_kernel(const int *__restrict__ hist1, const int *__restrict__ hist2, int *__restrict__ thr) {
    const int n = blockIdx.x, lane = threadIdx.x;
    const int *h = (LEVEL == 0 ? hist1 : hist2) + (size_t)n * kMlBins;
    const int want = LEVEL == 0 ? kMaxNms : thr[4 * n + 1];
    constexpr int PER = kMlBins / 32;
    int local = 0;
    for (int b = 0; b < PER; ++b) local += h[lane * PER + b];
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total < want || want <= 0) {                       // fewer candidates than wanted: keep them all
        if (lane == 0) {
            thr[4 * n + 2 * LEVEL] = kMlBins;              // every bin index is below it
            if (LEVEL == 0) { thr[4 * n + 1] = 0; thr[4 * n + 2] = kMlBins; }
        }
        return;
    }
    const int excl = incl - local;
    if (excl < want && want <= incl) {                     // exactly one lane
        int run = excl;
        for (int b = 0; b < PER; ++b) {
            const int v = h[lane * PER + b];
            if (run + v >= want) {
                thr[4 * n + 2 * LEVEL] = lane * PER + b;
                if (LEVEL == 0) thr[4 * n + 1] = want - run;
                break;
            }
            run += v;
        }
    }
}

#ifdef YB_NMS_PROFILE
extern "C" int yb_nms_profile_read(long long *out_host) {
    return (int)cudaMemcpyFromSymbol(out_host, g_nms_clk, sizeof(long long) * 16);
}
extern "C" int yb_nms_profile_read_warps(long long *out_host) {
    long long z[128] = {0};
    cudaMemcpyFromSymbol(out_host, g_nms_warp, sizeof(long long) * 128);
    return (int)cudaMemcpyToSymbol(g_nms_warp, z, sizeof(z));
}
#endif

}  // namespace yb

using namespace yb;

extern "C" size_t yb_nms_workspace_bytes(int n_images, int n_anchors) {
    if (n_images <= 0 || n_anchors <= 0) return 0;
    return carve_nms(nullptr, n_images, n_anchors).total_bytes;
}

// scan + class-parallel kernel + generic sweep on whatever layout `in` describes
template <typename T, bool SIGMOID>
static int launch_scan(const NmsInput &in, int n_images, int nc, int n_anchors, float conf_thres, const int32_t *class_filter,
                       int n_class_filter, const NmsWorkspace &w, cudaStream_t st) {
    constexpr int VW = ElemsPer16<T>::value;
    const T *score = static_cast<const T *>(in.score);
    if (n_anchors % VW == 0 && aligned16(score) && (in.score_stride * sizeof(T)) % 16 == 0) {
        dim3 grid((n_anchors / VW + kScanThreads - 1) / kScanThreads, n_images);
        nms_scan_kernel<T, VW, SIGMOID><<<grid, kScanThreads, 0, st>>>(score, in.score_stride, nc, n_anchors, conf_thres, class_filter,
                                                                       n_class_filter, w.count, w.cls, w.keys, w.a_pad);
    } else {
        dim3 grid((n_anchors + kScanThreads - 1) / kScanThreads, n_images);
        nms_scan_kernel<T, 1, SIGMOID><<<grid, kScanThreads, 0, st>>>(score, in.score_stride, nc, n_anchors, conf_thres, class_filter,
                                                                      n_class_filter, w.count, w.cls, w.keys, w.a_pad);
    }
    YB_LAUNCH_CHECK();
    return YB_OK;
}

static int run_nms(const NmsInput &in, int score_dtype, int sigmoid, int n_images, int nc, int n_anchors, float conf_thres,
                   double iou_thres, int max_det, int agnostic, const int32_t *class_filter, int n_class_filter,
                   float *out_rows, int32_t *out_count, int32_t *out_anchor, const NmsWorkspace &w, cudaStream_t st) {
    // "float IoU promoted to double > thr"  <=>  "IoU > largest float that is <= thr"
    float thr = (float)iou_thres;
    if ((double)thr > iou_thres) thr = nextafterf(thr, -INFINITY);

    YB_CUDA(cudaMemsetAsync(w.count, 0, w.zero_bytes, st));
    int rc;
    if (score_dtype == YB_F32)
        rc = sigmoid ? launch_scan<float, true>(in, n_images, nc, n_anchors, conf_thres, class_filter, n_class_filter, w, st)
                     : launch_scan<float, false>(in, n_images, nc, n_anchors, conf_thres, class_filter, n_class_filter, w, st);
    else
        rc = sigmoid ? launch_scan<__nv_bfloat16, true>(in, n_images, nc, n_anchors, conf_thres, class_filter, n_class_filter, w, st)
                     : launch_scan<__nv_bfloat16, false>(in, n_images, nc, n_anchors, conf_thres, class_filter, n_class_filter, w, st);
    if (rc != YB_OK) return rc;
    if (!agnostic && nc <= kClassMaxNc && n_anchors < (1 << kAnchorBits) && max_det <= kClassMaxDet) {
        const size_t csmem = sizeof(float4) * (size_t)kClassCap;      // boxes; later keys + histogram + selection
        YB_CUDA(cudaFuncSetAttribute(nms_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        // CTAs per image: enough to put every SM to work on small batches
        const int split = n_images >= 96 ? 1 : (n_images >= 48 ? 2 : (n_images >= 24 ? 4 : 8));
        YB_CUDA(launch_pdl(nms_class_kernel, dim3(split, n_images), dim3(kClassThreads), csmem, st, in.box, in.box_stride, nc, n_anchors,
                           w.count, w.cls, w.keys, w.a_pad, w.keys2, w.alive_g, w.tick, w.range, thr, max_det, w.mode, out_rows,
                           out_count, out_anchor));
        YB_LAUNCH_CHECK();
    }
    // generic path for whatever the class-parallel kernel left (mode == 0)
    const size_t smem = sizeof(unsigned long long) * (size_t)min(w.a_pad, kSortTile);
    if (n_anchors <= kRegCols * kSortThreads) {
        YB_CUDA(cudaFuncSetAttribute(nms_sweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        YB_CUDA(launch_pdl(nms_sweep_kernel<true, false>, dim3(n_images), dim3(kSortThreads), smem, st, in.box, in.box_stride, nc, n_anchors,
                           w.count, w.mode, w.cls, w.keys, w.a_pad, w.sbox, n_anchors, thr, max_det, agnostic, out_rows, out_count,
                           out_anchor));
    } else {
        YB_CUDA(cudaFuncSetAttribute(nms_sweep_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        YB_CUDA(launch_pdl(nms_sweep_kernel<false, false>, dim3(n_images), dim3(kSortThreads), smem, st, in.box, in.box_stride, nc, n_anchors,
                           w.count, w.mode, w.cls, w.keys, w.a_pad, w.sbox, n_anchors, thr, max_det, agnostic, out_rows, out_count,
                           out_anchor));
    }
    YB_LAUNCH_CHECK();
    return YB_OK;
}

extern "C" int yb_nms(const float *prediction, int n_images, int nc, int n_anchors, float conf_thres, double iou_thres,
                      int max_det, int agnostic, const int32_t *class_filter, int n_class_filter, float *out_rows,
                      int32_t *out_count, int32_t *out_anchor, void *workspace, size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_nms");
    YB_REQUIRE(prediction && out_rows && out_count && workspace, "yb_nms: null pointer");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0 && max_det > 0 && n_images <= 65535, "yb_nms: bad sizes");
    YB_REQUIRE(n_class_filter >= 0 && (n_class_filter == 0 || class_filter), "yb_nms: bad class filter");
    YB_REQUIRE(conf_thres >= 0.f, "yb_nms: conf_thres must be >= 0 (scores are ordered by their bit pattern)");
    if (workspace_bytes < yb_nms_workspace_bytes(n_images, n_anchors)) {
        set_error("yb_nms: workspace %zu B < required %zu B", workspace_bytes, yb_nms_workspace_bytes(n_images, n_anchors));
        return YB_ERR_WORKSPACE;
    }
    if (!aligned16(workspace)) {
        set_error("yb_nms: workspace must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    const NmsWorkspace w = carve_nms(workspace, n_images, n_anchors);
    NmsInput in;
    in.box = prediction;
    in.box_stride = (size_t)(4 + nc) * n_anchors;
    in.score = prediction + (size_t)4 * n_anchors;
    in.score_stride = in.box_stride;
    return run_nms(in, YB_F32, 0, n_images, nc, n_anchors, conf_thres, iou_thres, max_det, agnostic, class_filter, n_class_filter,
                   out_rows, out_count, out_anchor, w, static_cast<cudaStream_t>(stream));
}

// Model.inference after the network (src/model/model_builder.py:123-139) as one entry point: the head output's box
// channels are decoded (DFL expectation -> dist2bbox xywh -> x stride) into a compact (N, 4, A) buffer of the workspace,
// and the NMS kernels read their scores straight from the head output's class channels -- the (N, 4 + nc, A) tensor the
// reference concatenates (:136) never exists.
static size_t postprocess_box_bytes(int n_images, int n_anchors) { return round_up(sizeof(float) * 4 * (size_t)n_images * n_anchors, 64); }

extern "C" size_t yb_postprocess_workspace_bytes(int n_images, int n_anchors) {
    if (n_images <= 0 || n_anchors <= 0) return 0;
    return postprocess_box_bytes(n_images, n_anchors) + carve_nms(nullptr, n_images, n_anchors).total_bytes;
}

extern "C" int yb_postprocess(const void *head_out, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                              const float *anchors, const float *strides, int apply_sigmoid, float conf_thres,
                              double iou_thres, int max_det, int agnostic, const int32_t *class_filter, int n_class_filter,
                              float *out_rows, int32_t *out_count, int32_t *out_anchor, void *workspace,
                              size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_postprocess");
    YB_REQUIRE(head_out && anchors && strides && out_rows && out_count && workspace, "yb_postprocess: null pointer");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0 && max_det > 0 && n_images <= 65535, "yb_postprocess: bad sizes");
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_postprocess: dtype must be YB_F32 or YB_BF16");
    YB_REQUIRE(n_class_filter >= 0 && (n_class_filter == 0 || class_filter), "yb_postprocess: bad class filter");
    YB_REQUIRE(conf_thres >= 0.f, "yb_postprocess: conf_thres must be >= 0 (scores are ordered by their bit pattern)");
    if (workspace_bytes < yb_postprocess_workspace_bytes(n_images, n_anchors)) {
        set_error("yb_postprocess: workspace %zu B < required %zu B", workspace_bytes, yb_postprocess_workspace_bytes(n_images, n_anchors));
        return YB_ERR_WORKSPACE;
    }
    if (!aligned16(workspace)) {
        set_error("yb_postprocess: workspace must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    float *boxes = static_cast<float *>(workspace);
    const size_t image_stride = (size_t)(4 * reg_max + nc) * n_anchors;
    // DFL.forward + dist2bbox(xywh) + "* strides" (model_builder.py:127-133), fp32 boxes whatever the head's dtype
    if (int rc = yb_dfl_decode(head_out, dtype, n_images, reg_max, n_anchors, image_stride, anchors, strides, nullptr, boxes,
                               0 /* xywh */, 1, stream))
        return rc;
    const NmsWorkspace w = carve_nms(static_cast<char *>(workspace) + postprocess_box_bytes(n_images, n_anchors), n_images, n_anchors);
    const size_t esz = dtype == YB_BF16 ? 2 : 4;
    NmsInput in;
    in.box = boxes;
    in.box_stride = (size_t)4 * n_anchors;
    in.score = static_cast<const char *>(head_out) + esz * (size_t)4 * reg_max * n_anchors;
    in.score_stride = image_stride;
    return run_nms(in, dtype, apply_sigmoid, n_images, nc, n_anchors, conf_thres, iou_thres, max_det, agnostic, class_filter,
                   n_class_filter, out_rows, out_count, out_anchor, w, static_cast<cudaStream_t>(stream));
}

extern "C" size_t yb_nms_multilabel_workspace_bytes(int n_images) {
    if (n_images <= 0) return 0;
    return carve_ml(nullptr, n_images).total_bytes;
}

extern "C" int yb_nms_multilabel(const float *prediction, int n_images, int nc, int n_anchors, float conf_thres,
                                 double iou_thres, int max_det, int agnostic, const int32_t *class_filter,
                                 int n_class_filter, float *out_rows, int32_t *out_count, int32_t *out_anchor,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_nms_multilabel");
    YB_REQUIRE(prediction && out_rows && out_count && workspace, "yb_nms_multilabel: null pointer");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0 && max_det > 0 && n_images <= 65535, "yb_nms_multilabel: bad sizes");
    YB_REQUIRE((long long)n_anchors * nc < (1ll << 32), "yb_nms_multilabel: anchors x classes must fit 32 bits");
    YB_REQUIRE(n_class_filter >= 0 && (n_class_filter == 0 || class_filter), "yb_nms_multilabel: bad class filter");
    YB_REQUIRE(conf_thres >= 0.f, "yb_nms_multilabel: conf_thres must be >= 0 (scores are ordered by their bit pattern)");
    if (workspace_bytes < yb_nms_multilabel_workspace_bytes(n_images)) {
        set_error("yb_nms_multilabel: workspace %zu B < required %zu B", workspace_bytes,
                  yb_nms_multilabel_workspace_bytes(n_images));
        return YB_ERR_WORKSPACE;
    }
    if (!aligned16(workspace)) {
        set_error("yb_nms_multilabel: workspace must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    const MlWorkspace w = carve_ml(workspace, n_images);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float thr = (float)iou_thres;
    if ((double)thr > iou_thres) thr = nextafterf(thr, -INFINITY);
    YB_CUDA(cudaMemsetAsync(w.count, 0, w.zero_bytes, st));
    const dim3 grid((n_anchors + kMlThreads - 1) / kMlThreads, n_images);
    ml_pass_kernel<0><<<grid, kMlThreads, 0, st>>>(prediction, nc, n_anchors, conf_thres, class_filter, n_class_filter, w.hist1,
                                                   w.hist2, w.thr, w.count, w.keys);
    YB_LAUNCH_CHECK();
    ml_threshold_kernel<0><<<n_images, 32, 0, st>>>(w.hist1, w.hist2, w.thr);
    YB_LAUNCH_CHECK();
    ml_pass_kernel<1><<<grid, kMlThreads, 0, st>>>(prediction, nc, n_anchors, conf_thres, class_filter, n_class_filter, w.hist1,
                                                   w.hist2, w.thr, w.count, w.keys);
    YB_LAUNCH_CHECK();
    ml_threshold_kernel<1><<<n_images, 32, 0, st>>>(w.hist1, w.hist2, w.thr);
    YB_LAUNCH_CHECK();
    ml_pass_kernel<2><<<grid, kMlThreads, 0, st>>>(prediction, nc, n_anchors, conf_thres, class_filter, n_class_filter, w.hist1,
                                                   w.hist2, w.thr, w.count, w.keys);
    YB_LAUNCH_CHECK();
    const size_t smem = sizeof(unsigned long long) * (size_t)kSortTile;
    YB_CUDA(cudaFuncSetAttribute(nms_sweep_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_sweep_kernel<false, true><<<n_images, kSortThreads, smem, st>>>(prediction, (size_t)(4 + nc) * n_anchors, nc, n_anchors, w.count, nullptr, nullptr,
                                                                        w.keys, kMlCap, w.sbox, kMaxNms, thr, max_det, agnostic,
                                                                        out_rows, out_count, out_anchor);
    YB_LAUNCH_CHECK();
    return YB_OK;
}
