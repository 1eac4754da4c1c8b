// Batched class-aware NMS (sm_100a).  Replaces non_max_suppression (src/utils/model_utils.py:174-279)
// and the per-image torchvision.ops.nms call inside it, for all images of a batch in three launches:
//
//   nms_scan_kernel    reads (N, 4+nc, A) once with 128-bit loads: best class (first max), strict
//                      `> conf`, optional class filter; candidates are compacted with one atomic per
//                      warp into 64-bit (score, anchor) keys.
//   nms_sort_kernel    one CTA per image: bitonic sort of the keys in shared memory
//                      (score descending, ties -> lowest anchor).
//   nms_sweep_kernel   one CTA per image: greedy suppression in score order.  Each thread keeps its
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
    int a_pad;
    size_t zero_bytes, total_bytes;
};

static NmsWorkspace carve_nms(void *base, int n_images, int n_anchors) {
    NmsWorkspace w;
    char *p = static_cast<char *>(base);
    size_t off = 0;
    w.count = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images, 64);
    w.zero_bytes = off;
    w.cls = reinterpret_cast<int *>(p + off);
    off += round_up(sizeof(int) * (size_t)n_images * n_anchors, 64);
    w.a_pad = next_pow2(n_anchors);
    w.keys = reinterpret_cast<unsigned long long *>(p + off);
    off += sizeof(unsigned long long) * (size_t)n_images * w.a_pad;
    w.sbox = reinterpret_cast<float4 *>(p + off);
    off += (n_anchors > kRegCols * kSortThreads) ? sizeof(float4) * (size_t)n_images * n_anchors : 0;
    w.total_bytes = off;
    return w;
}

template <int VW>
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const float *__restrict__ pred, int nc, int n_anchors, float conf, const int *__restrict__ filter,
                int n_filter, int *__restrict__ count, int *__restrict__ cls_out, unsigned long long *__restrict__ keys,
                int a_pad) {
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
        const size_t base = ((size_t)n * (4 + nc) + 4) * n_anchors + a0;
        {
            Group<float, VW> row;
            row.load(pred + base);
#pragma unroll
            for (int v = 0; v < VW; ++v) { best[v] = row.get(v); arg[v] = 0; }
        }
        constexpr int U = 4;
        int c = 1;
        for (; c + U <= nc; c += U) {
            Group<float, VW> row[U];
#pragma unroll
            for (int u = 0; u < U; ++u) row[u].load(pred + base + (size_t)(c + u) * n_anchors);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    const float s = row[u].get(v);
                    if (s > best[v]) { best[v] = s; arg[v] = c + u; }      // first maximum wins
                }
        }
        for (; c < nc; ++c) {
            Group<float, VW> row;
            row.load(pred + base + (size_t)c * n_anchors);
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                const float s = row.get(v);
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

__global__ void __launch_bounds__(kSortThreads)
nms_sort_kernel(const int *__restrict__ count, unsigned long long *__restrict__ keys, int a_pad) {
    extern __shared__ unsigned long long s_keys[];
    const int n = blockIdx.x;
    const int cnt = count[n];
    if (cnt <= 1) return;
    unsigned long long *k = keys + (size_t)n * a_pad;
    const int n_pad = next_pow2(cnt);
    for (int t = cnt + threadIdx.x; t < n_pad; t += blockDim.x) k[t] = kSentinel;
    __syncthreads();
    cta_bitonic_sort(k, n_pad, s_keys);
}

// torchvision's IoU test on two xyxy boxes; thr is the largest float <= the double threshold
__device__ __forceinline__ bool iou_exceeds(const float4 &a, float area_a, const float4 &b, float thr) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ovr > thr;
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

// Resolve one block of 32 consecutive sorted boxes inside a warp.  `bx`/`alive` belong to lane r =
// position c0 + r.  Returns the mask of rows kept (already capped to `room`).
__device__ __forceinline__ unsigned resolve_block(const float4 &bx, bool alive, float thr, int room) {
    const int lane = threadIdx.x & 31;
    unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
    const float my_area = box_area(bx);
    unsigned sup = 0;                                    // bit r2: my box suppresses the later box r2
#pragma unroll 4
    for (int r2 = 1; r2 < 32; ++r2) {
        float4 o;
        o.x = __shfl_sync(0xffffffffu, bx.x, r2);
        o.y = __shfl_sync(0xffffffffu, bx.y, r2);
        o.z = __shfl_sync(0xffffffffu, bx.z, r2);
        o.w = __shfl_sync(0xffffffffu, bx.w, r2);
        if (r2 > lane && iou_exceeds(bx, my_area, o, thr)) sup |= 1u << r2;
    }
    unsigned kept = 0;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const unsigned s_r = __shfl_sync(0xffffffffu, sup, r);
        if ((alive_mask >> r) & 1u) {
            kept |= 1u << r;
            alive_mask &= ~s_r;
        }
    }
    while (__popc(kept) > room) kept &= ~(0x80000000u >> __clz(kept));   // drop the lowest-ranked keeps
    return kept;
}

__device__ __forceinline__ void emit_row(float *__restrict__ out_rows, int *__restrict__ out_anchor, int n, int max_det,
                                         int pos, const float *__restrict__ img, int n_anchors, unsigned long long key,
                                         int cls) {
    const int a = (int)key_anchor(key);
    const float x = __ldg(img + a), y = __ldg(img + n_anchors + a);
    const float dw = __fmul_rn(__ldg(img + 2 * (size_t)n_anchors + a), 0.5f);
    const float dh = __fmul_rn(__ldg(img + 3 * (size_t)n_anchors + a), 0.5f);
    float *row = out_rows + ((size_t)n * max_det + pos) * 6;
    row[0] = __fsub_rn(x, dw);
    row[1] = __fsub_rn(y, dh);
    row[2] = __fadd_rn(x, dw);
    row[3] = __fadd_rn(y, dh);
    row[4] = key_score(key);
    row[5] = (float)cls;
    if (out_anchor) out_anchor[(size_t)n * max_det + pos] = a;
}

// REG = true: up to kRegCols*1024 candidates, boxes and alive bits in registers.
// REG = false: any count up to kMaxNms, boxes in a global scratch, alive bits in shared memory.
template <bool REG>
__global__ void __launch_bounds__(kSortThreads, 1)
nms_sweep_kernel(const float *__restrict__ pred, int nc, int n_anchors, const int *__restrict__ count,
                 const int *__restrict__ cls, const unsigned long long *__restrict__ keys, int a_pad,
                 float4 *__restrict__ sbox, float thr, int max_det, int agnostic, float *__restrict__ out_rows,
                 int *__restrict__ out_count, int *__restrict__ out_anchor) {
    __shared__ float4 s_row[32];
    __shared__ float s_area[32];
    __shared__ int s_nrow;
    __shared__ unsigned s_alive[REG ? 1 : (kMaxNms + 31) / 32 + 1];

    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_cand = min(count[n], kMaxNms);
    if (n_cand == 0 || max_det <= 0) {
        if (tid == 0) out_count[n] = 0;
        return;
    }
    if (REG && n_cand > kRegCols * kSortThreads) return;      // host picks the other variant; never taken
    const float *img = pred + (size_t)n * (4 + nc) * n_anchors;
    const unsigned long long *k = keys + (size_t)n * a_pad;
    const int *cls_n = cls + (size_t)n * n_anchors;

    float4 box[REG ? kRegCols : 1];
    unsigned alive = 0;
    if (REG) {
#pragma unroll
        for (int t = 0; t < kRegCols; ++t) {
            const int j = tid + t * kSortThreads;
            box[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n_cand) {
                const int a = (int)key_anchor(k[j]);
                box[t] = load_offset_box(img, n_anchors, a, cls_n[a], agnostic != 0);
                alive |= 1u << t;
            }
        }
    } else {
        float4 *sb = sbox + (size_t)n * n_anchors;
        for (int j = tid; j < n_cand; j += kSortThreads) {
            const int a = (int)key_anchor(k[j]);
            sb[j] = load_offset_box(img, n_anchors, a, cls_n[a], agnostic != 0);
        }
        for (int w = tid; w < (n_cand + 31) / 32; w += kSortThreads)
            s_alive[w] = (w * 32 + 32 <= n_cand) ? 0xffffffffu : ((1u << (n_cand - w * 32)) - 1u);
    }
    __syncthreads();
    const float4 *sb = REG ? nullptr : sbox + (size_t)n * n_anchors;

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
                emit_row(out_rows, out_anchor, n, max_det, kept_total + rank, img, n_anchors, key,
                         cls_n[key_anchor(key)]);
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

}  // namespace yb

using namespace yb;

extern "C" size_t yb_nms_workspace_bytes(int n_images, int n_anchors) {
    if (n_images <= 0 || n_anchors <= 0) return 0;
    return carve_nms(nullptr, n_images, n_anchors).total_bytes;
}

extern "C" int yb_nms(const float *prediction, int n_images, int nc, int n_anchors, float conf_thres, double iou_thres,
                      int max_det, int agnostic, const int32_t *class_filter, int n_class_filter, float *out_rows,
                      int32_t *out_count, int32_t *out_anchor, void *workspace, size_t workspace_bytes, void *stream) {
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // "float IoU promoted to double > thr"  <=>  "IoU > largest float that is <= thr"
    float thr = (float)iou_thres;
    if ((double)thr > iou_thres) thr = nextafterf(thr, -INFINITY);

    YB_CUDA(cudaMemsetAsync(w.count, 0, w.zero_bytes, st));
    if (n_anchors % 4 == 0 && aligned16(prediction)) {
        dim3 grid((n_anchors / 4 + kScanThreads - 1) / kScanThreads, n_images);
        nms_scan_kernel<4><<<grid, kScanThreads, 0, st>>>(prediction, nc, n_anchors, conf_thres, class_filter,
                                                          n_class_filter, w.count, w.cls, w.keys, w.a_pad);
    } else {
        dim3 grid((n_anchors + kScanThreads - 1) / kScanThreads, n_images);
        nms_scan_kernel<1><<<grid, kScanThreads, 0, st>>>(prediction, nc, n_anchors, conf_thres, class_filter,
                                                          n_class_filter, w.count, w.cls, w.keys, w.a_pad);
    }
    YB_CUDA(cudaGetLastError());
    const size_t smem = sizeof(unsigned long long) * (size_t)min(w.a_pad, kSortTile);
    YB_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_sort_kernel<<<n_images, kSortThreads, smem, st>>>(w.count, w.keys, w.a_pad);
    YB_CUDA(cudaGetLastError());
    if (n_anchors <= kRegCols * kSortThreads)
        nms_sweep_kernel<true><<<n_images, kSortThreads, 0, st>>>(prediction, nc, n_anchors, w.count, w.cls, w.keys,
                                                                  w.a_pad, w.sbox, thr, max_det, agnostic, out_rows,
                                                                  out_count, out_anchor);
    else
        nms_sweep_kernel<false><<<n_images, kSortThreads, 0, st>>>(prediction, nc, n_anchors, w.count, w.cls, w.keys,
                                                                   w.a_pad, w.sbox, thr, max_det, agnostic, out_rows,
                                                                   out_count, out_anchor);
    YB_CUDA(cudaGetLastError());
    return YB_OK;
}
