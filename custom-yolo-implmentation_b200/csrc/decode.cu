// DFL softmax-expectation decode kernels (sm_100a).
//
//   dfl_decode_kernel   DFL.forward (src/model/model_blocks.py:278-280) fused with dist2bbox
//                       (src/utils/model_utils.py:120-129) and the stride multiply
//                       (src/model/model_builder.py:133 / src/training/train_model.py:109)
//   dist2bbox_kernel    dist2bbox alone, dim=1 layout
//   val_decode_*        decode_predictions (src/training/train_model.py:14-142): per anchor the best
//                       sigmoid score and its class, `>= conf` compaction, per-image top-k.
#include "sort.cuh"

namespace yb {

constexpr int kDecThreads = 128;

// dist2bbox arithmetic, in the reference's operation order (model_utils.py:122-129)
__device__ __forceinline__ float4 ltrb_to_box(float ax, float ay, float l, float t, float r, float b, bool xywh) {
    const float x1 = __fsub_rn(ax, l), y1 = __fsub_rn(ay, t), x2 = __fadd_rn(ax, r), y2 = __fadd_rn(ay, b);
    if (!xywh) return make_float4(x1, y1, x2, y2);
    return make_float4(__fmul_rn(__fadd_rn(x1, x2), 0.5f), __fmul_rn(__fadd_rn(y1, y2), 0.5f), __fsub_rn(x2, x1),
                       __fsub_rn(y2, y1));
}

template <typename T, int VW>
__global__ void __launch_bounds__(kDecThreads)
dfl_decode_kernel(const T *__restrict__ box_logits, size_t image_stride, int n_anchors,
                  const float *__restrict__ anchors, const float *__restrict__ strides, float *__restrict__ out_ltrb,
                  float *__restrict__ out_box, int xywh, int scale) {
    const int n = blockIdx.y;
    const int a0 = (blockIdx.x * kDecThreads + threadIdx.x) * VW;
    if (a0 >= n_anchors) return;
    const T *img = box_logits + (size_t)n * image_stride;
    float d[4][VW];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        Group<T, VW> row[kRegMax];
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) row[j].load(img + (size_t)(side * kRegMax + j) * n_anchors + a0);
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            float x[kRegMax], p[kRegMax];
#pragma unroll
            for (int j = 0; j < kRegMax; ++j) x[j] = row[j].get(v);
            d[side][v] = dfl_expectation16(x, p);
        }
    }
    const size_t o = (size_t)n * 4 * n_anchors + a0;
#pragma unroll
    for (int v = 0; v < VW; ++v) {
        if (out_ltrb) {
#pragma unroll
            for (int side = 0; side < 4; ++side) out_ltrb[o + (size_t)side * n_anchors + v] = d[side][v];
        }
        if (out_box) {
            const float ax = __ldg(anchors + a0 + v), ay = __ldg(anchors + n_anchors + a0 + v);
            float4 b = ltrb_to_box(ax, ay, d[0][v], d[1][v], d[2][v], d[3][v], xywh != 0);
            if (scale) {
                const float s = __ldg(strides + a0 + v);
                b.x = __fmul_rn(b.x, s); b.y = __fmul_rn(b.y, s); b.z = __fmul_rn(b.z, s); b.w = __fmul_rn(b.w, s);
            }
            out_box[o + v] = b.x;
            out_box[o + (size_t)n_anchors + v] = b.y;
            out_box[o + (size_t)2 * n_anchors + v] = b.z;
            out_box[o + (size_t)3 * n_anchors + v] = b.w;
        }
    }
}

__global__ void __launch_bounds__(256)
dist2bbox_kernel(const float *__restrict__ ltrb, const float *__restrict__ anchors, int n_anchors, int xywh,
                 float *__restrict__ out) {
    const int n = blockIdx.y;
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= n_anchors) return;
    const size_t o = (size_t)n * 4 * n_anchors + a;
    const float4 b = ltrb_to_box(__ldg(anchors + a), __ldg(anchors + n_anchors + a), __ldg(ltrb + o),
                                 __ldg(ltrb + o + n_anchors), __ldg(ltrb + o + 2 * (size_t)n_anchors),
                                 __ldg(ltrb + o + 3 * (size_t)n_anchors), xywh != 0);
    out[o] = b.x;
    out[o + n_anchors] = b.y;
    out[o + 2 * (size_t)n_anchors] = b.z;
    out[o + 3 * (size_t)n_anchors] = b.w;
}

__global__ void __launch_bounds__(256)
anchor_level_kernel(int h, int w, float stride, float *__restrict__ grid, float *__restrict__ st) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= h * w) return;
    grid[2 * i] = (float)(i % w);
    grid[2 * i + 1] = (float)(i / w);
    st[i] = stride;
}

// ---- validation decode ---------------------------------------------------------------------
struct ValWorkspace {
    int *count;                   // [N] candidates per image (zeroed every call)
    int *cls;                     // [N * A] best class per anchor
    unsigned long long *keys;     // [N * Apad] compacted candidate keys
    int a_pad;
    size_t zero_bytes, total_bytes;
};

static ValWorkspace carve_val(void *base, int n_images, int n_anchors) {
    ValWorkspace w;
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
    w.total_bytes = off;
    return w;
}

// best sigmoid score over the class channels (first maximum wins, as torch.max(dim)) and `>= conf`
template <typename T, int VW>
__global__ void __launch_bounds__(kDecThreads)
val_scan_kernel(const T *__restrict__ preds, int n_ch, int nc, int n_anchors, float conf, int *__restrict__ count,
                int *__restrict__ cls_out, unsigned long long *__restrict__ keys, int a_pad) {
    const int n = blockIdx.y;
    const int a0 = (blockIdx.x * kDecThreads + threadIdx.x) * VW;
    const int lane = threadIdx.x & 31;
    float best[VW];
    int arg[VW];
    int n_pass = 0;
    if (a0 < n_anchors) {
        // The reference takes max / argmax over the SIGMOIDS (train_model.py:116-119): first index among the largest
        // float sigmoid.  The sigmoid is monotone, so the pass tracks the largest LOGIT (first index) and the largest
        // logit before it; one sigmoid per anchor at the end.  Only when that earlier logit rounds to the same (or a
        // larger) sigmoid can an earlier class win the tie: that anchor is then redone the reference's way.
        const float ninf = -__int_as_float(0x7f800000);
        float bx[VW], sx[VW];
#pragma unroll
        for (int v = 0; v < VW; ++v) { bx[v] = ninf; sx[v] = ninf; arg[v] = 0; }
        const size_t base = ((size_t)n * n_ch + 4 * kRegMax) * n_anchors + a0;
        for (int c = 0; c < nc; ++c) {
            Group<T, VW> row;
            row.load(preds + base + (size_t)c * n_anchors);
#pragma unroll
            for (int v = 0; v < VW; ++v) {
                const float x = row.get(v);
                if (x > bx[v]) { sx[v] = bx[v]; bx[v] = x; arg[v] = c; }
            }
        }
        auto sigmoid = [](float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); };       // Tensor.sigmoid()
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            best[v] = bx[v] > ninf ? sigmoid(bx[v]) : -1.f;
            if (sx[v] > ninf && sigmoid(sx[v]) >= best[v]) {                          // rare: a tie in sigmoid space
                best[v] = -1.f;
                arg[v] = 0;
                for (int c = 0; c < nc; ++c) {
                    const float sg = sigmoid(load_as_float(preds + base + (size_t)c * n_anchors + v));
                    if (sg > best[v]) { best[v] = sg; arg[v] = c; }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VW; ++v) {
            cls_out[(size_t)n * n_anchors + a0 + v] = arg[v];
            n_pass += (best[v] >= conf) ? 1 : 0;
        }
    }
    // warp-aggregated slot allocation: one atomic per warp
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
    if (a0 < n_anchors) {
#pragma unroll
        for (int v = 0; v < VW; ++v)
            if (best[v] >= conf) keys[(size_t)n * a_pad + slot++] = make_score_key(best[v], (unsigned int)(a0 + v));
    }
}

// one CTA per image: order the candidates and emit rows [cx, cy, w, h, cls]
template <typename T>
__global__ void __launch_bounds__(kSortThreads)
val_emit_kernel(const T *__restrict__ preds, int n_ch, int n_anchors, const float *__restrict__ anchors,
                const float *__restrict__ strides, const int *__restrict__ count, const int *__restrict__ cls,
                unsigned long long *__restrict__ keys, int a_pad, int top_k, float *__restrict__ out_rows,
                int *__restrict__ out_count, int *__restrict__ out_anchor) {
    extern __shared__ unsigned long long s_keys[];
    const int n = blockIdx.x;
    const int cnt = count[n];
    unsigned long long *k = keys + (size_t)n * a_pad;
    const int n_out = min(cnt, top_k);
    if (threadIdx.x == 0) out_count[n] = n_out;
    if (cnt == 0) return;
    const int n_pad = next_pow2(cnt);
    // at most top_k candidates: the reference keeps them in anchor order (boolean-mask order); otherwise
    // torch.topk order = score descending.  Re-key by anchor in the first case, then one sort does both.
    for (int t = threadIdx.x; t < n_pad; t += blockDim.x) {
        if (t >= cnt) k[t] = kSentinel;
        else if (cnt <= top_k) k[t] = (unsigned long long)key_anchor(k[t]) << 32 | (k[t] >> 32);
    }
    __syncthreads();
    cta_bitonic_sort(k, n_pad, s_keys);
    const T *img = preds + (size_t)n * n_ch * n_anchors;
    // one warp per emitted row: 64 logits -> 4 expectations (lanes 0-15 / 16-31 hold one side each)
    const int lane = threadIdx.x & 31, bin = lane & 15;
    for (int r = threadIdx.x >> 5; r < n_out; r += kSortThreads / 32) {
        const unsigned long long key = k[r];
        const int a = (cnt <= top_k) ? (int)(key >> 32) : (int)key_anchor(key);
        float x_lo = load_as_float(img + (size_t)lane * n_anchors + a);
        float x_hi = load_as_float(img + (size_t)(lane + 32) * n_anchors + a);
        // gather each side's 16 logits into every lane of its half, then reuse the scalar routine so the
        // result is bit-identical to dfl_decode_kernel
        float xl[kRegMax], xh[kRegMax], p[kRegMax];
#pragma unroll
        for (int j = 0; j < kRegMax; ++j) {
            xl[j] = __shfl_sync(0xffffffffu, x_lo, (lane & 16) + j);
            xh[j] = __shfl_sync(0xffffffffu, x_hi, (lane & 16) + j);
        }
        const float d_lo = dfl_expectation16(xl, p), d_hi = dfl_expectation16(xh, p);
        const float dl = __shfl_sync(0xffffffffu, d_lo, 0), dt = __shfl_sync(0xffffffffu, d_lo, 16);
        const float dr = __shfl_sync(0xffffffffu, d_hi, 0), db = __shfl_sync(0xffffffffu, d_hi, 16);
        (void)bin;
        if (lane == 0) {
            const float s = __ldg(strides + a);
            const float4 b = ltrb_to_box(__ldg(anchors + a), __ldg(anchors + n_anchors + a), dl, dt, dr, db, true);
            float *row = out_rows + ((size_t)n * top_k + r) * 5;
            row[0] = __fmul_rn(b.x, s);
            row[1] = __fmul_rn(b.y, s);
            row[2] = __fmul_rn(b.z, s);
            row[3] = __fmul_rn(b.w, s);
            row[4] = (float)cls[(size_t)n * n_anchors + a];
            if (out_anchor) out_anchor[(size_t)n * top_k + r] = a;
        }
    }
}

template <typename T>
static bool vec_ok(const void *p, int n_anchors, size_t image_stride) {
    constexpr int VW = ElemsPer16<T>::value;
    return n_anchors % VW == 0 && image_stride % VW == 0 && aligned16(p);
}

template <typename T>
static int launch_decode(const T *x, int n_images, int n_anchors, size_t image_stride, const float *anchors,
                         const float *strides, float *out_ltrb, float *out_box, int fmt, int scale, cudaStream_t st) {
    constexpr int VW = ElemsPer16<T>::value;
    if (vec_ok<T>(x, n_anchors, image_stride)) {
        dim3 grid((n_anchors / VW + kDecThreads - 1) / kDecThreads, n_images);
        dfl_decode_kernel<T, VW><<<grid, kDecThreads, 0, st>>>(x, image_stride, n_anchors, anchors, strides, out_ltrb,
                                                               out_box, fmt == 0, scale);
    } else {
        dim3 grid((n_anchors + kDecThreads - 1) / kDecThreads, n_images);
        dfl_decode_kernel<T, 1><<<grid, kDecThreads, 0, st>>>(x, image_stride, n_anchors, anchors, strides, out_ltrb,
                                                              out_box, fmt == 0, scale);
    }
    YB_LAUNCH_CHECK();
    return YB_OK;
}

template <typename T>
static int launch_val(const T *preds, int n_images, int nc, int n_anchors, const float *anchors, const float *strides,
                      float conf, int top_k, float *out_rows, int *out_count, int *out_anchor, const ValWorkspace &w,
                      cudaStream_t st) {
    constexpr int VW = ElemsPer16<T>::value;
    const int n_ch = 4 * kRegMax + nc;
    YB_CUDA(cudaMemsetAsync(w.count, 0, w.zero_bytes, st));
    if (vec_ok<T>(preds, n_anchors, (size_t)n_ch * n_anchors)) {
        dim3 grid((n_anchors / VW + kDecThreads - 1) / kDecThreads, n_images);
        val_scan_kernel<T, VW><<<grid, kDecThreads, 0, st>>>(preds, n_ch, nc, n_anchors, conf, w.count, w.cls, w.keys,
                                                             w.a_pad);
    } else {
        dim3 grid((n_anchors + kDecThreads - 1) / kDecThreads, n_images);
        val_scan_kernel<T, 1><<<grid, kDecThreads, 0, st>>>(preds, n_ch, nc, n_anchors, conf, w.count, w.cls, w.keys,
                                                            w.a_pad);
    }
    YB_LAUNCH_CHECK();
    const size_t smem = sizeof(unsigned long long) * (size_t)min(w.a_pad, kSortTile);
    YB_CUDA(cudaFuncSetAttribute(val_emit_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    val_emit_kernel<T><<<n_images, kSortThreads, smem, st>>>(preds, n_ch, n_anchors, anchors, strides, w.count, w.cls,
                                                            w.keys, w.a_pad, top_k, out_rows, out_count, out_anchor);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_dfl_decode(const void *box_logits, int dtype, int n_images, int reg_max, int n_anchors,
                             size_t image_stride, const float *anchors, const float *strides, float *out_ltrb,
                             float *out_box, int box_format, int scale_by_stride, void *stream) {
    YB_NVTX("yb_dfl_decode");
    YB_REQUIRE(box_logits != nullptr, "yb_dfl_decode: null input");
    YB_REQUIRE(out_ltrb || out_box, "yb_dfl_decode: no output requested");
    YB_REQUIRE(!out_box || anchors, "yb_dfl_decode: anchors required for box output");
    YB_REQUIRE(!(out_box && scale_by_stride) || strides, "yb_dfl_decode: strides required when scaling");
    YB_REQUIRE(n_images > 0 && n_anchors > 0 && n_images <= 65535, "yb_dfl_decode: bad sizes");
    YB_REQUIRE(reg_max == kRegMax, "yb_dfl_decode: reg_max must be %d (got %d)", kRegMax, reg_max);
    YB_REQUIRE(box_format == 0 || box_format == 1, "yb_dfl_decode: box_format must be 0 (xywh) or 1 (xyxy)");
    YB_REQUIRE(image_stride >= (size_t)4 * kRegMax * n_anchors, "yb_dfl_decode: image_stride too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == YB_F32)
        return launch_decode<float>((const float *)box_logits, n_images, n_anchors, image_stride, anchors, strides,
                                    out_ltrb, out_box, box_format, scale_by_stride, st);
    if (dtype == YB_BF16)
        return launch_decode<__nv_bfloat16>((const __nv_bfloat16 *)box_logits, n_images, n_anchors, image_stride,
                                            anchors, strides, out_ltrb, out_box, box_format, scale_by_stride, st);
    set_error("yb_dfl_decode: dtype must be YB_F32 or YB_BF16");
    return YB_ERR_ARG;
}

extern "C" int yb_make_anchors(const int32_t *shapes_host, const float *strides_host, int n_levels, float *out_grid,
                               float *out_strides, void *stream) {
    YB_REQUIRE(shapes_host && strides_host && out_grid && out_strides, "yb_make_anchors: null pointer");
    YB_REQUIRE(n_levels > 0, "yb_make_anchors: no levels");
    size_t off = 0;
    for (int l = 0; l < n_levels; ++l) {
        const int h = shapes_host[2 * l], w = shapes_host[2 * l + 1];
        YB_REQUIRE(h > 0 && w > 0, "yb_make_anchors: level %d has shape (%d, %d)", l, h, w);
        anchor_level_kernel<<<(h * w + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            h, w, strides_host[l], out_grid + 2 * off, out_strides + off);
        YB_LAUNCH_CHECK();
        off += (size_t)h * w;
    }
    return YB_OK;
}

extern "C" int yb_dist2bbox(const float *ltrb, const float *anchors, int n_images, int n_anchors, int xywh,
                            float *out_box, void *stream) {
    YB_REQUIRE(ltrb && anchors && out_box, "yb_dist2bbox: null pointer");
    YB_REQUIRE(n_images > 0 && n_anchors > 0 && n_images <= 65535, "yb_dist2bbox: bad sizes");
    dim3 grid((n_anchors + 255) / 256, n_images);
    dist2bbox_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ltrb, anchors, n_anchors, xywh, out_box);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

extern "C" size_t yb_val_decode_workspace_bytes(int n_images, int n_anchors) {
    if (n_images <= 0 || n_anchors <= 0) return 0;
    return carve_val(nullptr, n_images, n_anchors).total_bytes;
}

extern "C" int yb_val_decode(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                             const float *anchors, const float *strides, float conf_thres, int top_k,
                             float *out_rows, int32_t *out_count, int32_t *out_anchor, void *workspace,
                             size_t workspace_bytes, void *stream) {
    YB_NVTX("yb_val_decode");
    YB_REQUIRE(preds && anchors && strides && out_rows && out_count && workspace, "yb_val_decode: null pointer");
    YB_REQUIRE(n_images > 0 && nc > 0 && n_anchors > 0 && top_k > 0 && n_images <= 65535, "yb_val_decode: bad sizes");
    YB_REQUIRE(reg_max == kRegMax, "yb_val_decode: reg_max must be %d (got %d)", kRegMax, reg_max);
    if (workspace_bytes < yb_val_decode_workspace_bytes(n_images, n_anchors)) {
        set_error("yb_val_decode: workspace %zu B < required %zu B", workspace_bytes,
                  yb_val_decode_workspace_bytes(n_images, n_anchors));
        return YB_ERR_WORKSPACE;
    }
    if (!aligned16(workspace)) {
        set_error("yb_val_decode: workspace must be 16-byte aligned");
        return YB_ERR_ALIGN;
    }
    const ValWorkspace w = carve_val(workspace, n_images, n_anchors);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == YB_F32)
        return launch_val<float>((const float *)preds, n_images, nc, n_anchors, anchors, strides, conf_thres, top_k,
                                 out_rows, out_count, out_anchor, w, st);
    if (dtype == YB_BF16)
        return launch_val<__nv_bfloat16>((const __nv_bfloat16 *)preds, n_images, nc, n_anchors, anchors, strides,
                                         conf_thres, top_k, out_rows, out_count, out_anchor, w, st);
    set_error("yb_val_decode: dtype must be YB_F32 or YB_BF16");
    return YB_ERR_ARG;
}
