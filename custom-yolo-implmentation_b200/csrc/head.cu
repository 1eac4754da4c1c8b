// Tail of Head.forward (src/model/head.py:86-121): per level `cat((box_l, cls_l), dim=1)`, flatten
// H x W, `cat(..., dim=2)` over the levels.  The reference copies every element twice (one cat per
// level, one final cat); here one launch moves each element once, and the adjoint (the backward of
// the two cats, which autograd would run as 2*n_levels strided slice copies) is the same kernel
// with the direction flipped.  Pure HBM copy: 2 * sizeof(T) bytes per element.
#include "common.cuh"

namespace yb {

constexpr int kHeadThreads = 256;
constexpr int kHeadIters = 4;          // vectors per thread along the anchor axis
constexpr int kHeadMaxLevels = 8;

struct HeadLevels {
    void *box[kHeadMaxLevels];         // (N, box_ch, hw_l)
    void *cls[kHeadMaxLevels];         // (N, nc, hw_l)
    int begin[kHeadMaxLevels + 1];     // first anchor of level l; begin[n_levels] = A
    int n_levels;
};

template <typename T, int VW>
__device__ __forceinline__ void copy_group(T *dst, const T *src) {
    if constexpr (VW * sizeof(T) == 16)
        stg_stream16(dst, ldg_stream16(src));
    else
        *dst = *src;
}

// row = n * C + c of the (N, C, A) tensor; the level tensors hold the same row at
// ((n * ch_l + c_l) * hw_l).  GATHER: levels -> packed.  !GATHER: packed -> levels.
template <typename T, int VW, bool GATHER>
__global__ void __launch_bounds__(kHeadThreads)
head_tail_kernel(HeadLevels lv, T *__restrict__ packed, int box_ch, int nc, int n_anchors) {
    const int channels = box_ch + nc;
    const int row = blockIdx.x;
    const int n = row / channels, c = row - n * channels;
    const bool is_box = c < box_ch;
    const int cl = is_box ? c : c - box_ch;
    const int chl = is_box ? box_ch : nc;
    T *prow = packed + (size_t)row * n_anchors;
    const int a_base = blockIdx.y * (kHeadThreads * kHeadIters * VW) + threadIdx.x * VW;
#pragma unroll
    for (int it = 0; it < kHeadIters; ++it) {
        const int a = a_base + it * kHeadThreads * VW;
        if (a >= n_anchors) break;
        int l = 0;
#pragma unroll
        for (int k = 1; k < kHeadMaxLevels; ++k)
            if (k < lv.n_levels && a >= lv.begin[k]) l = k;
        const int hw = lv.begin[l + 1] - lv.begin[l];
        T *lrow = static_cast<T *>(is_box ? lv.box[l] : lv.cls[l]) + ((size_t)n * chl + cl) * hw + (a - lv.begin[l]);
        if (GATHER)
            copy_group<T, VW>(prow + a, lrow);
        else
            copy_group<T, VW>(lrow, prow + a);
    }
}

template <typename T, bool GATHER>
static int launch_head(const HeadLevels &lv, void *packed, int n_images, int box_ch, int nc, int n_anchors, bool vec,
                       cudaStream_t stream) {
    constexpr int VW = 16 / sizeof(T);
    const unsigned rows = (unsigned)n_images * (unsigned)(box_ch + nc);
    if (vec) {
        dim3 grid(rows, (n_anchors + kHeadThreads * kHeadIters * VW - 1) / (kHeadThreads * kHeadIters * VW));
        head_tail_kernel<T, VW, GATHER><<<grid, kHeadThreads, 0, stream>>>(lv, static_cast<T *>(packed), box_ch, nc, n_anchors);
    } else {
        dim3 grid(rows, (n_anchors + kHeadThreads * kHeadIters - 1) / (kHeadThreads * kHeadIters));
        head_tail_kernel<T, 1, GATHER><<<grid, kHeadThreads, 0, stream>>>(lv, static_cast<T *>(packed), box_ch, nc, n_anchors);
    }
    YB_LAUNCH_CHECK();
    return YB_OK;
}

static int head_tail(bool gather, void *const *box_levels, void *const *cls_levels, const int32_t *hw_host, int n_levels,
                     int dtype, int n_images, int box_ch, int nc, void *packed, void *stream) {
    YB_REQUIRE(box_levels && cls_levels && hw_host && packed, "yb_head_%s: null pointer", gather ? "gather" : "scatter");
    YB_REQUIRE(n_levels > 0 && n_levels <= kHeadMaxLevels, "yb_head_%s: n_levels must be 1..%d (got %d)",
               gather ? "gather" : "scatter", kHeadMaxLevels, n_levels);
    YB_REQUIRE(n_images > 0 && box_ch > 0 && nc > 0, "yb_head_%s: bad sizes", gather ? "gather" : "scatter");
    YB_REQUIRE(dtype == YB_F32 || dtype == YB_BF16, "yb_head_%s: dtype must be YB_F32 or YB_BF16", gather ? "gather" : "scatter");
    YB_REQUIRE((long long)n_images * (box_ch + nc) < (1ll << 31), "yb_head_%s: too many rows", gather ? "gather" : "scatter");
    const int vw = dtype == YB_F32 ? 4 : 8;
    HeadLevels lv{};
    lv.n_levels = n_levels;
    long long total = 0;
    bool vec = (reinterpret_cast<uintptr_t>(packed) & 15) == 0;
    for (int l = 0; l < n_levels; ++l) {
        YB_REQUIRE(hw_host[l] > 0 && box_levels[l] && cls_levels[l], "yb_head_%s: level %d is empty",
                   gather ? "gather" : "scatter", l);
        lv.box[l] = box_levels[l];
        lv.cls[l] = cls_levels[l];
        lv.begin[l] = (int)total;
        total += hw_host[l];
        YB_REQUIRE(total < (1ll << 31), "yb_head_%s: too many anchors", gather ? "gather" : "scatter");
        vec = vec && hw_host[l] % vw == 0 && (reinterpret_cast<uintptr_t>(box_levels[l]) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(cls_levels[l]) & 15) == 0;
    }
    for (int l = n_levels; l <= kHeadMaxLevels; ++l) lv.begin[l] = (int)total;
    const int n_anchors = (int)total;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == YB_F32)
        return gather ? launch_head<float, true>(lv, packed, n_images, box_ch, nc, n_anchors, vec, s)
                      : launch_head<float, false>(lv, packed, n_images, box_ch, nc, n_anchors, vec, s);
    return gather ? launch_head<__nv_bfloat16, true>(lv, packed, n_images, box_ch, nc, n_anchors, vec, s)
                  : launch_head<__nv_bfloat16, false>(lv, packed, n_images, box_ch, nc, n_anchors, vec, s);
}

}  // namespace yb

extern "C" int yb_head_gather(const void *const *box_levels, const void *const *cls_levels, const int32_t *hw_host,
                              int n_levels, int dtype, int n_images, int box_ch, int nc, void *out, void *stream) {
    YB_NVTX("yb_head_gather");
    return yb::head_tail(true, const_cast<void *const *>(box_levels), const_cast<void *const *>(cls_levels), hw_host,
                         n_levels, dtype, n_images, box_ch, nc, out, stream);
}

extern "C" int yb_head_scatter(const void *grad, const int32_t *hw_host, int n_levels, int dtype, int n_images,
                               int box_ch, int nc, void *const *box_grads, void *const *cls_grads, void *stream) {
    YB_NVTX("yb_head_scatter");
    return yb::head_tail(false, box_grads, cls_grads, hw_host, n_levels, dtype, n_images, box_ch, nc,
                         const_cast<void *>(grad), stream);
}
