// GT wire format on the device: the reference hands the loss a Python list of N small (Mi, 5) device tensors
// (src/training/train_model.py:236, read at src/model/losses.py:206-208); the kernels want ONE (sum Mi, 5) fp32 buffer
// + offsets.  torch.cat over 128 slices costs ~0.4 ms of host time per step; here the host only writes a table
// (pointer, row pitch, first output row per image -- one small host-to-device copy) and one launch gathers the rows.
#include "common.cuh"

namespace yb {

// one warp per image; the table entry of image b: src pointer, row pitch in floats, first output row, row count
__global__ void __launch_bounds__(128) gather_gt_kernel(const yb_gt_source *__restrict__ table, int n_images,
                                                        float *__restrict__ out_gt) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= n_images) return;
    const yb_gt_source s = table[b];
    const float *src = static_cast<const float *>(s.rows);
    float *dst = out_gt + (size_t)s.first_row * 5;
    for (int i = lane; i < s.n_rows * 5; i += 32) {
        const int r = i / 5, c = i - r * 5;
        dst[i] = src[(size_t)r * s.row_pitch + c];
    }
}

}  // namespace yb

using namespace yb;

extern "C" int yb_gather_gt(const yb_gt_source *table_dev, int n_images, float *out_gt, void *stream) {
    YB_NVTX("yb_gather_gt");
    YB_REQUIRE(table_dev && n_images > 0, "yb_gather_gt: bad arguments");
    YB_REQUIRE(out_gt != nullptr, "yb_gather_gt: out_gt is null");
    gather_gt_kernel<<<(n_images + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, n_images, out_gt);
    YB_LAUNCH_CHECK();
    return YB_OK;
}
