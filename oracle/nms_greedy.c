/* CPU oracle: greedy hard-NMS over score-sorted boxes.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates the published algorithm of torchvision.ops.nms (third-party; not under
 * /root/reference; pinned 0.24.1 by environment.yml:29; called at
 * src/utils/model_utils.py:264).  Boxes arrive already sorted by descending score, xyxy, fp32,
 * with the reference's class offset already added (model_utils.py:262-263).
 *
 *   keep box i unless an earlier kept box k has  inter/(area_k + area_i - inter) > thr
 *
 * with every quantity in fp32, width/height clamped at 0, NO epsilon in the denominator, and the
 * comparison made against the double threshold (a float IoU promoted to double), exactly as the
 * CPU kernel of torchvision does.  Stops after max_det keeps (model_utils.py:265 `i[:max_det]`
 * — truncation of a greedy prefix equals early exit).
 */
#include <stdlib.h>

int nms_greedy_sorted(const float *boxes, int n, double thr, int max_det, int *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int kept = 0;
    for (int i = 0; i < n; ++i)
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    for (int i = 0; i < n && kept < max_det; ++i) {
        if (dead[i]) continue;
        keep[kept++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float ia = area[i];
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1, h = yy2 - yy1;
            if (w < 0.f) w = 0.f;
            if (h < 0.f) h = 0.f;
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    free(dead);
    free(area);
    return kept;
}
