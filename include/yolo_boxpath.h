/* yolo_boxpath.h — C ABI of the B200-native box-geometry hot path.
 *
 * The reference (DarylFernandes99/custom-yolo-implmentation) is pure Python and has no plugin /
 * FFI surface; its boundary for this path is a handful of Python call
 * signatures.  Each entry point below names the reference function(s) it replaces; the Python
 * mirror of those signatures (custom-yolo-implmentation_b200/{model,utils,training}/) binds
 * them through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the host except
 *     the *_host entry points, which say so.
 *   - return 0 on success, negative yb_status on error; yb_last_error() gives the text
 *     (thread-local).  No exceptions cross the ABI, no hidden allocations: the caller owns
 *     every buffer, including the workspace whose size a *_workspace_bytes() query returns.
 *   - tensors use the reference's layouts: head output (N, 4*reg_max + nc, A) channel-major with
 *     the anchor axis contiguous (src/model/head.py:119); anchors (2, A); strides (1, A).
 *   - dtype: YB_F32 or YB_BF16 for the head output and its gradient; all arithmetic is fp32
 *     (the reference upcasts at src/model/losses.py:142).
 */
#ifndef YOLO_BOXPATH_H
#define YOLO_BOXPATH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_ABI_VERSION 6

typedef enum { YB_F32 = 0, YB_BF16 = 1 } yb_dtype;

typedef enum {
    YB_OK = 0,
    YB_ERR_ARG = -1,        /* bad argument (null pointer, non-positive size, unsupported value) */
    YB_ERR_WORKSPACE = -2,  /* workspace too small */
    YB_ERR_CUDA = -3,       /* a CUDA runtime call failed; see yb_last_error() */
    YB_ERR_ALIGN = -4       /* pointer not aligned as documented */
} yb_status;

int yb_abi_version(void);
const char *yb_last_error(void);
/* Kernels this library has launched in this process so far (bench.py reports the delta as gpu_launches). */
long long yb_launch_count(void);
/* sizeof() of this header's host structs as the library was built with them -- which: 0 yb_tal_params, 1 yb_tal_grid
 * (= yb_anchor_grid), 2 yb_peer_exchange, 3 yb_gt_source; 0 for anything else.  A foreign-language binding checks its own
 * struct layouts against these before the first call (tests/test_host_logic.py does for the ctypes mirror). */
size_t yb_struct_size(int which);

/* The anchors described as the reference's pyramid of regular grids (make_anchors, src/utils/model_utils.py:60-70: per
 * level x fastest, (x0 + col, y0 + row), one stride per level): level l holds anchors start[l] .. start[l] + w[l]*h[l].
 * A HINT, never a precondition: anchors and strides stay inputs of every entry point, the kernels either verify the
 * hint bit for bit (yb_tal_assign) or use it only where a wrong value costs speed, not correctness (yb_loss_fwd_bwd). */
#define YB_TAL_MAX_LEVELS 8
typedef struct {
    int n_levels;                               /* 0: no hint */
    int start[YB_TAL_MAX_LEVELS], w[YB_TAL_MAX_LEVELS], h[YB_TAL_MAX_LEVELS];
    float stride[YB_TAL_MAX_LEVELS], x0[YB_TAL_MAX_LEVELS], y0[YB_TAL_MAX_LEVELS];
} yb_tal_grid;                                  /* host struct, read during the call */
typedef yb_tal_grid yb_anchor_grid;

/* ------------------------------------------------------------------------------------------
 * Training: fused decode + nearest-centre assignment + DFL/QFL loss + backward.
 * Replaces YoloDFLQFLoss.forward and the autograd backward of it
 * (src/model/losses.py:93-281; invoked at src/training/train_model.py:245, :248/:252).
 *
 *   preds        (N, 4*reg_max + nc, A)  dtype            head output
 *   anchors      (2, A) fp32, strides (1, A) fp32         (src/model/head.py:112-114)
 *   gt           (gt_total, 5) fp32 [cx, cy, w, h, cls]   all images' boxes concatenated
 *   gt_offsets   (N + 1) int32                            image b owns rows [off[b], off[b+1])
 *   gmax         max boxes of any one image (host-known from the list shapes; sizes the per-image match work units:
 *                a wrong value costs time, never correctness; 0 = unknown)
 *   grad_preds   (N, C, A) dtype or NULL                  d total_loss / d preds (NULL: forward only)
 *   out_loss     8 floats: [0] total  [1] mean DFL ("box_loss")  [2] mean QFL ("cls_loss")
 *                          [3] number of distinct matched anchors  [4..5] reserved (zero)
 *                          [6] 1 if an in-kernel dependency of the launch timed out (then [0] is NaN; never seen: the
 *                              launch's consumers sit behind their producers in block-dispatch order, csrc/common.cuh)
 *                          [7] number of GT rows whose class id lies outside [0, nc): the reference raises on
 *                              those (scatter_, src/model/losses.py:260); the kernels clamp the id to stay
 *                              memory-safe and report the count, so the caller can raise without an extra sync
 *   out_idx      (gt_total) int32 or NULL                 matched anchor per GT  (losses.py:215)
 *   out_iou      (gt_total) fp32  or NULL                 IoU soft target per GT (losses.py:256)
 *   out_per_image (2, N) fp32 or NULL                     per-image DFL and QFL terms
 *   flags        0, or YB_LOSS_NO_PRUNE / YB_LOSS_SPLIT_LAUNCH / YB_LOSS_FORCE_PROBE / YB_LOSS_NO_PDL (test and profiling aids,
 *                results identical), YB_LOSS_WS_CLEAN (the caller keeps the workspace zeroed: no memset node, see below)
 *   grid_hint    NULL, or the anchors as a pyramid of grids: lets the launch bound every GT's nearest-centre distance
 *                up front (a few probe anchors per GT), so that the box role prunes from its first tile on; used when
 *                gmax > 128, where the coarse pyramid levels would otherwise scan several chunks of GTs.  Results
 *                never depend on it (tests/test_gpu_loss.py::test_tile_pruning_never_changes_the_result)
 *   stage_events NULL, or three cudaEvent_t handles of the CALLER (as void*), recorded on `stream` before the
 *                launch, after it, and (YB_LOSS_SPLIT_LAUNCH) after match_kernel (bench.py's roofline leg)
 *
 * The quirks of the reference that decide results are kept (Q1-Q6, Q16).
 * Images with no boxes contribute their QFL term and count in the mean (Q6).  The all-empty
 * batch, on which the reference raises, is the caller's to reject (gt_total == 0 is accepted
 * here and yields the QFL-only loss).
 * ---------------------------------------------------------------------------------------- */
size_t yb_loss_workspace_bytes(int n_images, int n_anchors, int gt_total, int dtype);

int yb_loss_fwd_bwd(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                    const float *anchors, const float *strides,
                    const float *gt, const int32_t *gt_offsets, int gt_total, int gmax,
                    float lambda_cls, float lambda_dfl,
                    void *grad_preds, float *out_loss, int32_t *out_idx, float *out_iou, float *out_per_image,
                    void *workspace, size_t workspace_bytes, unsigned flags, const yb_anchor_grid *grid_hint,
                    void *const *stage_events, void *stream);

#define YB_LOSS_NO_PRUNE 1u      /* box role scans every (GT, tile) pair: the exactness tests compare with and without */
#define YB_LOSS_FORCE_PROBE 4u   /* run the probe role (grid_hint) whatever gmax says: the exactness tests exercise it on small inputs */
#define YB_LOSS_SPLIT_LAUNCH 2u  /* box, class and match roles as three launches instead of one (profiling the roles apart) */
#define YB_LOSS_NO_PDL 16u       /* plain launch instead of a programmatic dependent launch */
#define YB_LOSS_WS_CLEAN 8u      /* the caller vouches that the whole workspace is zero: no memset node in front of the launch.
                                    A call made with this flag leaves the workspace all zero again when its launch completes
                                    (its last CTA wipes the words the launch used), so a buffer that was zeroed once and is
                                    only ever handed to yb_loss_fwd_bwd with this flag stays valid from step to step, whatever
                                    the batch shape.  Not after a call that reported a stalled dependency (out_loss[6]). */

/* The reference's GT argument -- a Python list of N (Mi, 5) fp32 device tensors (src/training/train_model.py:236, read at
 * src/model/losses.py:206-208) -- gathered into the (sum Mi, 5) buffer yb_loss_fwd_bwd / yb_tal_assign take, by ONE launch
 * driven by a device table the caller fills with one small copy (no per-image slicing or concatenation on the host).
 *   table_dev  n_images entries: source pointer, row pitch in floats (>= 5), first output row, number of rows */
typedef struct {
    const void *rows;       /* device pointer to the image's (Mi, >=5) fp32 rows; may be NULL when n_rows == 0 */
    int32_t row_pitch;      /* floats between consecutive rows */
    int32_t first_row;      /* where the image's rows start in out_gt (= gt_offsets[b]) */
    int32_t n_rows;         /* Mi */
    int32_t reserved;
} yb_gt_source;
int yb_gather_gt(const yb_gt_source *table_dev, int n_images, float *out_gt, void *stream);

/* grad *= *scale (device scalar), in place; returns without touching memory when *scale == 1.
 * Used by the autograd bridge for `loss.backward()` under a GradScaler
 * (src/training/train_model.py:247-253). */
int yb_scale_grad(void *grad, int dtype, size_t n_elements, const float *scale, void *stream);

/* Same as yb_loss_fwd_bwd but with HOST buffers (pinned for full speed): copies preds / gt to the
 * device buffers the caller provides, runs the fused path, copies out_loss (and grad, if
 * grad_host != NULL) back, and waits for the stream.  This is the end-to-end entry point that
 * bench.py's `e2e` leg times. */
int yb_loss_fwd_bwd_host(const void *preds_host, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                         const float *anchors, const float *strides,
                         const float *gt_host, const int32_t *gt_offsets_host, int gt_total, int gmax,
                         float lambda_cls, float lambda_dfl,
                         void *preds_dev, float *gt_dev, int32_t *gt_offsets_dev, void *grad_dev,
                         float *out_loss_dev, float *out_loss_host, void *grad_host,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Training, task-aligned variant (BASELINE.json north_star; NO counterpart in the reference — see
 * SURVEY.md §0.1 / §8(a'); specified by oracle/tal_oracle.py):  pairwise CIoU over anchors x GT,
 * metric = sigmoid(cls)^alpha * CIoU^beta, per-GT top-k (ties -> lowest anchor), conflicts to the larger
 * CIoU (ties -> lowest GT), normalised target scores; CIoU + DFL + BCE-with-logits loss and backward.
 * Two calls with the normaliser in between, so the caller can all-reduce it (SUM / world) under DDP:
 *
 *   yb_tal_assign  -> out_stats[0] = sum of target scores of this rank (un-clamped), [1] = #foreground,
 *                     [2] = 1 if a grid hint was given and rejected;
 *                     out_assigned_gt (N, A) int32 (-1 background) / out_target_score (N, A) fp32: optional
 *   yb_tal_loss    <- tss_dev: device scalar, the normaliser to use (clamped at 1 inside);
 *                  -> grad_preds (or NULL), out_loss: [0] total [1] box (CIoU) [2] cls (BCE) [3] dfl
 *                     [4] normaliser used [5] #foreground [6] grid hint rejected [7] #GT rows with a class id outside [0, nc)
 *                     (clamped to stay memory-safe; the module raises, as the reference's scatter_ would)
 * The same workspace must be passed to both calls (it carries the assignment).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int topk;                                   /* anchors per GT, 1..16 (public default 10) */
    float alpha, beta;                          /* metric = sigmoid(cls)^alpha * CIoU^beta (0.5, 6.0) */
    float lambda_box, lambda_cls, lambda_dfl;   /* loss weights */
    int vfl;                                    /* 0: plain BCE class term; 1: varifocal weighting (north_star "VFL-BCE"): */
    float vfl_alpha, vfl_gamma;                 /*    weight = vfl_alpha * sigmoid(x)^vfl_gamma on background cells
                                                      (differentiated), = the target score on the positive cell */
    unsigned flags;                             /* 0, or YB_TAL_WS_CLEAN: the caller vouches that the first 1152 + 4 * n_images (rounded up to 64) bytes of the workspace (the counters) are
                                                   zero -- true for a buffer that was zeroed once and has only seen complete
                                                   yb_tal_assign + yb_tal_loss pairs since (yb_tal_loss's last kernel wipes the
                                                   step's counters): no memset node in front of the step */
} yb_tal_params;                                /* host struct, read during the call; pass the SAME values to both calls */

/* Optional hint: the anchors as a pyramid of regular grids, what the reference's make_anchors produces
 * (src/utils/model_utils.py:60-70): level l owns anchors [start, start + w*h), x fastest, anchor (row, col) =
 * (x0 + col, y0 + row) in grid units, one stride per level.  The kernels VERIFY the hint against the anchor / stride
 * arrays on every call (bit for bit) and fall back to a structure-free scan when it does not hold — results never
 * depend on it, only the speed of the candidate enumeration does.  out_stats[2] = 1 reports a rejected hint. */
/* Optional peer exchange of the normaliser: instead of returning between the two calls for a collective, yb_tal_assign's
 * last kernel stores this rank's [sum of target scores, #foreground] into every rank's mailbox (peer memory over NVLink /
 * NVSwitch) and yb_tal_loss's first kernel waits on the own mailbox for all `world` entries of step `seq` and averages them
 * in rank order (every rank gets the bit-identical normaliser) — see csrc/peer.cu.  Replaces the one collective this path
 * has (the suggested `yolo_allreduce_small(ncclComm_t, ...)` of §8(b) — NCCL's ~35 us floor for an 8-byte
 * message is what this removes).  Set-up, once per process group, single node:
 *   yb_peer_mailbox_alloc -> own mailbox;  yb_peer_mailbox_export -> 64-byte CUDA IPC handle to send to the other ranks;
 *   yb_peer_mailbox_open(handle of rank r) -> mailbox[r] as mapped in this process;  mailbox[rank] = the own one.
 * `seq` must be the same on all ranks for one step, start at 1 and increase by 1 per step. */
#define YB_TAL_WS_CLEAN 1u
#define YB_PEER_MAX_WORLD 16
#define YB_PEER_HANDLE_BYTES 64
typedef struct {
    int world, rank;
    unsigned int seq;
    void *mailbox[YB_PEER_MAX_WORLD];
} yb_peer_exchange;                             /* host struct, read during the call */
size_t yb_peer_mailbox_bytes(void);
int yb_peer_mailbox_alloc(void **mailbox_out);
int yb_peer_mailbox_free(void *mailbox);
int yb_peer_mailbox_export(void *mailbox, void *handle_out_64);
int yb_peer_mailbox_open(const void *handle_64, void **peer_mailbox_out);
int yb_peer_mailbox_close(void *peer_mailbox);

size_t yb_tal_workspace_bytes(int n_images, int n_anchors, int gt_total, int dtype, int topk);

/* Decode, assignment and everything of the loss that does not need the normaliser (the foreground anchors' CIoU / DFL
 * terms and box-logit gradients, kept un-normalised in the workspace). */
int yb_tal_assign(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                  const float *anchors, const float *strides, const float *gt, const int32_t *gt_offsets,
                  int gt_total, const yb_tal_params *params, const yb_tal_grid *grid_hint /* or NULL */,
                  const yb_peer_exchange *peers /* or NULL */,
                  float *out_stats, int32_t *out_assigned_gt, float *out_target_score,
                  void *workspace, size_t workspace_bytes, void *stream);

/* The dense class pass (reads the class logits, writes the whole gradient) and the loss scalars. */
int yb_tal_loss(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors, int gt_total,
                const yb_tal_params *params, const float *tss_dev /* ignored when peers != NULL */,
                const yb_peer_exchange *peers /* or NULL: the struct given to yb_tal_assign */,
                void *grad_preds, float *out_loss, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Decode.  Replaces DFL.forward (src/model/model_blocks.py:278-280), dist2bbox
 * (src/utils/model_utils.py:120-129) and the decode block of decode_predictions
 * (src/training/train_model.py:36-109) / Model.inference (src/model/model_builder.py:123-133).
 *
 *   box_logits  (N, 4*reg_max, A) rows of a tensor whose image stride is `image_stride` elements
 *   out_ltrb    (N, 4, A) fp32 or NULL    expected bin per side, grid units
 *   out_box     (N, 4, A) fp32 or NULL    xywh (box_format 0) or xyxy (1), times stride if scale_by_stride
 * ---------------------------------------------------------------------------------------- */
int yb_dfl_decode(const void *box_logits, int dtype, int n_images, int reg_max, int n_anchors, size_t image_stride,
                  const float *anchors, const float *strides, float *out_ltrb, float *out_box,
                  int box_format, int scale_by_stride, void *stream);

/* make_anchors (src/utils/model_utils.py:18-70; called at src/model/head.py:94, :112): integer cell
 * indices, x fastest within a level, levels concatenated, plus the per-anchor stride.
 *   shapes_host  n_levels x (h, w) int32 on the HOST     strides_host  n_levels floats on the HOST
 *   out_grid (A, 2) fp32 = (col, row) WITHOUT the offset (the caller adds it in the target dtype so
 *   that the reference's rounding sequence is kept)       out_strides (A, 1) fp32 */
int yb_make_anchors(const int32_t *shapes_host, const float *strides_host, int n_levels, float *out_grid,
                    float *out_strides, void *stream);

/* Tail of Head.forward (src/model/head.py:86-121): `cat((box_l, cls_l), 1)` per level, flatten H x W,
 * `cat(..., 2)` over the levels -- one launch, each element moved once (the reference moves it twice).
 *   box_levels / cls_levels   HOST arrays of n_levels DEVICE pointers: (N, box_ch, H_l*W_l) / (N, nc, H_l*W_l),
 *                             contiguous, dtype YB_F32 or YB_BF16 (the outputs of the last 1x1 convs, head.py:52, :60)
 *   hw_host                   n_levels int32 on the HOST: H_l * W_l          (n_levels <= 8)
 *   out                       (N, box_ch + nc, sum H_l*W_l)
 * yb_head_scatter is the adjoint (the backward of the two cats): grad (N, C, A) -> the 2 * n_levels
 * contiguous conv-output gradients. */
int yb_head_gather(const void *const *box_levels, const void *const *cls_levels, const int32_t *hw_host, int n_levels,
                   int dtype, int n_images, int box_ch, int nc, void *out, void *stream);
int yb_head_scatter(const void *grad, const int32_t *hw_host, int n_levels, int dtype, int n_images, int box_ch, int nc,
                    void *const *box_grads, void *const *cls_grads, void *stream);

/* ltrb (N, 4, A) fp32 + anchors (2, A) -> box (N, 4, A); dist2bbox with dim=1. */
int yb_dist2bbox(const float *ltrb, const float *anchors, int n_images, int n_anchors, int xywh, float *out_box,
                 void *stream);

/* Validation decode: decode + sigmoid + best class + `>= conf` + top-k by score, per image.
 * Replaces decode_predictions (src/training/train_model.py:14-142).
 *   out_rows  (N, top_k, 5) fp32 [cx, cy, w, h, cls]   out_count (N) int32   out_anchor (N, top_k) int32 or NULL
 * Row order: anchor order when at most top_k candidates pass, else score descending (ties: lowest anchor). */
size_t yb_val_decode_workspace_bytes(int n_images, int n_anchors);
int yb_val_decode(const void *preds, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                  const float *anchors, const float *strides, float conf_thres, int top_k,
                  float *out_rows, int32_t *out_count, int32_t *out_anchor,
                  void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Eval: batched class-aware NMS.  Replaces non_max_suppression (src/utils/model_utils.py:174-279)
 * including the torchvision.ops.nms call inside it (:264), for multi_label=False, labels=().
 *
 *   prediction (N, 4 + nc, A) fp32: xywh pixels + per-class scores
 *   iou_thres is a double so that "IoU > thr" is decided exactly as torchvision's CPU kernel does
 *   class_filter: NULL or n_class_filter int32 class ids to keep (the `classes` argument)
 *   out_rows  (N, max_det, 6) fp32 [x1, y1, x2, y2, conf, cls]  score descending
 *   out_count (N) int32        out_anchor (N, max_det) int32 or NULL (anchor index of each row)
 * The reference's wall-clock abort (:212, :275-277) is not reproduced.
 * ---------------------------------------------------------------------------------------- */
size_t yb_nms_workspace_bytes(int n_images, int n_anchors);
int yb_nms(const float *prediction, int n_images, int nc, int n_anchors,
           float conf_thres, double iou_thres, int max_det, int agnostic,
           const int32_t *class_filter, int n_class_filter,
           float *out_rows, int32_t *out_count, int32_t *out_anchor,
           void *workspace, size_t workspace_bytes, void *stream);

/* Model.inference after the network (src/model/model_builder.py:123-139: split, DFL, dist2bbox, "* strides", cat, NMS)
 * as ONE entry point.  The box channels of the head output are decoded into a compact (N, 4, A) fp32 buffer inside the
 * workspace; the NMS kernels take their scores straight from the head output's class channels -- raw logits, as the
 * reference feeds them, or through a sigmoid when apply_sigmoid != 0 -- so the (N, 4 + nc, A) tensor the
 * reference concatenates (:136) never exists.  Everything else as yb_nms.
 *   head_out  (N, 4*reg_max + nc, A) fp32 or bf16      anchors (2, A), strides (1, A) fp32 */
size_t yb_postprocess_workspace_bytes(int n_images, int n_anchors);
int yb_postprocess(const void *head_out, int dtype, int n_images, int nc, int reg_max, int n_anchors,
                   const float *anchors, const float *strides, int apply_sigmoid, float conf_thres, double iou_thres,
                   int max_det, int agnostic, const int32_t *class_filter, int n_class_filter,
                   float *out_rows, int32_t *out_count, int32_t *out_anchor,
                   void *workspace, size_t workspace_bytes, void *stream);

/* multi_label=True (src/utils/model_utils.py:240-242): every (anchor, class) pair with score > conf_thres is
 * a candidate; the max_nms = 30000 best of an image (:211, :259; score descending, ties -> lower
 * anchor * nc + class, the order of the reference's nonzero()) go through the same greedy NMS.
 * Same arguments and outputs as yb_nms; its own workspace. */
size_t yb_nms_multilabel_workspace_bytes(int n_images);
int yb_nms_multilabel(const float *prediction, int n_images, int nc, int n_anchors, float conf_thres, double iou_thres,
                      int max_det, int agnostic, const int32_t *class_filter, int n_class_filter,
                      float *out_rows, int32_t *out_count, int32_t *out_anchor,
                      void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * IoU / box utilities (fp32).
 *   yb_xywh2xyxy      src/utils/model_utils.py:153-172
 *   yb_bbox_iou       src/model/losses.py:9-40   element-wise, xywh, eps 1e-6, b1_y2 slip kept;
 *                     grad_box1 (M,4) or NULL receives d(sum_i go[i]*iou[i]) / d box1
 *   yb_box_iou        src/utils/model_utils.py:131-151  pairwise xyxy, eps argument
 *   yb_box_iou_batch  src/training/metrics.py:6-41      pairwise xywh, eps 1e-6
 * ---------------------------------------------------------------------------------------- */
int yb_xywh2xyxy(const float *in, size_t n_boxes, float *out, void *stream);
int yb_bbox_iou(const float *box1, const float *box2, int m, float *out_iou,
                const float *grad_out, float *grad_box1, void *stream);
int yb_box_iou(const float *box1, int n, const float *box2, int m, float eps, float *out, void *stream);
int yb_box_iou_batch(const float *box1, int n, const float *box2, int m, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Validation bookkeeping.  Replaces DetectionMetrics.update (src/training/metrics.py:68-160; called per
 * image at src/training/train_model.py:326-328) for a whole batch: greedy class-matched TP/FP/FN.
 *   pred_rows   (N, row_stride, 5) fp32 [cx, cy, w, h, cls] (the layout yb_val_decode writes), pred_count (N)
 *   pred_scores (N, row_stride) or NULL; rows with score < score_threshold are ignored (:82-86)
 *   gt / gt_offsets as in yb_loss_fwd_bwd; gmax = most targets of one image (<= 1024)
 *   skip_empty_targets != 0: images without targets add nothing, as in the reference's validation loop, which calls
 *               update() only `if gt_box[i].numel() > 0` (train_model.py:327); 0 = update()'s own rule (their
 *               predictions count as false positives, metrics.py:98-104)
 *   counters    uint64[8 + 4*nc], ACCUMULATED (zero them to reset):
 *               [0] tp [1] fp [2] fn [3] total_predictions [4] total_ground_truths
 *               [8..] class_tp, then class_fp, class_fn, class_gt_count (nc each)
 * ---------------------------------------------------------------------------------------- */
size_t yb_detection_counters_bytes(int nc);
int yb_detection_match(const float *pred_rows, int row_stride, const int32_t *pred_count, const float *pred_scores,
                       float score_threshold, const float *gt, const int32_t *gt_offsets, int gmax, int n_images,
                       int nc, float iou_threshold, int skip_empty_targets, uint64_t *counters, void *stream);

/* quality_focal_loss (src/model/losses.py:46-57) on dense (M, C) logits/targets, fp32:
 * out_loss[0] = -sum(...)/M; grad_scores (M, C) or NULL = d out_loss / d pred_scores (times *grad_out if given). */
size_t yb_qfl_workspace_bytes(size_t n_elements);
int yb_quality_focal_loss(const float *pred_scores, const float *target_scores, int m, int c, float beta,
                          float *out_loss, float *grad_scores, void *workspace, size_t workspace_bytes, void *stream);

/* distribution_focal_loss (src/model/losses.py:63-78): pred_dist (M, R) logits, target (M,) in [0, R-1):
 * out_loss[0] = mean over M; grad_dist (M, R) or NULL = d out_loss / d pred_dist. */
int yb_distribution_focal_loss(const float *pred_dist, const float *target, int m, int r,
                               float *out_loss, float *grad_dist, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLO_BOXPATH_H */
