"""B200-native box-geometry hot path of the custom YOLOv11-style detector.

Sub-modules mirror the reference's module paths for this path:

  model.losses            YoloDFLQFLoss, bbox_iou, quality_focal_loss, distribution_focal_loss
  model.model_blocks      DFL
  utils.model_utils       make_anchors, dist2bbox, box_iou, xywh2xyxy, non_max_suppression
  training.train_model    decode_predictions
  training.metrics        box_iou_batch
  training.distributed_setup   reduce_value / reduce_loss_stats (one fused small all-reduce)

All arithmetic runs in hand-written sm_100a CUDA behind the C-ABI declared in
``include/yolo_boxpath.h`` (``csrc/``, loaded by ``_cabi``).  There is no CPU fallback: calling
an op without the built library or with a non-CUDA tensor raises.
"""
__version__ = "0.1.0"
