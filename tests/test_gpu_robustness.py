"""Inputs a trained or diverging network really produces, and shapes at the edges of the launch geometry:
extreme logits, degenerate GT boxes, tied / degenerate NMS boxes, GTs no anchor falls into, many small images.
Every case is checked against the CPU oracle (same tolerances as the parity tests)."""
import pytest
import torch

from oracle import loss_oracle as L
from oracle import nms_oracle as N
from oracle import tal_oracle as T
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200.utils import model_utils as U
from custom_yolo_implmentation_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


def _loss_vs_oracle(preds, gts, anchors, strides, nc, dev, rtol=1e-5):
    x = preds.to(dev).requires_grad_(True)
    crit = P.YoloDFLQFLoss(num_classes=nc)
    loss, parts = crit(x, [g.to(dev) for g in gts], anchors.to(dev), strides.to(dev))
    loss.backward()
    ora = L.loss_forward_backward(preds, gts, anchors, strides, nc)
    assert torch.isfinite(loss).item() and torch.isfinite(x.grad).all().item()
    assert abs(loss.item() - ora.total.item()) <= rtol * abs(ora.total.item()) + 1e-7
    scale = ora.grad.abs().max().item()
    assert (x.grad.cpu() - ora.grad).abs().max().item() <= rtol * scale
    return ora


def test_loss_with_saturated_logits(cuda_device):
    """Class logits down to -30 / up to +15 and one-hot-sharp DFL rows: no overflow, same numbers as the oracle."""
    preds, gts, anchors, strides = syn.make_loss_inputs(2, 7, 160, 8, 501)
    g = torch.Generator().manual_seed(502)
    preds[:, 64:] = torch.randn(preds[:, 64:].shape, generator=g) * 9.0 - 8.0
    preds[:, :64] = preds[:, :64] * 12.0                                    # near one-hot softmax over the 16 bins
    preds[0, 64:, :7] = torch.tensor([-30.0, -20.0, -12.0, 0.0, 6.0, 12.0, 15.0])
    _loss_vs_oracle(preds, gts, anchors, strides, 7, cuda_device)


def test_loss_with_degenerate_gt_boxes(cuda_device):
    """Zero-size, huge, out-of-image and duplicated GT boxes, a class id given as a float with a fraction."""
    preds, gts, anchors, strides = syn.make_loss_inputs(3, 5, 128, 4, 511)
    gts[0] = torch.tensor([[64.0, 64.0, 0.0, 0.0, 1.0],                     # zero area
                           [64.0, 64.0, 500.0, 400.0, 2.0],                 # larger than the image
                           [-40.0, 300.0, 20.0, 10.0, 0.0],                 # centre outside the image
                           [64.0, 64.0, 0.0, 0.0, 1.0]])                    # exact duplicate of the first
    gts[1] = torch.tensor([[10.5, 11.5, 3.0, 2.0, 3.9]])                    # .long() truncates the class to 3
    gts[2] = torch.zeros(0, 5)
    ora = _loss_vs_oracle(preds, gts, anchors, strides, 5, cuda_device)
    assert ora.idx[0][0] == ora.idx[0][3]                                   # the duplicates share their anchor


def test_loss_many_small_images_and_tiny_grids(cuda_device):
    """300 images (more than a wave of CTAs per tile index) on a 32 px input: A = 21 anchors, one ragged tile."""
    preds, gts, anchors, strides = syn.make_loss_inputs(300, 3, 32, 3, 521)
    assert preds.shape[2] == 21
    _loss_vs_oracle(preds, gts, anchors, strides, 3, cuda_device)


def _nms_vs_oracle(x, conf, iou, dev, nc, **kw):
    rows, count, anchor = U.batched_nms_raw(x.to(dev), conf, iou, 300, nc, kw.get("agnostic", False), None, want_anchor=True)
    ora = N.nms_forward(x, conf, iou, agnostic=kw.get("agnostic", False), max_det=300, nc=nc)
    for b in range(x.shape[0]):
        k = int(count[b])
        assert k == ora.rows[b].shape[0]
        assert torch.equal(anchor[b, :k].cpu().long(), ora.keep_anchor[b])
        assert torch.equal(rows[b, :k].cpu(), ora.rows[b])


@pytest.mark.parametrize("agnostic", [False, True])
def test_nms_ties_and_degenerate_boxes(agnostic, cuda_device):
    """Equal scores (ties go to the lower anchor), identical boxes, zero-area and inverted (negative w / h) boxes."""
    x = syn.make_nms_input(2, 4, 160, 531)
    a = x.shape[2]
    x[0, 4:, : a // 2] = x[0, 4:, a // 2: 2 * (a // 2)]                     # pairs of anchors with equal scores
    x[0, :4, 10:20] = x[0, :4, 30:40]                                       # identical boxes
    x[0, 2, 40:60] = 0.0                                                    # zero width
    x[1, 2, 5:25] = -x[1, 2, 5:25]                                          # negative width
    x[1, 3, 15:35] = -x[1, 3, 15:35]                                        # negative height (some both)
    x[1, 4:, 100:140] = 0.5                                                 # a plateau of equal scores across classes
    _nms_vs_oracle(x, 0.001, 0.5, cuda_device, 4, agnostic=agnostic)
    _nms_vs_oracle(x, 0.001, 0.95, cuda_device, 4, agnostic=agnostic)


def test_nms_boxes_wider_than_the_class_offset(cuda_device):
    """Candidates spanning more than 7680 px break the class-parallel decomposition: the kernel must notice
    and take the generic path (classes then DO interact through the offset, exactly as in the reference)."""
    x = syn.make_nms_input(2, 3, 160, 541)
    x[0, 0, :6] = torch.tensor([0.0, 7700.0, 15400.0, 20.0, 7690.0, 40.0])  # centres far apart
    x[0, 2, :6] = torch.tensor([60.0, 80.0, 60.0, 7800.0, 30.0, 16000.0])   # and very wide boxes
    _nms_vs_oracle(x, 0.001, 0.3, cuda_device, 3)


def test_tal_gt_without_inside_anchor_and_out_of_range_class(cuda_device):
    """A GT so small that no anchor centre lies inside it gets no foreground anchor; a class id beyond nc is clamped
    by the kernel (the oracle indexes it, so it is kept in range there) and images without GT are skipped."""
    preds, gts, anchors, strides = syn.make_loss_inputs(3, 6, 160, 5, 551)
    gts[0] = torch.cat((gts[0], torch.tensor([[41.0, 41.0, 1.0, 1.0, 2.0]])))        # between the stride-8 centres 36 and 44
    gts[1] = torch.zeros(0, 5)
    dev = cuda_device
    gt, off, counts = P.pack_gt([g.to(dev) for g in gts], dev)
    out, grad, tr = P.fused_tal_loss(preds.to(dev), gt, off, anchors.to(dev), strides.to(dev), 6, 1.5, 1.0, 1.5, want_trace=True)
    ora = T.tal_forward_backward(preds, gts, anchors, strides, 6)
    asg = tr["assigned_gt"].cpu().long()
    assert asg.equal(ora.assigned_gt)
    assert not (asg[0] == gts[0].shape[0] - 1).any()                                  # the tiny GT owns nothing
    assert (asg[1] == -1).all()
    assert abs(out[0].item() - ora.total.item()) <= 1e-5 * abs(ora.total.item())
    assert (grad.cpu() - ora.grad).abs().max().item() <= 1e-5 * ora.grad.abs().max().item()


def test_val_decode_class_ties_in_sigmoid_space(cuda_device):
    """decode_predictions takes the argmax over float SIGMOIDS (train_model.py:116-119): two different logits can
    round to the same sigmoid, and then the EARLIER class wins even if its logit is the smaller one.  The kernel
    tracks logits and must fall back to the reference's rule exactly in that case."""
    from custom_yolo_implmentation_b200.training.train_model import decode_predictions_raw
    from oracle import decode_oracle as D
    anchors, strides = syn.anchor_grid(160)
    a = anchors.shape[1]
    preds = syn.make_preds(2, 6, a, 601, cls_mean=-2.0, cls_std=1.0)
    cls = preds[:, 64:]
    cls[0, :, 0:40] = torch.tensor([17.0, 18.0, 20.0, -3.0, 19.0, 30.0])[:, None]      # all saturate to 1.0: class 0 wins
    cls[0, :, 40:80] = torch.tensor([-3.0, 9.99995, 10.0, 9.9999, -1.0, 0.0])[:, None]  # near-ties around 10
    cls[1, :, 0:40] = torch.tensor([-120.0, -110.0, -105.0, -130.0, -104.0, -200.0])[:, None]   # all underflow to 0
    cls[1, 2, 40:80] = 5.0
    cls[1, 4, 40:80] = 5.0                                                              # exactly equal logits: first wins
    rows, count, anchor = decode_predictions_raw(preds.to(cuda_device), anchors.to(cuda_device), strides.to(cuda_device),
                                                 0.0, a, 6, want_anchor=True)
    ora = D.val_decode(preds, anchors, strides, 0.0, a, 6)
    for b in range(2):
        k = int(count[b])
        assert k == a == ora.rows[b].shape[0]                                           # conf 0: every anchor is kept, anchor order
        assert torch.equal(anchor[b, :k].cpu().long(), ora.anchor[b])
        assert torch.equal(rows[b, :k, 4].cpu(), ora.rows[b][:, 4])                     # class ids, bit-exact
