"""World-size-2 run of the CUDA path on ONE GPU (both ranks on cuda:0, gloo carrying the CUDA tensors —
NCCL refuses two ranks on one device): the batch shards, the nearest-centre path exchanges nothing but
the loss statistics, and the task-aligned path all-reduces its normaliser between its two ABI calls — once with a
collective, once through peer-mapped mailboxes (CUDA IPC works between two processes on one device too)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from custom_yolo_implmentation_b200.model import losses as P
    from custom_yolo_implmentation_b200.training import distributed_setup as DS
    from test_gpu_tal import make_inputs

    dev = torch.device("cuda:0")
    n_global, nc = 6, 20
    preds, gts, anchors, strides = make_inputs(n_global, nc, 320, 30, 91)
    lo, hi = DS.shard_batch(n_global, rank, world)
    x = preds[lo:hi].to(dev).requires_grad_(True)
    g = [t.to(dev) for t in gts[lo:hi]]
    a, s = anchors.to(dev), strides.to(dev)

    # nearest-centre path: per-image means, reduced with ONE all-reduce
    crit = P.YoloDFLQFLoss(num_classes=nc)
    loss, _ = crit(x, g, a, s)
    loss.backward()
    red = DS.reduce_loss_stats(crit.last_stats, hi - lo).cpu()
    grad_a = x.grad.detach().cpu().clone()
    x.grad = None

    # task-aligned path: the normaliser is the global sum of target scores / world
    tal = P.YoloDFLQFLoss(num_classes=nc, assigner="tal")
    loss_t, d = tal(x, g, a, s)
    loss_t.backward()
    res = {"red": red, "grad_a": grad_a, "tal_total": loss_t.item(), "tal_dict": d, "tal_norm": tal.last_stats[4].item(),
           "tal_grad": x.grad.detach().cpu(), "lo": lo, "hi": hi}

    # the same exchange through peer-mapped mailboxes (csrc/peer.cu): the assign call's last kernel stores into both
    # ranks' mailboxes, the loss call's first kernel polls the own one.  Seven steps in a row: the ring of 4 slots wraps.
    px = DS.enable_peer_exchange()
    assert DS.default_peer_exchange() is px and px.world == 2
    peer = []
    for step in range(7):
        x.grad = None
        loss_p, _ = tal(x, g, a, s)
        loss_p.backward()
        peer.append((loss_p.item(), tal.last_stats[4].item(), tal.last_stats[5].item()))
    res.update(peer=peer, peer_grad=x.grad.detach().cpu())
    torch.cuda.synchronize()
    dist.barrier()
    px.close()
    torch.save(res, out.format(rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_shard_the_batch_and_exchange_only_statistics(tmp_path, cuda_device):
    from oracle import tal_oracle as T
    from custom_yolo_implmentation_b200.model import losses as P
    from test_gpu_tal import make_inputs
    out = str(tmp_path / "r{}.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = [torch.load(out.format(k), weights_only=False) for k in range(2)]
    n_global, nc = 6, 20
    preds, gts, anchors, strides = make_inputs(n_global, nc, 320, 30, 91)
    dev = cuda_device

    # nearest-centre: the reduced statistics equal the unsharded batch; gradients are per-sample (scaled by N_local)
    crit = P.YoloDFLQFLoss(num_classes=nc)
    x = preds.to(dev).requires_grad_(True)
    loss, _ = crit(x, [g.to(dev) for g in gts], anchors.to(dev), strides.to(dev))
    loss.backward()
    full = crit.last_stats.cpu()
    assert torch.equal(r[0]["red"], r[1]["red"])
    for k in range(3):
        assert abs(r[0]["red"][k].item() - full[k].item()) <= 2e-6 * abs(full[k].item())
    assert r[0]["red"][4].item() == float(n_global)
    both = torch.cat((r[0]["grad_a"], r[1]["grad_a"]))              # each rank divides by its own 3 images, the full batch by 6
    assert torch.allclose(both * (3.0 / n_global), x.grad.cpu(), rtol=1e-5, atol=1e-9)

    # task-aligned: both ranks used the same normaliser = global sum / world; each shard's loss is the oracle's with it
    whole = T.tal_forward(preds, gts, anchors, strides, nc)
    norm = max(whole.tss / 2.0, 1.0)
    assert abs(r[0]["tal_norm"] - norm) <= 1e-5 * norm and r[0]["tal_norm"] == r[1]["tal_norm"]
    for k in range(2):
        lo, hi = r[k]["lo"], r[k]["hi"]
        ora = T.tal_forward_backward(preds[lo:hi], gts[lo:hi], anchors, strides, nc, tss_override=whole.tss / 2.0)
        assert abs(r[k]["tal_total"] - ora.total.item()) <= 1e-5 * abs(ora.total.item())
        assert (r[k]["tal_grad"] - ora.grad).abs().max().item() <= 1e-5 * ora.grad.abs().max().item()
        # peer mailboxes: every step gives what the collective gave (normaliser equal up to the summation order)
        for total, norm_p, nfg in r[k]["peer"]:
            assert abs(norm_p - norm) <= 1e-6 * norm and abs(total - r[k]["tal_total"]) <= 2e-6 * abs(total)
        assert (r[k]["peer_grad"] - r[k]["tal_grad"]).abs().max().item() <= 2e-6 * r[k]["tal_grad"].abs().max().item()
    # ... and both ranks used the BIT-identical normaliser in every step (out_loss[5], the foreground count, is the rank's own)
    assert [p[1] for p in r[0]["peer"]] == [p[1] for p in r[1]["peer"]]
