"""World-size-2 checks of the N>1 path on the CPU (gloo): the batch shards with no data-path
collective; the only exchange is one small all-reduce of the loss statistics."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from custom_yolo_implmentation_b200.training import distributed_setup as DS
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    from oracle import loss_oracle as L

    n_global, nc = 4, 5
    preds, gts, anchors, strides = syn.make_loss_inputs(n_global, nc, 96, 6, 77)
    lo, hi = DS.shard_batch(n_global, rank, world)
    tr = L.loss_forward(preds[lo:hi], gts[lo:hi], anchors, strides, nc)          # the rank's shard (oracle on CPU)
    fg = float(sum(len(set(i.tolist())) for i in tr.idx))
    stats = torch.tensor([tr.total.item(), tr.dfl_mean.item(), tr.cls_mean.item(), fg, 0, 0, 0, 0])
    red = DS.reduce_loss_stats(stats, hi - lo)
    mean3 = [DS.reduce_value(v, average=True) for v in (tr.total.item(), tr.dfl_mean.item(), tr.cls_mean.item())]
    tbuf = torch.tensor([float(rank + 1)])
    tsum = DS.reduce_value(tbuf, average=False)           # a Python float, and the tensor is reduced in place (reference :58-63)
    assert isinstance(tsum, float) and tbuf.item() == tsum
    if rank == 0:
        full = L.loss_forward(preds, gts, anchors, strides, nc)
        torch.save({"red": red, "mean3": mean3, "tsum": tsum, "full": [full.total.item(), full.dfl_mean.item(), full.cls_mean.item()],
                    "fg_full": float(sum(len(set(i.tolist())) for i in full.idx))}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_loss_stats_match_the_unsharded_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    red, full = r["red"], r["full"]
    # per-image means over the global batch == the loss of the unsharded batch (every image is independent)
    for k in range(3):
        assert abs(red[k].item() - full[k]) <= 2e-6 * abs(full[k])
        assert abs(r["mean3"][k] - full[k]) <= 2e-6 * abs(full[k])       # the reference's three reduce_value calls
    assert red[3].item() == r["fg_full"] and red[4].item() == 4.0
    assert r["tsum"] == 3.0


def test_reduce_helpers_are_identity_without_a_process_group():
    from custom_yolo_implmentation_b200.training import distributed_setup as DS
    s = torch.tensor([1.0, 2.0, 3.0, 4.0, 0, 0, 0, 0])
    out = DS.reduce_loss_stats(s, 8)
    assert torch.allclose(out[:4], s[:4]) and out[4].item() == 8.0
    assert DS.shard_batch(10, 1, 3) == (3, 6)
