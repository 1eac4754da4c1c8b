#!/usr/bin/env python
"""Benchmark of the box-geometry hot path (BASELINE.json metric: images/s through
decode+assign+loss(+bwd) and through NMS; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the fused decode + assign + DFL/QFL loss + backward over one batch of
synthetic head outputs.  Workload at every N: cfg2 of BASELINE.json per GPU (batch 128, 640x640,
8400 anchors, 80 classes, <=100 GT/img, fp32) — i.e. cfg3 (global batch 1024) at N=8: weak scaling.

  value       whole-job images/s with inputs resident in HBM, timed with CUDA events on the launching
              stream, max over ranks (inputs are 619 MB/step: larger than the 126 MB L2).
  e2e         the same metric through the public drop-in API (YoloDFLQFLoss.forward + backward) with
              HOST inputs: every step copies preds / GT from pinned host memory and reads the loss
              scalars back.
  roofline    dominant kernel (fused_main_kernel): algorithmic bytes / CUDA-event duration vs the
              measured copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference
              the CPU oracle port of the reference's loss (oracle/loss_oracle.py; the reference is
              Python and cannot travel to the box) on all host cores, on a bounded sample.
  nms         cfg4 (batch 64, 8400 candidates, conf 0.001, IoU 0.7, max_det 300) through yb_nms.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(batch_per_gpu=128, imgsz=640, nc=80, gmax=100, reg_max=16)
CPU_SAMPLE_IMAGES = 16
METRIC = "images/sec through decode+assign+loss(+bwd)"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class gpu_local_cpus:
    """Context manager: run on the CPUs NVML names as local to GPU `index` (first-touch places pinned host memory on
    that NUMA node; a staging buffer on the far socket halves the host-to-device rate).  Restores the affinity."""

    def __init__(self, index):
        self.index, self.old, self.note = index, None, "unchanged"

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.old = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self.note = f"{len(os.sched_getaffinity(0))} of {len(self.old)} cpus"
        except Exception as e:                              # not permitted / not supported: leave the placement alone
            self.note = f"unavailable ({type(e).__name__})"
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except OSError:
                pass
        return False


class ClockSampler:
    """SM clock / throttle reasons of one GPU through NVML, in-process.

    NVML queries take the driver lock for milliseconds and stall kernel launches, so nothing is
    sampled while the timed loop is ENQUEUEING; the samples are taken right after the last launch of
    the timed region, while the GPU is still draining the queued steps (i.e. under load, inside the
    timed region, before the closing synchronize)."""

    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.rows, self.h, self.mx, self.sample_ms = [], None, None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.bits = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                         pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]
        except Exception:
            self.h = None

    def sample(self, n=1, gap=0.002):
        if self.h is None:
            return
        for i in range(n):
            t0 = time.perf_counter()
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), [nm for nm, b in zip(self.NAMES, self.bits) if r & b]))
            except Exception:
                pass
            self.sample_ms.append((time.perf_counter() - t0) * 1e3)
            if i + 1 < n:
                time.sleep(gap)

    def summary(self):
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.mx) if self.mx else None,
                "reasons": sorted({n for r in self.rows for n in r[1]}), "samples": len(sm),
                "nvml_ms_per_sample": round(statistics.mean(self.sample_ms), 3) if self.sample_ms else None}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_loss_images_per_s(steps, warmup, seed=1236):
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    from oracle import loss_oracle

    torch.set_num_threads(os.cpu_count() or 1)
    n = CPU_SAMPLE_IMAGES
    preds, gts, anchors, strides = syn.make_loss_inputs(n, CFG["nc"], CFG["imgsz"], CFG["gmax"], seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss_oracle.loss_forward_backward(preds, gts, anchors, strides, CFG["nc"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return n / med, med, torch.get_num_threads()


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    ips, med, threads = cpu_loss_images_per_s(steps, warmup)
    sample = (f"{CPU_SAMPLE_IMAGES} images of the cfg2 workload per step (640x640, 8400 anchors, nc=80, <=100 GT/img, fp32), "
              f"oracle/loss_oracle.py fwd+bwd on torch CPU, median of {steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: batch 128/GPU, 640x640 (8400 anchors, reg_max 16), 80 classes, <=100 GT/img, "
                                   "fused decode+assign+loss+backward", "cpu_sample_images": CPU_SAMPLE_IMAGES},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from custom_yolo_implmentation_b200 import _cabi
    from custom_yolo_implmentation_b200.model.losses import YoloDFLQFLoss, fused_loss, fused_tal_loss, pack_gt, pack_gt_host
    from custom_yolo_implmentation_b200.training.distributed_setup import reduce_loss_stats
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    from custom_yolo_implmentation_b200.utils.model_utils import batched_nms_raw

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    n, nc, imgsz, gmax = CFG["batch_per_gpu"], CFG["nc"], CFG["imgsz"], CFG["gmax"]
    preds_h, gts_h, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, 1236 + rank)
    a = preds_h.shape[2]
    preds = preds_h.to(dev)
    anchors_d, strides_d = anchors.to(dev), strides.to(dev)
    gt, off, counts = pack_gt([g.to(dev) for g in gts_h], dev)
    gmax_real = max(counts)
    bytes_per_step = 2 * preds.numel() * preds.element_size()          # read once + gradient written once

    def step():
        out, grad, _ = fused_loss(preds, gt, off, gmax_real, anchors_d, strides_d, nc, 1.0, 1.5, want_grad=True)
        return out, grad

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):             # results bound exactly as in the timed loop, so the caching
        out, grad = step()                           # allocator already owns both 619 MB gradient blocks
    if world > 1:
        reduce_loss_stats(out, n)                    # NCCL connects its channels lazily on the first collective
    barrier()
    launches0 = _cabi.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    if True:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out, grad = step()
        if world > 1:
            # the path has no per-step exchange (every image is independent); like the reference, which
            # all-reduces its logging scalars once per epoch (src/training/train_model.py:285-288), the
            # loss statistics are reduced once per run — one 8-float message, inside the timed region
            reduced = reduce_loss_stats(out, n)
        ev1.record()
        clocks.sample(3)                     # the GPU is still draining the queued steps: samples under load
        barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = _cabi.launch_count - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * n * args.steps / (elapsed_ms * 1e-3)
    loss_val = float(out[0].item())

    # ---- per-kernel durations for the roofline (separate pass, events on the launching stream) ----
    lib.yb_stage_timing(1)
    stage = []
    import ctypes
    buf = (ctypes.c_float * 3)()
    for _ in range(max(5, min(args.steps, 20))):
        step()
        _cabi.check(lib.yb_loss_last_stage_ms(buf), "yb_loss_last_stage_ms")
        stage.append([buf[0], buf[1], buf[2]])
    lib.yb_stage_timing(0)
    main_ms, match_ms, fin_ms = (statistics.mean(s[i] for s in stage) for i in range(3))   # launch order
    peak, peak_src = measured_peak_gbs()
    # fused_main_kernel = box role (reads the 64 box channels, writes their gradient) + class role (reads the nc
    # class channels, writes their gradient): every byte of preds read once, every byte of grad written once
    main_bytes = bytes_per_step
    main_gbs = main_bytes / (main_ms * 1e-3) / 1e9
    traffic = None                                   # dram read+write per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and (n, nc, a, preds.element_size()) == (128, 80, 8400, 4):
        traffic = json.load(open(tpath)).get("fused_main_kernel")
    roofline = {"bound": "hbm", "kernel": "fused_main_kernel", "achieved": main_gbs, "peak": peak, "unit": "GB/s",
                "frac": main_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": main_bytes, "ms_per_launch": main_ms,
                "other_kernels": {"match_kernel": {"ms": match_ms}, "finalize_kernel": {"ms": fin_ms}},
                "whole_step": {"algorithmic_bytes": bytes_per_step, "GB/s": bytes_per_step / (ms_per_step * 1e-3) / 1e9,
                               "frac": bytes_per_step / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- end to end through the public API, host inputs ----
    # Every step copies its own inputs from pinned host memory (the head output and the packed GT wire format of
    # data/collate.py::collate_fn_packed) and reads its loss scalars back.  As a data loader would, the copy of
    # step i+1 is issued on a copy stream while step i computes; all K copies lie inside the timed region.
    crit = YoloDFLQFLoss(num_classes=nc)
    with gpu_local_cpus(local) as numa_note:           # pinned staging buffers on the GPU's own NUMA node
        preds_pin = preds_h.pin_memory()
        packed_pin = pack_gt_host(gts_h, pin_memory=True)
    h2d = preds_pin.numel() * preds_pin.element_size() + packed_pin.gt.numel() * 4 + packed_pin.offsets.numel() * 4
    d2h = 3 * 4
    copy_stream = torch.cuda.Stream(device=dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            x = preds_pin.to(dev, non_blocking=True)
            pg = packed_pin.to(dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        return x, pg, done

    def e2e_run(k):
        parts, nxt = None, fetch()
        for i in range(k):
            x, pg, done = nxt
            torch.cuda.current_stream().wait_event(done)
            if i + 1 < k:
                nxt = fetch()
            x.record_stream(torch.cuda.current_stream())
            pg.gt.record_stream(torch.cuda.current_stream())
            pg.offsets.record_stream(torch.cuda.current_stream())
            x.requires_grad_(True)
            loss, parts = crit(x, pg, anchors_d, strides_d)            # parts: one D2H copy of the loss scalars
            loss.backward()
        return parts

    e2e_steps = max(3, min(args.steps, 10))
    parts = e2e_run(2)
    barrier()
    ev0.record()
    parts = e2e_run(e2e_steps)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)

    # ---- NMS (cfg4), reported alongside ----
    nms = None
    if rank == 0 or world > 1:
        y = syn.make_nms_input(64, nc, imgsz, 2024 + rank).to(dev)
        for _ in range(3):
            rows, cnt, _ = batched_nms_raw(y, 0.001, 0.7, 300, nc)
        barrier()
        ev0.record()
        nms_steps = max(5, min(args.steps, 20))
        for _ in range(nms_steps):
            rows, cnt, _ = batched_nms_raw(y, 0.001, 0.7, 300, nc)
        ev1.record()
        barrier()
        nms_ms = ev0.elapsed_time(ev1) / nms_steps
        if world > 1:
            t = torch.tensor([nms_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nms_ms = float(t.item())
        nms_bytes = y.numel() * 4 + 64 * 300 * 6 * 4
        nms = {"metric": "images/sec through batched class-aware NMS", "value": world * 64 / (nms_ms * 1e-3), "unit": "images/s",
               "ms_per_step": nms_ms, "config": {"workload": "cfg4: batch 64/GPU, 8400 candidates, nc=80, conf 0.001, IoU 0.7, max_det 300"},
               "hbm_frac_of_scan_roofline": nms_bytes / (nms_ms * 1e-3) / 1e9 / peak, "kept_min": int(cnt.min().item())}

    # ---- extra lines (not the headline): the task-aligned variant and the dense bf16 config ----
    def time_steps(fn, k):
        for _ in range(3):
            keep = fn()
        barrier()
        ev0.record()
        for _ in range(k):
            keep = fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / k
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    extra_steps = max(5, min(args.steps, 50))
    tal_ms = time_steps(lambda: fused_tal_loss(preds, gt, off, anchors_d, strides_d, nc, 1.5, 1.0, 1.5), extra_steps)
    tal = {"metric": "images/sec through decode+TAL assign+CIoU/DFL/BCE loss+bwd (no reference counterpart; parity vs in-repo oracle)",
           "value": world * n / (tal_ms * 1e-3), "unit": "images/s", "ms_per_step": tal_ms,
           "hbm_frac_whole_step": bytes_per_step / (tal_ms * 1e-3) / 1e9 / peak,
           "exchange": "all-reduce of [sum target scores, #fg] between assign and loss" if world > 1 else "none (1 GPU)"}
    p5, g5, a5, s5 = syn.make_loss_inputs(32, nc, 1280, 300, 1240 + rank, dtype=torch.bfloat16)
    p5 = p5.to(dev); a5 = a5.float().to(dev); s5 = s5.float().to(dev)
    gt5, off5, c5 = pack_gt([g.to(dev) for g in g5], dev)
    cfg5_ms = time_steps(lambda: fused_loss(p5, gt5, off5, max(c5), a5, s5, nc, 1.0, 1.5), extra_steps)
    cfg5_bytes = 2 * p5.numel() * p5.element_size()
    cfg5 = {"config": "cfg5: batch 32/GPU, 1280x1280 (33600 anchors), <=300 GT/img, bf16 head outputs", "value": world * 32 / (cfg5_ms * 1e-3),
            "unit": "images/s", "ms_per_step": cfg5_ms, "hbm_frac_whole_step": cfg5_bytes / (cfg5_ms * 1e-3) / 1e9 / peak}
    del p5, gt5

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ips, med, threads = cpu_loss_images_per_s(3, 1)
            cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{CPU_SAMPLE_IMAGES} images of the same workload per step, oracle/loss_oracle.py fwd+bwd on torch CPU, "
                             f"median of 3 steps ({med * 1e3:.0f} ms/step)"}
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "cfg2 per GPU: batch 128, 640x640 (8400 anchors, reg_max 16), 80 classes, <=100 GT/img, "
                                       "fused decode+assign+loss+backward (cfg3 = global batch 1024 at 8 GPUs)",
                           "global_batch": world * n, "gt_boxes_per_step_per_gpu": int(sum(counts)),
                           "l2_policy": "inputs+outputs are 1.24 GB per step, larger than the 126 MB L2",
                           "parallelism": (f"batch-sharded x{world}: no data-path collective, one 8-float all-reduce of the loss statistics per run"
                                           if world > 1 else "single GPU")},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                        "api": "YoloDFLQFLoss.forward + loss.backward on pinned host inputs (head output + packed GT); the next step's copy is prefetched on a copy stream",
                        "pinned_numa": numa_note.note},
                "gpu_launches": launches, "clocks": clocks.summary(), "loss": loss_val, "nms": nms, "tal": tal,
                "cfg5_bf16": cfg5}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
