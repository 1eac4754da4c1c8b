"""cfg5 through the Python entry point vs the bare ABI loop: where do 10 us per step go?"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from custom_yolo_implmentation_b200.model.losses import fused_loss, pack_gt
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
nc = 80
for seed in (1240, 1236):
    p5, g5, a5, s5 = syn.make_loss_inputs(32, nc, 1280, 300, seed, dtype=torch.bfloat16)
    p5 = p5.to(dev); a5 = a5.float().to(dev); s5 = s5.float().to(dev)
    gt5, off5, c5 = pack_gt([g.to(dev) for g in g5], dev)
    for hint in ("auto", None):
        def step():
            return fused_loss(p5, gt5, off5, max(c5), a5, s5, nc, 1.0, 1.5, grid_hint=hint)
        for _ in range(5): keep = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(200): keep = step()
        e1.record(); th = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"seed {seed} hint {hint}: gpu {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/step, host {th / 200 * 1e6:.1f} us/step, gt {sum(c5)} gmax {max(c5)}", flush=True)
