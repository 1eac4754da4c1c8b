"""H2D bandwidth from pinned memory: one copy vs chunked copies on several streams."""
import torch, time
dev = torch.device('cuda:0')
n = 128 * 144 * 8400
h = torch.empty(n, dtype=torch.float32).pin_memory()
h.normal_()
d = torch.empty(n, dtype=torch.float32, device=dev)
def timed(f, reps=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def single(): d.copy_(h, non_blocking=True)
print('single copy  %.3f ms  %.1f GB/s' % (timed(single) * 1e3, n * 4 / timed(single) / 1e9))
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    step = (n + k - 1) // k
    def chunked():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    t = timed(chunked)
    print('%d streams    %.3f ms  %.1f GB/s' % (k, t * 1e3, n * 4 / t / 1e9))
for mb in (16, 64):
    step = mb * 1024 * 1024 // 4
    def seq():
        for i in range(0, n, step): d[i:i + step].copy_(h[i:i + step], non_blocking=True)
    t = timed(seq)
    print('%d MB chunks, one stream %.3f ms  %.1f GB/s' % (mb, t * 1e3, n * 4 / t / 1e9))
