import os, time, torch, pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print('cpus allowed', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], '...')
try:
    words = pynvml.nvmlDeviceGetCpuAffinity(h, 4)
    print('gpu0 ideal cpu mask words', [hex(w) for w in words])
except Exception as e:
    print('affinity query failed', e)
os.system("lscpu | grep -i 'numa\\|socket\\|model name' | head -8")
dev = torch.device('cuda:0')
n = 128 * 144 * 8400
def rate(h_t):
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(2): d.copy_(h_t, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): d.copy_(h_t, non_blocking=True)
    torch.cuda.synchronize(); return n * 4 * 5 / (time.perf_counter() - t0) / 1e9
a = torch.empty(n).pin_memory(); a.zero_()
print('default placement  %.1f GB/s' % rate(a))
old = os.sched_getaffinity(0)
try:
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print('now on', len(os.sched_getaffinity(0)), 'cpus')
    b = torch.empty(n).pin_memory(); b.zero_()
    print('gpu-local placement %.1f GB/s' % rate(b))
except Exception as e:
    print('set affinity failed', e)
os.sched_setaffinity(0, old)
