import ctypes, os, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libcopy256.so'))
lib.run_copy.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
dev = torch.device('cuda:0')
n = 128 * 144 * 8400
src = torch.randn(n, device=dev); dst = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
names = {0: '128-bit x4', 1: '128-bit x8', 2: '256-bit x2', 3: '256-bit x4', 4: '256-bit x4 + L2::evict_first', 9: 'torch copy_'}
for rep in range(2):
    for mode in (0, 1, 2, 3, 4, 9):
        f = (lambda: dst.copy_(src)) if mode == 9 else (lambda: lib.run_copy(mode, src.data_ptr(), dst.data_ptr(), n * 4, st))
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if mode != 9: assert torch.equal(src, dst)
        print(f'{names[mode]:32s} {ms*1e3:7.1f} us  {2*n*4/ms/1e6:7.0f} GB/s')
