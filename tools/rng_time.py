"""generator timing: streams x normals, device memory, CUDA events (HTN_RNG_COOP selects the kernel)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htnorm_b200 as h
shapes = [(1, 1000), (256, 9216), (4096, 2500), (8192, 1000), (16384, 1000), (32768, 1000), (65536, 1000)]
for gen in (0, 1):
    for ns, n in shapes:
        out = torch.empty(ns, n, dtype=torch.float64, device="cuda")
        b = h._lib.Batch()
        h.lib.htn_batch_init(C.byref(b), C.c_size_t(ns))
        b.gen, b.seed0, b.mem, b.device = gen, 12345, h._lib.HTN_MEM_DEVICE, 0
        b.stream = torch.cuda.current_stream().cuda_stream
        fn = lambda: h.lib.htn_std_normal_fill(C.byref(b), C.c_size_t(n), C.c_void_p(out.data_ptr()), None)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"gen {gen} streams {ns:6d} x {n:5d}: {ms*1e3:9.1f} us  {ns*n/ms/1e6:8.1f} G normals/s", flush=True)
