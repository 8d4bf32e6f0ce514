"""Device-side timing of the ziggurat generator alone (htn_std_normal_fill), CUDA events on the torch stream.
HTN_RNG_CFG selects the kernel configuration (see htnk_normal_fill)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htnorm_b200 as h

dev = "cuda:0"
for gen in ("pcg64", "xrs128p"):
    for (B, n) in [(65536, 1000), (8192, 4096), (1 << 20, 64)]:
        z = torch.empty((B, n), dtype=torch.float64, device=dev)
        b = h.api._batch(B, gen, None, 1, False, 0, 0)
        f = lambda: h.lib.htn_std_normal_fill(C.byref(b), C.c_size_t(n), C.c_void_p(z.data_ptr()), None)
        for _ in range(2): f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts)
        print(f"cfg={os.environ.get('HTN_RNG_CFG','0')} {gen} B={B} n={n}: {t*1e3:.1f} us  {B*n/t/1e6:.1f} Gnormals/s  {B*n*8/t/1e6:.0f} GB/s", flush=True)
