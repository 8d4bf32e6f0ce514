"""Device-side time of the shared-covariance hyperplane path over a sweep of orders k (odd, unaligned and aligned),
k2 and batch sizes: a check that no shape falls off a performance cliff (slow epilogue, fallback kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h
import cases
dev = "cuda:0"

def timeit(fn, n=4, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

shapes = [(1000, 1, 65536), (999, 1, 65536), (1001, 1, 65536), (1024, 1, 65536), (1000, 4, 65536), (1000, 16, 65536),
          (1000, 17, 65536), (500, 1, 65536), (250, 2, 65536), (2000, 1, 32768), (4000, 8, 16384), (1000, 1, 1000),
          (1000, 1, 100), (129, 1, 65536), (100, 3, 262144)]
for (k, k2, B) in shapes:
    rng = np.random.default_rng(0)
    cov = cases.spd(rng, k); g = rng.standard_normal((k2, k)); r = rng.standard_normal(k2); mu = rng.standard_normal(k)
    t = [torch.from_numpy(x).to(dev) for x in (mu, cov, g, r)]
    out = torch.empty((B, k), dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
    f = lambda: h.hyperplane_truncated_mvnorm_batch(B, *t, seed0=1, out=out, info=info)
    ms = timeit(f)
    flops = k ** 3 / 3 + B * (k * k + 4 * k * k2)
    resid = float((out @ t[2].T - t[3]).abs().max())
    print(f"k={k:5d} k2={k2:3d} B={B:7d}: {ms:8.3f} ms  {B / ms * 1e3:10.3e} samples/s  {flops / ms / 1e9:6.2f} TFLOP/s  "
          f"out {8 * B * k / ms / 1e6:7.1f} GB/s  |Gx-r| {resid:.1e}", flush=True)
