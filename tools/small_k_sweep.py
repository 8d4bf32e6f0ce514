"""independent small problems, k2 = 4, 2^19 problems: time per k (HTN_SMALL_IMPL=0 selects the generic kernel)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cases
import htnorm_b200 as h
dev = "cuda:0"
for k in (16, 32, 48, 64):
    P = 1 << 19
    rng = np.random.default_rng(k)
    base = torch.from_numpy(np.stack([cases.spd(rng, k) for _ in range(64)])).to(dev)
    g_ = torch.Generator(device=dev).manual_seed(2)
    cov = base.repeat(P // 64, 1, 1); cov.mul_(1.0 + torch.rand(P, 1, 1, dtype=torch.float64, device=dev, generator=g_))
    g = torch.randn(P, 4, k, dtype=torch.float64, device=dev, generator=g_); r = torch.randn(P, 4, dtype=torch.float64, device=dev, generator=g_)
    mu = torch.randn(P, k, dtype=torch.float64, device=dev, generator=g_)
    out = torch.empty(P, k, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
    f = lambda: h.hyperplane_truncated_mvnorm_batch(P, mu, cov, g, r, independent=True, seed0=1, out=out, info=info)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    nbytes = 8.0 * (k * k + 4 * k + 2 * k + 4) * P
    print(f"k={k} k2=4 P=2^19: {ms:.3f} ms  {P/ms/1e3:.1f} M problems/s  {nbytes/ms/1e6:.0f} GB/s = {nbytes/ms/1e6/6535.1:.2f} of HBM roofline", flush=True)
