"""batched mid-size independent problems: P problems of order k, k2 = 4, device-resident; ms and fraction of the DMMA peak
on SURVEY 8(d)'s algorithmic flops (own covariance per problem)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cases
import htnorm_b200 as h
dev = "cuda:0"
peak = h.fp64_tensor_peak_tflops(0, 50.0)
shapes = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(256, 4096), (512, 1024), (1000, 256), (192, 4096)]
for k, P in shapes:
    k2 = 4
    rng = np.random.default_rng(k)
    base = torch.from_numpy(np.stack([cases.spd(rng, k) for _ in range(8)])).to(dev)
    g_ = torch.Generator(device=dev).manual_seed(2)
    cov = base.repeat(P // 8, 1, 1); cov.mul_(1.0 + torch.rand(P, 1, 1, dtype=torch.float64, device=dev, generator=g_))
    g = torch.randn(P, k2, k, dtype=torch.float64, device=dev, generator=g_); r = torch.randn(P, k2, dtype=torch.float64, device=dev, generator=g_)
    mu = torch.randn(P, k, dtype=torch.float64, device=dev, generator=g_)
    out = torch.empty(P, k, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
    f = lambda: h.hyperplane_truncated_mvnorm_batch(P, mu, cov, g, r, independent=True, seed0=1, out=out, info=info)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = P * (k**3 / 3 + k * k + 2 * k * k * k2 + 2 * k * k2 * k2 + k2**3 / 3 + 2 * k2 * k2 + 4 * k * k2)
    resid = float((torch.einsum("bck,bk->bc", g[:64], out[:64]) - r[:64]).abs().max())
    print(f"k={k} k2={k2} P={P}: {ms:.3f} ms  {P/ms*1e3:.0f} problems/s  {flops/ms/1e9:.2f} TFLOP/s = {flops/ms/1e9/peak:.2f} of DMMA peak  bad={int((info!=0).sum())} resid={resid:.1e}", flush=True)
