"""BASELINE.json configs 3-5 (per-GPU shapes), device-resident, CUDA events; prints one line per config.
cfg3: structured precision p=2000 n=500, Omega diagonal, B=4096 (A dense and A diagonal)
cfg4: p=8192 n=1024 dense/dense, struct_mean, B=256
cfg5 (one GPU's shard): hyperplane k=16384 k2=64, B=8192, shared covariance"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h

dev = "cuda:0"
which = sys.argv[1:] or ["cfg3", "cfg3d", "cfg4", "cfg5"]

def spd_dev(d, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.randn(d, d, dtype=torch.float64, device=dev, generator=g)
    s = t @ t.T / d
    s += torch.eye(d, dtype=torch.float64, device=dev)
    return (s + s.T) / 2

def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

def randn(*s, seed=0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    return torch.randn(*s, dtype=torch.float64, device=dev, generator=g)

for w in which:
    if w in ("cfg3", "cfg3d"):
        p, n, B = 2000, 500, 4096
        a = spd_dev(p, 3) if w == "cfg3" else torch.diag(torch.rand(p, dtype=torch.float64, device=dev) + 0.5)
        phi = randn(n, p, seed=4) / p ** 0.5
        om = torch.diag(torch.rand(n, dtype=torch.float64, device=dev) + 0.5)
        mu = randn(p, seed=5)
        out = torch.empty(B, p, dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
        at = 0 if w == "cfg3" else 1
        f = lambda: h.structured_precision_mvnorm_batch(B, mu, a, phi, om, False, at, 1, seed0=1, out=out, info=info)
        ms = timeit(f)
        print(f"{w}: p={p} n={n} B={B} a_type={at}: {ms:.3f} ms  {B/ms*1e3:.3e} samples/s", flush=True)
    elif w == "cfg4":
        p, n, B = 8192, 1024, 256
        a = spd_dev(p, 6); om = spd_dev(n, 7)
        phi = randn(n, p, seed=8) / p ** 0.5
        t = randn(n, seed=9)
        out = torch.empty(B, p, dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
        f = lambda: h.structured_precision_mvnorm_batch(B, t, a, phi, om, True, 0, 0, seed0=1, out=out, info=info)
        ms = timeit(f)
        flops = p**3/3 + 2*p*p*n + 2*n*n*p + n**3/3 + n**3 + B*(p*p + n*n + 4*n*p + 2*n*n + n)
        print(f"{w}: p={p} n={n} B={B} dense/dense struct_mean: {ms:.3f} ms  {flops/ms/1e9:.2f} TFLOP/s algorithmic", flush=True)
    elif w == "cfg5":
        k, k2, B = 16384, 64, 8192
        cov = spd_dev(k, 10)
        g = randn(k2, k, seed=11); r = randn(k2, seed=12); mu = randn(k, seed=13)
        out = torch.empty(B, k, dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
        f = lambda: h.hyperplane_truncated_mvnorm_batch(B, mu, cov, g, r, seed0=1, out=out, info=info)
        ms = timeit(f, n=2)
        flops = k**3/3 + 2*k*k*k2 + 2*k*k2*k2 + k2**3/3 + B*(k*k + 4*k*k2 + 2*k2*k2)
        resid = float((out @ g.T - r).abs().max())
        print(f"{w}: k={k} k2={k2} B={B}: {ms:.2f} ms  {flops/ms/1e9:.2f} TFLOP/s algorithmic  max|Gx-r|={resid:.2e}", flush=True)
    del_vars = None
    torch.cuda.empty_cache()
