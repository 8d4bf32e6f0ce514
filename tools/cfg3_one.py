"""one cfg3 call after warm-up (for launch lists): python tools/cfg3_one.py [cfg3|cfg4]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cases
import htnorm_b200 as h
import bench_configs as bc
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dev = "cuda:0"
if which == "cfg3":
    rng = np.random.default_rng(3)
    p, n, B = 2000, 500, 4096
    a = torch.from_numpy(cases.spd(rng, p)).to(dev)
    phi = torch.from_numpy(rng.standard_normal((n, p)) / np.sqrt(p)).to(dev)
    om = torch.from_numpy(np.diag(rng.uniform(0.5, 1.5, n))).to(dev)
    mu = torch.from_numpy(rng.standard_normal(p)).to(dev)
    args = (B, mu, a, phi, om, False, 0, 1)
else:
    p, n, B = 8192, 1024, 256
    a = bc.spd_lowrank_device(torch, p, dev, 41); om = bc.spd_lowrank_device(torch, n, dev, 42)
    g_ = torch.Generator(device=dev).manual_seed(43)
    phi = torch.randn(n, p, dtype=torch.float64, device=dev, generator=g_) / np.sqrt(p)
    mu = torch.randn(n, dtype=torch.float64, device=dev, generator=g_)
    args = (B, mu, a, phi, om, True, 0, 0)
out = torch.empty(B, p, dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(3):
    h.structured_precision_mvnorm_batch(*args, seed0=12345, out=out, info=info)
torch.cuda.synchronize()
