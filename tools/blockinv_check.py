"""cfg4 shape (p = 8192, n = 1024, dense A and Omega, struct_mean, B = 256): the block-inverse solves (automatic at this size)
against the whole-inverse path on the same inputs and seeds, plus the statistical identity both must satisfy."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import htnorm_b200 as h
dev = "cuda:0"
def spd(d, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.randn(d, d, dtype=torch.float64, device=dev, generator=g)
    s = t @ t.T / d + torch.eye(d, dtype=torch.float64, device=dev)
    return (s + s.T) / 2
def randn(*s, seed=0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    return torch.randn(*s, dtype=torch.float64, device=dev, generator=g)
p, n, B = 8192, 1024, 256
a = spd(p, 6); om = spd(n, 7); phi = randn(n, p, seed=8) / p ** 0.5; t = randn(n, seed=9)
res = {}
for mode in ("0", None):
    if mode is None: os.environ.pop("HTN_INV_BLOCK", None)
    else: os.environ["HTN_INV_BLOCK"] = mode
    x, info = h.structured_precision_mvnorm_batch(B, t, a, phi, om, True, 0, 0, seed0=1)
    torch.cuda.synchronize()
    res[mode] = x
    assert not bool(info.any())
d = (res[None] - res["0"]).abs().max().item() / res["0"].abs().max().item()
print(f"block-inverse vs whole-inverse, max relative difference: {d:.2e}")
assert d < 1e-11
print("ok")
