import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cases
import htnorm_b200 as h
dev = "cuda:0"; k, P, k2 = 256, 4096, 4
rng = np.random.default_rng(k)
base = torch.from_numpy(np.stack([cases.spd(rng, k) for _ in range(8)])).to(dev)
cov = base.repeat(P // 8, 1, 1).contiguous()
g = torch.randn(P, k2, k, dtype=torch.float64, device=dev); r = torch.randn(P, k2, dtype=torch.float64, device=dev)
mu = torch.randn(P, k, dtype=torch.float64, device=dev)
out = torch.empty(P, k, dtype=torch.float64, device=dev); info = torch.empty(P, dtype=torch.int32, device=dev)
for _ in range(2):
    h.hyperplane_truncated_mvnorm_batch(P, mu, cov, g, r, independent=True, seed0=1, out=out, info=info)
torch.cuda.synchronize()
