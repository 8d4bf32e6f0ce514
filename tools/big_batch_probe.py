import os, sys
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h, cases
K = 1000
rng = np.random.default_rng(0)
cov = cases.spd(rng, K); g = np.ones((1, K)); r = np.zeros(1); mean = rng.standard_normal(K)
d_in = [torch.from_numpy(x).cuda() for x in (mean, cov, g, r)]
for B in (262144, 1 << 20):
    out = torch.empty((B, K), dtype=torch.float64, device="cuda"); info = torch.empty(B, dtype=torch.int32, device="cuda")
    f = lambda i: h.hyperplane_truncated_mvnorm_batch(B, *d_in, seed0=1 + i, out=out, info=info)
    f(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(1); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    resid = float(out.sum(dim=1).abs().max())
    # shard invariance against a small call at an offset in the second Z~ chunk
    o2 = torch.empty((64, K), dtype=torch.float64, device="cuda"); i2 = torch.empty(64, dtype=torch.int32, device="cuda")
    off = B - 64
    h.hyperplane_truncated_mvnorm_batch(64, *d_in, seed0=2 + off, out=o2, info=i2); torch.cuda.synchronize()
    same = bool(torch.equal(o2, out[off:]))
    print(f"B={B}: {ms:.2f} ms -> {B/ms*1e3:.3e} samples/s, max|sum x|={resid:.2e}, tail rows == separate call: {same}, info nonzero {int((info!=0).sum())}", flush=True)
    del out, info
    torch.cuda.empty_cache()
