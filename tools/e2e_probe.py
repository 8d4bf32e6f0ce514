import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h, cases
B, K = 65536, 1000
rng = np.random.default_rng(0)
cov = cases.spd(rng, K); g = np.ones((1, K)); r = np.zeros(1); mean = rng.standard_normal(K)
h_in = [torch.from_numpy(x).pin_memory() for x in (mean, cov, g, r)]
h_out = torch.empty((B, K), dtype=torch.float64).pin_memory(); h_info = torch.empty(B, dtype=torch.int32).pin_memory()
np_in = [x.numpy() for x in h_in]
def step(i): h.hyperplane_truncated_mvnorm_batch(B, *np_in, gen="pcg64", seed0=1 + i * B, out=h_out.numpy(), info=h_info.numpy())
step(0); step(1)
ts = []
for i in range(5):
    t0 = time.perf_counter(); step(2 + i); ts.append(time.perf_counter() - t0)
print(f"HTN_SINK_CHUNK={os.environ.get('HTN_SINK_CHUNK','8192')}: best {min(ts)*1e3:.2f} ms  median {sorted(ts)[2]*1e3:.2f} ms -> {B/min(ts):.3e} samples/s")
