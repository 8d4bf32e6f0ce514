import os, sys, time
ROOT = "/root/repo" if os.path.isdir("/root/repo/htnorm_b200") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h, cases
K = 1000
rng = np.random.default_rng(0)
cov = cases.spd(rng, K); g = np.ones((1, K)); r = np.zeros(1); mean = rng.standard_normal(K)
h_in = [torch.from_numpy(x).pin_memory() for x in (mean, cov, g, r)]
np_in = [x.numpy() for x in h_in]
d_in = [torch.from_numpy(x).cuda() for x in (mean, cov, g, r)]
for B in (128, 2048, 65536):
    h_out = torch.empty((B, K), dtype=torch.float64).pin_memory(); h_info = torch.empty(B, dtype=torch.int32).pin_memory()
    d_out = torch.empty((B, K), dtype=torch.float64, device="cuda"); d_info = torch.empty(B, dtype=torch.int32, device="cuda")
    def host(i): h.hyperplane_truncated_mvnorm_batch(B, *np_in, seed0=1 + i, out=h_out.numpy(), info=h_info.numpy())
    def dev(i):
        h.hyperplane_truncated_mvnorm_batch(B, *d_in, seed0=1 + i, out=d_out, info=d_info); torch.cuda.synchronize()
    for f, name in ((host, "host"), (dev, "device+sync")):
        f(0); f(1)
        ts = []
        for i in range(7):
            t0 = time.perf_counter(); f(2 + i); ts.append(time.perf_counter() - t0)
        print(f"B={B:6d} {name:12s}: best {min(ts)*1e3:.3f} ms  median {sorted(ts)[3]*1e3:.3f} ms", flush=True)

# single-sample API (reference-compatible entry point, generator advanced on the GPU and written back)
gen = h.HTNGenerator(7)
out1 = np.empty(K)
def one(i): gen.hyperplane_truncated_mvnorm(mean, cov, g, r, out=out1)
one(0); one(1)
ts = []
for i in range(9):
    t0 = time.perf_counter(); one(i); ts.append(time.perf_counter() - t0)
print(f"single-sample htn_hyperplane_truncated_mvn, k={K} (pageable numpy inputs): best {min(ts)*1e3:.3f} ms  median {sorted(ts)[4]*1e3:.3f} ms", flush=True)
