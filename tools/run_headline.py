"""One or more passes of a named workload in device mode (inputs resident in HBM); used under ncu."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h
import cases
ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=1000); ap.add_argument("--k2", type=int, default=1)
ap.add_argument("--batch", type=int, default=65536); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--sp", action="store_true"); ap.add_argument("--n", type=int, default=500)
ap.add_argument("--atype", type=int, default=0); ap.add_argument("--otype", type=int, default=1)
a = ap.parse_args()
dev = "cuda:0"
rng = np.random.default_rng(0)
if a.sp:
    p, n = a.k, a.n
    A = cases.spd(rng, p); Om = cases.spd(rng, n)
    if a.atype: A = np.diag(np.diag(A))
    if a.otype: Om = np.diag(np.diag(Om))
    phi = rng.standard_normal((n, p)) / np.sqrt(p); mu = rng.standard_normal(p)
    t = [torch.from_numpy(x).to(dev) for x in (mu, A, phi, Om)]
    f = lambda: h.structured_precision_mvnorm_batch(a.batch, *t, False, a.atype, a.otype, seed0=1)
else:
    cov = cases.spd(rng, a.k); g = rng.standard_normal((a.k2, a.k)); r = rng.standard_normal(a.k2); mu = rng.standard_normal(a.k)
    t = [torch.from_numpy(x).to(dev) for x in (mu, cov, g, r)]
    out = torch.empty((a.batch, a.k), dtype=torch.float64, device=dev); info = torch.empty(a.batch, dtype=torch.int32, device=dev)
    f = lambda: h.hyperplane_truncated_mvnorm_batch(a.batch, *t, seed0=1, out=out, info=info)
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1), flush=True)
