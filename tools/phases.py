"""Phase timing (setup / coeff / gemm) of the shared-covariance hyperplane path for a given shape."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h
import cases
k, k2, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = "cuda:0"
rng = np.random.default_rng(0)
cov = cases.spd(rng, k); g = rng.standard_normal((k2, k)); r = rng.standard_normal(k2); mu = rng.standard_normal(k)
t = [torch.from_numpy(x).to(dev) for x in (mu, cov, g, r)]
out = torch.empty((B, k), dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
for rep in range(3):
    h.phase_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.hyperplane_truncated_mvnorm_batch(B, *t, seed0=1, out=out, info=info); e1.record(); torch.cuda.synchronize()
    ph = h.phase_timing()
    print(f"k={k} k2={k2} B={B}: total {e0.elapsed_time(e1):.2f} ms  phases {ph}  setup TF {(k**3/3+2*k*k*k2)/(ph['setup']*1e-3)/1e12:.1f}  gemm TF {B*(k*k+2*k*k2)/(ph['gemm']*1e-3)/1e12:.1f}", flush=True)
resid = float((out @ t[2].T - t[3]).abs().max()); print("max |Gx - r| =", resid)
