"""latency of ONE draw through the reference's own signature (host arrays in / out), k = 1000, k2 = 1"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import htnorm_b200 as h
import bench
mean, cov, g, r = bench.make_inputs(np)
gen = h.HTNGenerator(12345)
out = np.empty(1000)
for _ in range(5): gen.hyperplane_truncated_mvnorm(mean, cov, g, r, out=out)
t0 = time.perf_counter()
n = 50
for _ in range(n): gen.hyperplane_truncated_mvnorm(mean, cov, g, r, out=out)
print(f"single draw k=1000 (library generator, pageable host arrays): {(time.perf_counter()-t0)/n*1e3:.3f} ms per call")
for k in (64, 256):
    rng = np.random.default_rng(k)
    import cases
    c = cases.spd(rng, k); G = rng.standard_normal((2, k)); rr = rng.standard_normal(2); mu = rng.standard_normal(k)
    o = np.empty(k)
    for _ in range(5): gen.hyperplane_truncated_mvnorm(mu, c, G, rr, out=o)
    t0 = time.perf_counter()
    for _ in range(n): gen.hyperplane_truncated_mvnorm(mu, c, G, rr, out=o)
    print(f"single draw k={k}: {(time.perf_counter()-t0)/n*1e3:.3f} ms per call")
