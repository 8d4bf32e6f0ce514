import os, sys
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import htnorm_b200 as h, cases
c = cases.ht_case("k7_dense_333"); s = cases.sp_case("nn_plain"); m = cases.sp_case("nn_struct_mid")
mid = cases.ht_independent_case(nproblems=40, k=136, k2=3, rng_seed=1); sml = cases.ht_independent_case(nproblems=64)
def used(): f, t = torch.cuda.mem_get_info(); return (t - f) / 2**20
for it in range(3):
    for i in range(400):
        h.hyperplane_truncated_mvnorm_batch(300, c["mean"], c["cov"], c["g"], c["r"], seed0=i)
        h.structured_precision_mvnorm_batch(50, s["mean"], s["a"], s["phi"], s["omega"], s["struct_mean"], s["a_type"], s["o_type"], seed0=i)
        if i % 8 == 0:   # forked inversion + early Y1 (n, p large enough for the side streams)
            h.structured_precision_mvnorm_batch(200, m["mean"], m["a"], m["phi"], m["omega"], m["struct_mean"], m["a_type"], m["o_type"], seed0=i)
        h.HTNGenerator(i).hyperplane_truncated_mvnorm(c["mean"], c["cov"], c["g"], c["r"])
        if i % 4 == 0:   # independent problems: warp-per-problem kernel and the batched mid-size path (split over two streams)
            h.hyperplane_truncated_mvnorm_batch(64, sml["mean"], sml["cov"], sml["g"], sml["r"], independent=True, seed0=i)
            os.environ["HTN_MID_SPLIT"] = "8"
            h.hyperplane_truncated_mvnorm_batch(40, mid["mean"], mid["cov"], mid["g"], mid["r"], independent=True, seed0=i)
            os.environ.pop("HTN_MID_SPLIT")
    torch.cuda.synchronize()
    print(f"after {400*(it+1)} x 3 calls: {used():.0f} MiB in use on the device", flush=True)
