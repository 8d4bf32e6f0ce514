#!/usr/bin/env python
"""Golden rows at the FULL sizes of BASELINE.json configs[3] and configs[4] from the UNMODIFIED reference compiled
here (oracle/_ref): tests/golden/fullsize.npz.  One reference call per row (the reference refactorises per call:
a 16384 x 16384 dpotrf each), OpenBLAS multi-threaded: a few minutes in all.  Run where /root/reference exists:
    OPENBLAS_NUM_THREADS=8 python tools/make_golden_fullsize.py
Stored next to each row set: a fingerprint of the seeded inputs (tools/make_golden.py `digest`)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cases  # noqa: E402
from make_golden import digest  # noqa: E402
from oracle import Reference  # noqa: E402


def main():
    ref = Reference()
    out = {}
    c = cases.cfg4_inputs()
    t0 = time.time()
    seeds = (c["seed0"] + c["spot"]).astype(np.uint64)
    x, info, rc = ref.structured(c["gen"], len(seeds), c["mean"], c["a"], c["phi"], c["omega"], c["struct_mean"],
                                 c["a_type"], c["o_type"], seeds=seeds)
    assert rc == 0 and not info.any()
    out["cfg4__x"], out["cfg4__spot"] = x, c["spot"]
    out["cfg4__digest"] = digest(c["mean"], c["a"], c["phi"], c["omega"])
    print("cfg4: %d rows in %.1f s" % (len(seeds), time.time() - t0), flush=True)
    c = cases.cfg5_inputs()
    t0 = time.time()
    seeds = (c["seed0"] + c["spot"]).astype(np.uint64)
    x, info, rc = ref.hyperplane(c["gen"], len(seeds), c["mean"], c["cov"], c["g"], c["r"], seeds=seeds)
    assert rc == 0 and not info.any()
    out["cfg5__x"], out["cfg5__spot"] = x, c["spot"]
    out["cfg5__digest"] = digest(c["mean"], c["cov"], c["g"], c["r"])
    resid = np.abs(x.astype(np.longdouble) @ c["g"].T.astype(np.longdouble) - c["r"]).max()
    print("cfg5: %d rows in %.1f s, reference's own max |Gx - r| = %.2e" % (len(seeds), time.time() - t0, float(resid)),
          flush=True)
    path = os.path.join(ROOT, "tests", "golden", "fullsize.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
