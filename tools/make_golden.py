#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled here (oracle/_ref).

The reference's own tests pin properties only (tests/test_pyhtnorm.py:34-124), never values, so the
value-level golden vectors are outputs of the reference itself on seeded inputs (tests/cases.py).
Run where /root/reference exists:   python tools/make_golden.py
Each file stores the reference outputs (+ a checksum of the inputs, so a drifting input builder is
caught before it is mistaken for a parity failure).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import Reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(*arrays):
    """Fingerprint of the seeded inputs.  The builders go through BLAS (t @ t.T), whose last bit depends on
    the CPU, so the fingerprint is a few float statistics compared with a tolerance, not a hash."""
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        out += [a.sum(), (a * a).sum(), a[::97].sum(), float(a.size)]
    return np.array(out)


def main():
    ref = Reference()
    os.makedirs(GOLD, exist_ok=True)
    seeds = np.array(cases.RNG_SEEDS, dtype=np.uint64)
    out = {}
    for gen in ("pcg64", "xrs128p"):
        out[f"raw_{gen}"] = ref.raw(gen, len(seeds), 64, seeds=seeds)
        z, nd = ref.normals(gen, len(seeds), 3000, seeds=seeds)
        out[f"normals_{gen}"] = z
        out[f"ndraws_{gen}"] = nd
    np.savez_compressed(os.path.join(GOLD, "rng.npz"), seeds=seeds, **out)

    ht = {}
    for name in cases.HT_CASES:
        c = cases.ht_case(name)
        x, info, rc = ref.hyperplane(c["gen"], c["nsamples"], c["mean"], c["cov"], c["g"], c["r"],
                                     diag=c["diag"], seed0=c["seed0"])
        ht[name + "__x"] = x
        ht[name + "__info"] = info
        ht[name + "__digest"] = digest(c["mean"], c["cov"], c["g"], c["r"])
    c = cases.ht_independent_case()
    x, info, rc = ref.hyperplane(c["gen"], c["nsamples"], c["mean"], c["cov"], c["g"], c["r"],
                                 independent=True, seed0=c["seed0"])
    ht["independent__x"] = x
    ht["independent__info"] = info
    ht["independent__digest"] = digest(c["mean"], c["cov"], c["g"], c["r"])
    np.savez_compressed(os.path.join(GOLD, "hyperplane.npz"), **ht)

    sp = {}
    for name in cases.SP_CASES:
        c = cases.sp_case(name)
        x, info, rc = ref.structured(c["gen"], c["nsamples"], c["mean"], c["a"], c["phi"], c["omega"],
                                     c["struct_mean"], c["a_type"], c["o_type"], seed0=c["seed0"])
        sp[name + "__x"] = x
        sp[name + "__info"] = info
        sp[name + "__digest"] = digest(c["mean"], c["a"], c["phi"], c["omega"])
    np.savez_compressed(os.path.join(GOLD, "structured.npz"), **sp)
    # the -DHTNORM_COLMAJOR build (R binding): g / phi handed over column-major
    refc = Reference(colmajor=True)
    cm = {}
    for name in ("k4_dense_64", "k20_dense_100", "k3_diag_40", "k1_dense_50"):
        c = cases.ht_case(name)
        gcm = np.ascontiguousarray(c["g"].T)          # k x k2 row-major == k2 x k column-major
        x, info, rc = refc.hyperplane(c["gen"], c["nsamples"], c["mean"], c["cov"], gcm.reshape(c["g"].shape),
                                      c["r"], diag=c["diag"], seed0=c["seed0"])
        cm["ht_" + name + "__x"] = x
    for name in ("nn_plain", "nn_struct", "dn_struct", "nd_plain_mid", "in_plain"):
        c = cases.sp_case(name)
        pcm = np.ascontiguousarray(c["phi"].T).reshape(c["phi"].shape)
        x, info, rc = refc.structured(c["gen"], c["nsamples"], c["mean"], c["a"], pcm, c["omega"],
                                      c["struct_mean"], c["a_type"], c["o_type"], seed0=c["seed0"])
        cm["sp_" + name + "__x"] = x
    np.savez_compressed(os.path.join(GOLD, "colmajor.npz"), **cm)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
