#!/usr/bin/env bash
cd /root/repo
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import htnorm_b200 as h
from oracle import Oracle
o = Oracle()
for gen in ("pcg64", "xrs128p"):
    for ns, n in [(1, 1), (3, 40), (7, 3000), (33, 31)]:
        z, nd = h.standard_normals(ns, n, gen=gen, seed0=7)
        oz, ond = o.normals(gen, ns, n, seed0=7)
        bad = np.argwhere(z.view(np.uint64) != oz.view(np.uint64))
        print(gen, ns, n, "ndraws equal:", np.array_equal(nd, ond), "mismatches:", len(bad), "max abs", np.abs(z - oz).max())
        if len(bad):
            b, i = bad[0]
            print("  first mismatch stream", b, "pos", i, "draw j =", n - 1 - i, z[b, i], oz[b, i], "nd", nd[b], ond[b])
            bb = bad[bad[:, 0] == b]
            print("  positions in that stream:", (n - 1 - bb[:, 1])[::-1][:20])
PY
