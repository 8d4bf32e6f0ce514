"""CPU: pin the oracle (plain-C restatement) against the golden vectors generated from the unmodified
reference (tools/make_golden.py) and against the reference binary itself when it is present."""
import hashlib
import os

import numpy as np
import pytest

import cases
from conftest import REL_TOL, ROOT, rel_err


def digest(*arrays):
    """Fingerprint of the seeded inputs.  The builders go through BLAS (t @ t.T), whose last bit depends on
    the CPU, so the fingerprint is a few float statistics compared with a tolerance, not a hash."""
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        out += [a.sum(), (a * a).sum(), a[::97].sum(), float(a.size)]
    return np.array(out)


# known answers extracted from the compiled reference
KNOWN_U64 = {
    ("pcg64", 12345): [0x942315b95a694068, 0xce625b6b0ccae580, 0xfe847abf0a980280, 0x0440f8fe6ea3db20],
    ("pcg64", 0): [0x412162a4d88b9532, 0x01a0e8a6edd319f1],
    ("pcg64", 10): [0x887109f4e30ecc37],
    ("xrs128p", 12345): [0x56805f3ea0e50a8d, 0xe8c8fb10e799d07a, 0x67b37a0b3f115e50, 0x23a7130d9ba9ee0a],
    ("xrs128p", 0): [0x509946a41cd733a3],  # = SplitMix64(0)[0] + SplitMix64(0)[1], a public known answer
}
KNOWN_NORMALS_PCG64_12345 = [0.864850857323908, -0.6901464382644238, 1.464212930958405,
                             -0.10960342167368162, -0.41965070960937056, 1.720698313831757,
                             -0.3554770333072357, -1.0935129658967742]


@pytest.mark.parametrize("key", list(KNOWN_U64))
def test_known_answer_streams(oracle, key):
    gen, seed = key
    want = KNOWN_U64[key]
    got = oracle.raw(gen, 1, len(want), seed0=seed)[0]
    assert [int(v) for v in got] == want


def test_known_answer_normals(oracle):
    z, nd = oracle.normals("pcg64", 1, 8, seed0=12345)
    # reverse fill (reference src/htnorm_distributions.c:12-13): the first draw lands in out[n-1]
    assert z[0][::-1].tolist() == KNOWN_NORMALS_PCG64_12345


@pytest.mark.parametrize("gen", ["pcg64", "xrs128p"])
def test_rng_golden(oracle, gold, gen):
    g = gold["rng"]
    seeds = g["seeds"]
    raw = oracle.raw(gen, len(seeds), 64, seeds=seeds)
    assert np.array_equal(raw, g[f"raw_{gen}"])
    z, nd = oracle.normals(gen, len(seeds), 3000, seeds=seeds)
    assert np.array_equal(z.view(np.uint64), g[f"normals_{gen}"].view(np.uint64))
    assert np.array_equal(nd, g[f"ndraws_{gen}"])




@pytest.mark.parametrize("name", cases.HT_CASES)
def test_hyperplane_golden(oracle, gold, name):
    c = cases.ht_case(name)
    g = gold["hyperplane"]
    assert np.allclose(digest(c["mean"], c["cov"], c["g"], c["r"]), g[name + "__digest"], rtol=1e-10), "input builder drifted"
    x, info, rc = oracle.hyperplane(c["gen"], c["nsamples"], c["mean"], c["cov"], c["g"], c["r"],
                                    diag=c["diag"], seed0=c["seed0"])
    assert np.array_equal(info, g[name + "__info"])
    tol = 1e-7 if name == "readme_k1000" else REL_TOL  # README cov = t t^T is ill-conditioned (kappa ~ 1e9)
    assert rel_err(x, g[name + "__x"]) < tol


def test_hyperplane_independent_golden(oracle, gold):
    c = cases.ht_independent_case()
    g = gold["hyperplane"]
    assert np.allclose(digest(c["mean"], c["cov"], c["g"], c["r"]), g["independent__digest"], rtol=1e-10)
    x, info, rc = oracle.hyperplane(c["gen"], c["nsamples"], c["mean"], c["cov"], c["g"], c["r"],
                                    independent=True, seed0=c["seed0"])
    assert rc == 0 and not info.any()
    assert rel_err(x, g["independent__x"]) < REL_TOL


@pytest.mark.parametrize("name", cases.SP_CASES)
def test_structured_golden(oracle, gold, name):
    c = cases.sp_case(name)
    g = gold["structured"]
    assert np.allclose(digest(c["mean"], c["a"], c["phi"], c["omega"]), g[name + "__digest"], rtol=1e-10)
    x, info, rc = oracle.structured(c["gen"], c["nsamples"], c["mean"], c["a"], c["phi"], c["omega"],
                                    c["struct_mean"], c["a_type"], c["o_type"], seed0=c["seed0"])
    assert np.array_equal(info, g[name + "__info"])
    assert rel_err(x, g[name + "__x"]) < REL_TOL


def test_error_codes_and_nan(oracle):
    """zero cov -> 1; two equal rows of G -> 2 with the unprojected y left in out;
    NaN mean -> NaN out with info 0 (reference tests/test_pyhtnorm.py:46-48, 66-68)."""
    k = 10
    rng = np.random.default_rng(3)
    cov = cases.spd(rng, k)
    x, info, rc = oracle.hyperplane("pcg64", 2, np.zeros(k), np.zeros((k, k)), np.ones((2, k)), np.zeros(2))
    assert rc == 1 and info.tolist() == [1, 1]
    x, info, rc = oracle.hyperplane("pcg64", 2, np.zeros(k), cov, np.ones((2, k)), np.zeros(2))
    assert rc == 2 and info.tolist() == [2, 2] and np.isfinite(x).all()
    mu = np.zeros(k)
    mu[0] = np.nan
    x, info, rc = oracle.hyperplane("pcg64", 1, mu, cov, np.ones((1, k)), np.zeros(1))
    assert rc == 0 and np.isnan(x).all()


def test_reference_properties(oracle):
    """The properties the reference's own tests pin (tests/test_pyhtnorm.py:51-64, 97-103)."""
    c = cases.ht_case("k3_diag_40")
    xd, _, _ = oracle.hyperplane("pcg64", 4, c["mean"], c["cov"], c["g"], c["r"], diag=True, seed0=1)
    xf, _, _ = oracle.hyperplane("pcg64", 4, c["mean"], c["cov"], c["g"], c["r"], diag=False, seed0=1)
    assert np.allclose(xd, xf, rtol=0, atol=1e-12)       # diag=True == dense path on a diagonal cov
    k = 64
    rng = np.random.default_rng(5)
    cov = cases.spd(rng, k)
    x, _, _ = oracle.hyperplane("xrs128p", 8, rng.standard_normal(k), cov, np.ones((1, k)), np.zeros(1))
    assert np.abs(x.sum(axis=1)).max() < 1e-12           # G = 1^T, r = 0  =>  sum(x) = 0
    c = cases.sp_case("dd_plain")
    a = oracle.structured("pcg64", 4, c["mean"], c["a"], c["phi"], c["omega"], False, 1, 1, seed0=2)[0]
    b = oracle.structured("pcg64", 4, c["mean"], c["a"], c["phi"], c["omega"], False, 0, 0, seed0=2)[0]
    assert np.allclose(a, b, rtol=0, atol=1e-12)          # regular == diagonal types, same seed


def test_oracle_vs_reference_binary(oracle, reference):
    """Direct comparison with the compiled reference on fresh inputs (skipped on the GPU box only if
    oracle/_ref did not travel)."""
    rng = np.random.default_rng(99)
    for gen in ("pcg64", "xrs128p"):
        assert np.array_equal(oracle.raw(gen, 3, 500, seed0=4), reference.raw(gen, 3, 500, seed0=4))
        zo, no = oracle.normals(gen, 16, 4000, seed0=5)
        zr, nr = reference.normals(gen, 16, 4000, seed0=5)
        assert np.array_equal(zo.view(np.uint64), zr.view(np.uint64)) and np.array_equal(no, nr)
    k, k2 = 80, 5
    cov = cases.spd(rng, k)
    g = rng.standard_normal((k2, k))
    r = rng.standard_normal(k2)
    mu = rng.standard_normal(k)
    xo = oracle.hyperplane("pcg64", 8, mu, cov, g, r, seed0=1)[0]
    xr = reference.hyperplane("pcg64", 8, mu, cov, g, r, seed0=1)[0]
    assert rel_err(xo, xr) < 1e-12
    # the drop-in headers (include/htnorm.h) and the reference binary agree on struct layouts
    assert reference.lib.refh_sizeof_ht_config() == 56
    assert reference.lib.refh_sizeof_sp_config() == 64
    assert reference.lib.refh_sizeof_rng() == 24


def test_ziggurat_moments(oracle):
    z, _ = oracle.normals("pcg64", 4, 50000, seed0=1)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.02


def test_fullsize_golden_inputs_are_reproducible():
    """tests/golden/fullsize.npz (configs[3] / configs[4] at full size) was generated from seeded inputs that the GPU
    box rebuilds: the fingerprint stored with the golden rows must match the rebuilt arrays (cfg4 here: 0.5 GiB;
    cfg5's 2 GiB covariance is rebuilt - and fingerprinted implicitly by the parity assertion - in the GPU test)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden import digest
    gold = np.load(os.path.join(ROOT, "tests", "golden", "fullsize.npz"))
    c = cases.cfg4_inputs()
    d = digest(c["mean"], c["a"], c["phi"], c["omega"])
    assert np.allclose(d, gold["cfg4__digest"], rtol=1e-9, atol=1e-9)
    assert gold["cfg4__x"].shape == (4, 8192) and gold["cfg5__x"].shape == (2, 16384)
    assert np.isfinite(gold["cfg4__x"]).all() and np.isfinite(gold["cfg5__x"]).all()
