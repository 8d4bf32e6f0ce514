"""Seeded input builders shared by the golden-vector generator (tools/make_golden.py), the CPU tests
and the GPU parity tests.  Inputs follow SURVEY.md section 8(d): well-conditioned SPD matrices
(T T^T / d + I, kappa ~ 5-6) so that 1e-10 parity is attainable, plus the reference's own fixtures
(tests/test_pyhtnorm.py:7-31) and its README example (README.md:120-125)."""
import numpy as np


def spd(rng, d):
    t = rng.standard_normal((d, d))
    s = t @ t.T / d + np.eye(d)
    return (s + s.T) / 2


def ht_case(name):
    """-> dict(mean, cov, g, r, diag, gen, seed0, nsamples)"""
    if name == "pytest_fixture":  # reference tests/test_pyhtnorm.py:7-18, generator PCG64(10)
        rng = np.random.default_rng(0)
        k1, k2 = 100, 20
        temp = rng.random((k1, k1))
        cov = temp @ temp.T + np.diag(rng.random(k1))
        g = rng.random((k2, k1))
        r = rng.random(k2)
        mean = rng.random(k1)
        return dict(mean=mean, cov=cov, g=g, r=r, diag=False, gen="pcg64", seed0=10, nsamples=8)
    if name == "readme_k1000":  # reference README.md:120-125 (BASELINE.json configs[0]), PCG64(12345)
        rng = np.random.default_rng(1)
        k = 1000
        t = rng.random((k, k))
        cov = t @ t.T
        g = np.ones((1, k))
        r = np.zeros(1)
        mean = rng.random(k)
        return dict(mean=mean, cov=cov, g=g, r=r, diag=False, gen="pcg64", seed0=12345, nsamples=4)
    spec = {
        # name: (k2, k, diag, gen, seed0, nsamples, rng_seed)
        "k1_dense_50": (1, 50, False, "pcg64", 100, 16, 11),
        "k4_dense_64": (4, 64, False, "pcg64", 12345, 32, 12),
        "k4_dense_64_xrs": (4, 64, False, "xrs128p", 12345, 32, 12),
        "k20_dense_100": (20, 100, False, "xrs128p", 3, 16, 13),
        "k3_diag_40": (3, 40, True, "pcg64", 5, 16, 14),
        "k1_diag_40": (1, 40, True, "xrs128p", 6, 16, 15),
        "k7_dense_333": (7, 333, False, "pcg64", 77, 8, 16),
        "k1_dense_1000_spd": (1, 1000, False, "pcg64", 12345, 8, 17),
        "k64_dense_1024": (64, 1024, False, "xrs128p", 9, 4, 18),
    }[name]
    k2, k, diag, gen, seed0, nsamples, rs = spec
    rng = np.random.default_rng(rs)
    cov = spd(rng, k)
    if diag:
        cov = np.diag(np.diag(cov))
    g = rng.standard_normal((k2, k))
    r = rng.standard_normal(k2)
    mean = rng.standard_normal(k)
    return dict(mean=mean, cov=cov, g=g, r=r, diag=diag, gen=gen, seed0=seed0, nsamples=nsamples)


HT_CASES = ["pytest_fixture", "readme_k1000", "k1_dense_50", "k4_dense_64", "k4_dense_64_xrs",
            "k20_dense_100", "k3_diag_40", "k1_diag_40", "k7_dense_333", "k1_dense_1000_spd",
            "k64_dense_1024"]


def ht_independent_case(nproblems=48, k=64, k2=4, rng_seed=2):
    """BASELINE.json configs[1] in miniature: independent problems, each with its own inputs."""
    rng = np.random.default_rng(rng_seed)
    cov = np.stack([spd(rng, k) for _ in range(nproblems)])
    g = rng.standard_normal((nproblems, k2, k))
    r = rng.standard_normal((nproblems, k2))
    mean = rng.standard_normal((nproblems, k))
    return dict(mean=mean, cov=cov, g=g, r=r, diag=False, gen="pcg64", seed0=12345,
                nsamples=nproblems)


def sp_case(name):
    """-> dict(mean, a, phi, omega, struct_mean, a_type, o_type, gen, seed0, nsamples)"""
    if name == "pytest_fixture":  # reference tests/test_pyhtnorm.py:21-31
        rng = np.random.default_rng(0)
        n = 50
        temp = rng.random((n, n))
        cov = temp @ temp.T + np.diag(rng.random(n))
        omega, phi = np.linalg.eigh(cov)
        omega = np.diag(omega)
        a = np.diag(rng.random(n) + 1)
        mean = rng.random(n)
        return dict(mean=mean, a=a, phi=np.ascontiguousarray(phi), omega=omega, struct_mean=False,
                    a_type=1, o_type=1, gen="pcg64", seed0=10, nsamples=8)
    spec = {
        # name: (n, p, a_type, o_type, struct_mean, gen, seed0, nsamples, rng_seed)
        "nn_plain": (5, 20, 0, 0, False, "pcg64", 5, 16, 21),
        "nn_struct": (5, 20, 0, 0, True, "xrs128p", 5, 16, 22),
        "dd_plain": (5, 20, 1, 1, False, "pcg64", 6, 16, 23),
        "id_struct": (5, 20, 2, 1, True, "pcg64", 7, 16, 24),
        "nd_plain": (6, 30, 0, 1, False, "xrs128p", 8, 16, 25),
        "dn_struct": (6, 30, 1, 0, True, "pcg64", 9, 16, 26),
        "ii_plain": (6, 30, 2, 2, False, "pcg64", 10, 16, 27),
        "ni_struct": (6, 30, 0, 2, True, "xrs128p", 11, 16, 28),
        "in_plain": (7, 33, 2, 0, False, "pcg64", 12, 16, 29),
        "nd_plain_mid": (50, 200, 0, 1, False, "pcg64", 12345, 8, 30),      # cfg3 in miniature
        "nn_struct_mid": (64, 300, 0, 0, True, "pcg64", 12345, 8, 31),     # cfg4 in miniature
        "nd_plain_big": (150, 700, 0, 1, False, "xrs128p", 1, 4, 32),
    }[name]
    n, p, a_type, o_type, sm, gen, seed0, nsamples, rs = spec
    rng = np.random.default_rng(rs)
    a = spd(rng, p)
    omega = spd(rng, n)
    if a_type != 0:
        a = np.diag(np.diag(a))
    if o_type != 0:
        omega = np.diag(np.diag(omega))
    phi = rng.standard_normal((n, p)) / np.sqrt(p)
    mean = rng.standard_normal(n if sm else p)
    return dict(mean=mean, a=a, phi=phi, omega=omega, struct_mean=sm, a_type=a_type, o_type=o_type,
                gen=gen, seed0=seed0, nsamples=nsamples)


SP_CASES = ["pytest_fixture", "nn_plain", "nn_struct", "dd_plain", "id_struct", "nd_plain",
            "dn_struct", "ii_plain", "ni_struct", "in_plain", "nd_plain_mid", "nn_struct_mid",
            "nd_plain_big"]

RNG_SEEDS = [0, 1, 10, 12345, 2**32 - 1, 2**63 + 5, 2**64 - 1]


# ---- BASELINE.json configs at FULL size (configs[3], configs[4]) ------------------------------------------------
# Inputs must be rebuilt identically on the GPU box from a seed, cheaply (no O(d^3) BLAS on the host): SPD matrices of
# the form  diag(U(1,2)) + V V^T / d  with V d x 64 - dense, kappa ~ 3, 2 * 64 * d^2 flops to build.  The golden rows
# come from the unmodified reference on the same arrays (tools/make_golden_fullsize.py -> tests/golden/fullsize.npz).

def spd_lowrank(rng, d, rank=64):
    v = rng.standard_normal((d, rank))
    s = v @ v.T
    s /= d
    s[np.diag_indices(d)] += rng.uniform(1.0, 2.0, d)
    return s


def cfg4_inputs():
    """configs[3]: structured precision + structured mean, p = 8192, n = 1024, dense A and Omega."""
    rng = np.random.default_rng(404)
    p, n = 8192, 1024
    a = spd_lowrank(rng, p)
    omega = spd_lowrank(rng, n)
    phi = rng.standard_normal((n, p)) / np.sqrt(p)
    t = rng.standard_normal(n)
    return dict(mean=t, a=a, phi=phi, omega=omega, struct_mean=True, a_type=0, o_type=0, gen="pcg64", seed0=12345,
                spot=np.array([0, 1, 128, 255]))


def cfg5_inputs():
    """configs[4], one GPU's shard: hyperplane-truncated MVN k = 16384, k2 = 64.  G is scaled by 1/sqrt(k) so that
    G x is O(1) and the constraint residual can be evaluated to ~1e-16 in float64."""
    rng = np.random.default_rng(505)
    k, k2 = 16384, 64
    cov = spd_lowrank(rng, k)
    g = rng.standard_normal((k2, k)) / np.sqrt(k)
    r = rng.standard_normal(k2)
    mean = rng.standard_normal(k)
    return dict(mean=mean, cov=cov, g=g, r=r, gen="pcg64", seed0=12345, spot=np.array([0, 8191]))
