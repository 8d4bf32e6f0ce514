"""The reference's own callers against this library.

* The reference's Cython wrapper (pyhtnorm/_htnorm.pyx), compiled UNMODIFIED against include/ (with the
  htnorm_distributions.h shim) and linked to libhtnorm_b200.so by oracle/build_pyhtnorm.sh, runs the reference's
  own pytest file (tests/test_pyhtnorm.py:34-124, staged byte for byte into the git-ignored oracle/_ref/) UNCHANGED.
* The R binding's view (src/htnorm_r_wrapper.c:55-98, src/Makevars:1): the single-sample signatures of
  libhtnorm_b200_colmajor.so take column-major g / phi, checked against goldens of the reference's own
  -DHTNORM_COLMAJOR build (tests/golden/colmajor.npz, tools/make_golden.py).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from conftest import GOLD, REL_TOL, ROOT, rel_err

WRAP = os.path.join(ROOT, "oracle", "_ref", "pyhtnorm_b200")


def _have_wrapper():
    return os.path.isdir(os.path.join(WRAP, "pyhtnorm")) and os.path.exists(os.path.join(WRAP, "test_pyhtnorm.py"))


def _run(code_or_args, is_code=True):
    args = [sys.executable, "-c", code_or_args] if is_code else [sys.executable] + code_or_args
    return subprocess.run(args, cwd=WRAP, capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=WRAP))


def test_reference_wrapper_builds_and_imports():
    """CPU: the unmodified .pyx compiled against our headers imports and resolves every symbol it binds in
    libhtnorm_b200.so (init_*_config, htn_*_mvn); it is linked to this library and to nothing of the reference's."""
    if not _have_wrapper():
        if not os.path.isdir("/root/reference"):
            pytest.skip("reference wrapper not built (no /root/reference here)")
        subprocess.run([os.path.join(ROOT, "oracle", "build_pyhtnorm.sh")], check=True)
    out = _run("import pyhtnorm\n"
               "assert callable(pyhtnorm.hyperplane_truncated_mvnorm) and callable(pyhtnorm.structured_precision_mvnorm)\n"
               "print('ok')\n")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
    ext = next(os.path.join(WRAP, "pyhtnorm", f) for f in os.listdir(os.path.join(WRAP, "pyhtnorm"))
               if f.startswith("_htnorm") and f.endswith(".so"))
    ldd = subprocess.run(["ldd", ext], capture_output=True, text=True).stdout
    assert "libhtnorm_b200.so" in ldd and "libhtnorm_ref" not in ldd and "openblas" not in ldd.lower()


@pytest.mark.gpu
def test_reference_pytest_file_runs_unchanged():
    """GPU: `pytest test_pyhtnorm.py` - the reference's three test functions, NumPy generators through the rng_t
    callbacks (pyhtnorm/_htnorm.pyx:116-122), every sample drawn by libhtnorm_b200.so."""
    assert _have_wrapper(), "oracle/_ref/pyhtnorm_b200 missing: run oracle/build_pyhtnorm.sh where /root/reference exists"
    out = _run(["-m", "pytest", "-q", "-p", "no:cacheprovider", "test_pyhtnorm.py"], is_code=False)
    assert out.returncode == 0 and "3 passed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    # the wrapper's samples are this library's: the same NumPy seed through our ctypes mirror gives the same numbers
    code = ("import numpy as np, pyhtnorm, sys\n"
            "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import cases, htnorm_b200 as h\n"
            "c = cases.ht_case('k4_dense_64')\n"
            "a = pyhtnorm.hyperplane_truncated_mvnorm(c['mean'], c['cov'], c['g'], c['r'], random_state=7)\n"
            "b = h.hyperplane_truncated_mvnorm(c['mean'], c['cov'], c['g'], c['r'], random_state=7)\n"
            "assert np.array_equal(a, b), np.abs(a - b).max()\n"
            "s = cases.sp_case('nn_plain')\n"
            "a = pyhtnorm.structured_precision_mvnorm(s['mean'], s['a'], s['phi'], s['omega'], random_state=3)\n"
            "b = h.structured_precision_mvnorm(s['mean'], s['a'], s['phi'], s['omega'], random_state=3)\n"
            "assert np.array_equal(a, b)\n"
            "print('same')\n") % (ROOT, os.path.join(ROOT, "tests"))
    out = _run(code)
    assert out.returncode == 0 and "same" in out.stdout, out.stdout + out.stderr[-2000:]


def _load_colmajor():
    import htnorm_b200 as h
    L = h._lib
    lib = C.CDLL(os.path.join(os.path.dirname(L.LIB_PATH), "libhtnorm_b200_colmajor.so"))
    lib.rng_pcg64_new_seeded.restype = C.POINTER(L.RngT)
    lib.rng_pcg64_new_seeded.argtypes = [C.c_uint64]
    lib.rng_xrs128p_new_seeded.restype = C.POINTER(L.RngT)
    lib.rng_xrs128p_new_seeded.argtypes = [C.c_uint64]
    lib.rng_free.argtypes = [C.POINTER(L.RngT)]
    lib.rng_free.restype = None
    lib.htn_hyperplane_truncated_mvn.argtypes = [C.POINTER(L.RngT), C.POINTER(L.HtConfig), C.c_void_p]
    lib.htn_structured_precision_mvn.argtypes = [C.POINTER(L.RngT), C.POINTER(L.SpConfig), C.c_void_p]
    lib.htn_version.restype = C.c_char_p
    return lib, L


def test_colmajor_library_exports_the_same_boundary():
    import htnorm_b200 as h
    lib, L = _load_colmajor()
    for name in h.PUBLIC_SYMBOLS:
        assert hasattr(lib, name), name
    assert b"column-major" in lib.htn_version() and b"column-major" not in h.lib.htn_version()


@pytest.mark.gpu
def test_colmajor_single_sample_signature_matches_reference_colmajor_build():
    """What the R binding calls (src/htnorm_r_wrapper.c:55-67, 85-98): htn_*_mvn(rng, &config, out) with R's
    column-major matrices, on the library built with the compile-time HTNORM_COLMAJOR default."""
    lib, L = _load_colmajor()
    gold = np.load(os.path.join(GOLD, "colmajor.npz"))
    new = {"pcg64": lib.rng_pcg64_new_seeded, "xrs128p": lib.rng_xrs128p_new_seeded}
    for name in ("k4_dense_64", "k20_dense_100", "k3_diag_40", "k1_dense_50"):
        c = cases.ht_case(name)
        k2, k = c["g"].shape
        gcm = np.ascontiguousarray(c["g"].T)                 # k x k2 row-major == k2 x k column-major bytes
        want = gold["ht_" + name + "__x"]
        for i in range(3):
            rng = new[c["gen"]](c["seed0"] + i)
            conf = L.HtConfig()
            lib.init_ht_config(C.byref(conf), C.c_size_t(k2), C.c_size_t(k), C.c_void_p(c["mean"].ctypes.data),
                               C.c_void_p(c["cov"].ctypes.data), C.c_void_p(gcm.ctypes.data),
                               C.c_void_p(c["r"].ctypes.data), C.c_bool(bool(c["diag"])))
            out = np.empty(k)
            assert lib.htn_hyperplane_truncated_mvn(rng, C.byref(conf), out.ctypes.data) == 0
            lib.rng_free(rng)
            assert rel_err(out[None], want[i][None]) < REL_TOL, (name, i)
    for name in ("nn_plain", "nn_struct", "dn_struct", "in_plain"):
        c = cases.sp_case(name)
        n, p = c["phi"].shape
        pcm = np.ascontiguousarray(c["phi"].T)
        want = gold["sp_" + name + "__x"]
        for i in range(3):
            rng = new[c["gen"]](c["seed0"] + i)
            conf = L.SpConfig()
            lib.init_sp_config(C.byref(conf), C.c_size_t(n), C.c_size_t(p), C.c_void_p(c["mean"].ctypes.data),
                               C.c_void_p(c["a"].ctypes.data), C.c_void_p(pcm.ctypes.data),
                               C.c_void_p(c["omega"].ctypes.data), C.c_bool(bool(c["struct_mean"])),
                               C.c_int(c["a_type"]), C.c_int(c["o_type"]))
            out = np.empty(p)
            assert lib.htn_structured_precision_mvn(rng, C.byref(conf), out.ctypes.data) == 0
            lib.rng_free(rng)
            assert rel_err(out[None], want[i][None]) < REL_TOL, (name, i)
