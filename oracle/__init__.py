"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle and to the compiled reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; ``htnorm_b200`` never does.

* ``Oracle``   - ``oracle/_build/liboracle.so``: plain-C restatement (``htnorm_oracle.c``).
* ``Reference`` - ``oracle/_ref/libref_harness.so`` -> ``libhtnorm_ref.so``: the UNMODIFIED reference
  compiled from ``/root/reference/src`` in place (``oracle/build_ref.sh``), linked to scipy's OpenBLAS.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GEN = {"pcg64": 0, "xrs128p": 1, 0: 0, 1: 1}
TYPES = {"normal": 0, "regular": 0, "diagonal": 1, "identity": 2, 0: 0, 1: 1, 2: 2}

_u64p = C.POINTER(C.c_uint64)
_dblp = C.POINTER(C.c_double)
_intp = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint32)


def build():
    """(Re)build the oracle; builds oracle/_ref too when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _seeds(seeds):
    return None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)


class Oracle:
    """Plain-C CPU restatement (kind "port")."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_raw_fill.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_size_t, _u64p]
        L.orc_raw_fill.restype = None
        L.orc_normal_fill.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_size_t, _dblp, _u32p]
        L.orc_normal_fill.restype = None
        L.orc_hyperplane_batch.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_int, C.c_int,
                                           C.c_int, _dblp, _dblp, _dblp, _dblp, C.c_int, _dblp, _intp]
        L.orc_hyperplane_batch.restype = C.c_int
        L.orc_structured_batch.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_int, C.c_int,
                                           _dblp, _dblp, _dblp, _dblp, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _dblp, _intp]
        L.orc_structured_batch.restype = C.c_int
        L.orc_potrf.argtypes = [C.c_int, _dblp, _dblp]
        L.orc_potrf.restype = C.c_int

    def raw(self, gen, nstreams, nwords, seeds=None, seed0=0):
        s = _seeds(seeds)
        out = np.empty((nstreams, nwords), dtype=np.uint64)
        self.lib.orc_raw_fill(GEN[gen], _ptr(s, _u64p), seed0, nstreams, nwords, _ptr(out, _u64p))
        return out

    def normals(self, gen, nstreams, n, seeds=None, seed0=0):
        s = _seeds(seeds)
        out = np.empty((nstreams, n), dtype=np.float64)
        nd = np.empty(nstreams, dtype=np.uint32)
        self.lib.orc_normal_fill(GEN[gen], _ptr(s, _u64p), seed0, nstreams, n, _ptr(out, _dblp),
                                 _ptr(nd, _u32p))
        return out, nd

    def hyperplane(self, gen, nsamples, mean, cov, g, r, diag=False, independent=False, seeds=None,
                   seed0=0):
        mean, cov, g, r = _f64(mean), _f64(cov), _f64(g), _f64(r)
        k2, k = g.shape[-2], g.shape[-1]
        s = _seeds(seeds)
        out = np.full((nsamples, k), np.nan)
        info = np.zeros(nsamples, dtype=np.int32)
        rc = self.lib.orc_hyperplane_batch(GEN[gen], _ptr(s, _u64p), seed0, nsamples,
                                           int(independent), k2, k, _ptr(mean, _dblp),
                                           _ptr(cov, _dblp), _ptr(g, _dblp), _ptr(r, _dblp),
                                           int(diag), _ptr(out, _dblp), _ptr(info, _intp))
        return out, info, rc

    def structured(self, gen, nsamples, mean, a, phi, omega, struct_mean=False, a_type=0, o_type=0,
                   paper_sign=False, seeds=None, seed0=0):
        mean, a, phi, omega = _f64(mean), _f64(a), _f64(phi), _f64(omega)
        n, p = phi.shape
        s = _seeds(seeds)
        out = np.full((nsamples, p), np.nan)
        info = np.zeros(nsamples, dtype=np.int32)
        rc = self.lib.orc_structured_batch(GEN[gen], _ptr(s, _u64p), seed0, nsamples, n, p,
                                           _ptr(mean, _dblp), _ptr(a, _dblp), _ptr(phi, _dblp),
                                           _ptr(omega, _dblp), int(struct_mean), TYPES[a_type],
                                           TYPES[o_type], int(paper_sign), _ptr(out, _dblp),
                                           _ptr(info, _intp))
        return out, info, rc

    def potrf(self, a):
        a = _f64(a)
        n = a.shape[0]
        l = np.empty((n, n))
        info = self.lib.orc_potrf(n, _ptr(a, _dblp), _ptr(l, _dblp))
        return l, info


class Reference:
    """The unmodified reference library (kind "reference"); raises FileNotFoundError if not built."""

    def __init__(self, colmajor=False):
        """colmajor=True: the reference compiled with -DHTNORM_COLMAJOR (what its R binding builds,
        src/Makevars:1): g / phi are column-major, symmetric inputs are read through the lower triangle."""
        path = os.path.join(HERE, "_ref", "libref_harness_colmajor.so" if colmajor else "libref_harness.so")
        if not os.path.exists(path):
            if os.path.isdir(os.environ.get("HTNORM_REFERENCE", "/root/reference")):
                build()
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.refh_raw_fill.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_size_t, _u64p]
        L.refh_raw_fill.restype = None
        L.refh_normal_fill.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_size_t, _dblp, _u32p]
        L.refh_normal_fill.restype = C.c_int
        L.refh_hyperplane_batch.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_int,
                                            C.c_size_t, C.c_size_t, _dblp, _dblp, _dblp, _dblp,
                                            C.c_int, _dblp, _intp, C.c_int, _dblp]
        L.refh_hyperplane_batch.restype = C.c_int
        L.refh_structured_batch.argtypes = [C.c_int, _u64p, C.c_uint64, C.c_size_t, C.c_size_t,
                                            C.c_size_t, _dblp, _dblp, _dblp, _dblp, C.c_int, C.c_int,
                                            C.c_int, _dblp, _intp, C.c_int, _dblp]
        L.refh_structured_batch.restype = C.c_int
        for f in ("refh_sizeof_ht_config", "refh_sizeof_sp_config", "refh_sizeof_rng"):
            getattr(L, f).restype = C.c_size_t
        self.last_seconds = None

    def raw(self, gen, nstreams, nwords, seeds=None, seed0=0):
        s = _seeds(seeds)
        out = np.empty((nstreams, nwords), dtype=np.uint64)
        self.lib.refh_raw_fill(GEN[gen], _ptr(s, _u64p), seed0, nstreams, nwords, _ptr(out, _u64p))
        return out

    def normals(self, gen, nstreams, n, seeds=None, seed0=0):
        s = _seeds(seeds)
        out = np.empty((nstreams, n), dtype=np.float64)
        nd = np.empty(nstreams, dtype=np.uint32)
        rc = self.lib.refh_normal_fill(GEN[gen], _ptr(s, _u64p), seed0, nstreams, n,
                                       _ptr(out, _dblp), _ptr(nd, _u32p))
        assert rc == 0
        return out, nd

    def hyperplane(self, gen, nsamples, mean, cov, g, r, diag=False, independent=False, seeds=None,
                   seed0=0, nthreads=1):
        mean, cov, g, r = _f64(mean), _f64(cov), _f64(g), _f64(r)
        k2, k = g.shape[-2], g.shape[-1]
        s = _seeds(seeds)
        out = np.full((nsamples, k), np.nan)
        info = np.zeros(nsamples, dtype=np.int32)
        sec = C.c_double(0)
        rc = self.lib.refh_hyperplane_batch(GEN[gen], _ptr(s, _u64p), seed0, nsamples,
                                            int(independent), k2, k, _ptr(mean, _dblp),
                                            _ptr(cov, _dblp), _ptr(g, _dblp), _ptr(r, _dblp),
                                            int(diag), _ptr(out, _dblp), _ptr(info, _intp), nthreads,
                                            C.byref(sec))
        self.last_seconds = sec.value
        first = int(info[np.nonzero(info)[0][0]]) if info.any() else 0
        return out, info, rc or first

    def structured(self, gen, nsamples, mean, a, phi, omega, struct_mean=False, a_type=0, o_type=0,
                   seeds=None, seed0=0, nthreads=1):
        mean, a, phi, omega = _f64(mean), _f64(a), _f64(phi), _f64(omega)
        n, p = phi.shape
        s = _seeds(seeds)
        out = np.full((nsamples, p), np.nan)
        info = np.zeros(nsamples, dtype=np.int32)
        sec = C.c_double(0)
        rc = self.lib.refh_structured_batch(GEN[gen], _ptr(s, _u64p), seed0, nsamples, n, p,
                                            _ptr(mean, _dblp), _ptr(a, _dblp), _ptr(phi, _dblp),
                                            _ptr(omega, _dblp), int(struct_mean), TYPES[a_type],
                                            TYPES[o_type], _ptr(out, _dblp), _ptr(info, _intp),
                                            nthreads, C.byref(sec))
        self.last_seconds = sec.value
        first = int(info[np.nonzero(info)[0][0]]) if info.any() else 0
        return out, info, rc or first


def have_reference():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_harness.so"))
