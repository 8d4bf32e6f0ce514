"""Python mirror of the reference's user-facing interface, over the C ABI.

Names, argument meaning and error behaviour follow the reference's wrapper:

* ``hyperplane_truncated_mvnorm`` / ``structured_precision_mvnorm``  <-> pyhtnorm/_htnorm.pyx:125-210, 213-308
* ``HTNGenerator``  <-> R/htnorm.r:132-181 (object holding one of the library's own generators)

plus the batched calls that are new here (``*_batch``): many samples per call, host arrays (NumPy) or
device arrays (torch CUDA tensors: pointers are passed straight through, work is enqueued on the
current torch stream).  torch is plumbing only: allocation, streams, ``torch.distributed``.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GEN, TYPES, Batch, HtConfig, SpConfig, check, lib

_SHAPE_ERR = "`cov` and `g` have incompatible shapes" \
    # reference message: pyhtnorm/_htnorm.pyx:193-195


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np_in(x, name, ndim):
    a = np.asarray(x)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    if not a.flags.c_contiguous:
        # the reference rejects non C-contiguous input (pyhtnorm/_htnorm.pyx:126-129, test :37-38)
        raise ValueError(f"`{name}` must be C-contiguous")
    if a.ndim != ndim:
        raise ValueError(f"`{name}` must have {ndim} dimension(s)")
    return a


class HTNGenerator:
    """One of the library's generators ("pcg64" = 2 x PCG32 as in the reference, or "xrs128p"),
    advanced on the GPU by every draw (R/htnorm.r:132-181; src/htnorm_r_wrapper.c:26-45)."""

    def __init__(self, seed=None, gen="pcg64"):
        if gen not in ("pcg64", "xrs128p"):
            raise ValueError("`gen` must be 'pcg64' or 'xrs128p'")
        if seed is not None:
            if isinstance(seed, bool) or not isinstance(seed, (int, np.integer)) or seed < 0:
                raise ValueError("`seed` must be a non-negative integer or None")
        self.gen = gen
        if seed is None:
            self._rng = lib.rng_pcg64_new() if gen == "pcg64" else lib.rng_xrs128p_new()
        elif gen == "pcg64":
            self._rng = lib.rng_pcg64_new_seeded(C.c_uint64(int(seed) & (2**64 - 1)))
        else:
            self._rng = lib.rng_xrs128p_new_seeded(C.c_uint64(int(seed) & (2**64 - 1)))
        if not self._rng:
            raise MemoryError("could not allocate a generator")

    def __del__(self):
        rng = getattr(self, "_rng", None)
        if rng:
            lib.rng_free(rng)
            self._rng = None

    @property
    def state(self):
        nwords = 4 if self.gen == "pcg64" else 2
        return list((C.c_uint64 * nwords).from_address(self._rng.contents.base))

    def hyperplane_truncated_mvnorm(self, mean, cov, g, r, diag=False, out=None):
        mean, cov = _np_in(mean, "mean", 1), _np_in(cov, "cov", 2)
        g, r = _np_in(g, "g", 2), _np_in(r, "r", 1)
        k2, k = g.shape
        if cov.shape != (k, k) or mean.shape[0] != k:
            raise ValueError(_SHAPE_ERR)
        if r.shape[0] != k2:
            raise ValueError("`r` and `g` have incompatible shapes")
        if out is None:
            out = np.empty(k)
        elif out.shape != (k,) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("`out` must be a C-contiguous float64 array of the same length as `mean`")
        conf = HtConfig()
        lib.init_ht_config(C.byref(conf), C.c_size_t(k2), C.c_size_t(k), C.c_void_p(mean.ctypes.data),
                           C.c_void_p(cov.ctypes.data), C.c_void_p(g.ctypes.data),
                           C.c_void_p(r.ctypes.data), C.c_bool(bool(diag)))
        check(lib.htn_hyperplane_truncated_mvn(self._rng, C.byref(conf), out.ctypes.data),
              "hyperplane_truncated_mvnorm")
        return out

    def structured_precision_mvnorm(self, mean, a, phi, omega, mean_structured=False, a_type=0, o_type=0,
                                    out=None):
        if a_type not in (0, 1, 2) or o_type not in (0, 1, 2):
            raise ValueError("`a_type` and `o_type` must be one of {0, 1, 2}")
        mean, a = _np_in(mean, "mean", 1), _np_in(a, "a", 2)
        phi, omega = _np_in(phi, "phi", 2), _np_in(omega, "omega", 2)
        n, p = phi.shape
        if a.shape != (p, p) or omega.shape != (n, n):
            raise ValueError("`omega` and `a` both need to be square matrices matching `phi`")
        if mean.shape[0] != (n if mean_structured else p):
            raise ValueError("`mean` has the wrong length")
        if out is None:
            out = np.empty(p)  # always p (the reference sizes it by len(mean))
        elif out.shape != (p,) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("`out` must be a C-contiguous float64 array of length phi.shape[1]")
        conf = SpConfig()
        lib.init_sp_config(C.byref(conf), C.c_size_t(n), C.c_size_t(p), C.c_void_p(mean.ctypes.data),
                           C.c_void_p(a.ctypes.data), C.c_void_p(phi.ctypes.data),
                           C.c_void_p(omega.ctypes.data), C.c_bool(bool(mean_structured)), C.c_int(a_type),
                           C.c_int(o_type))
        check(lib.htn_structured_precision_mvn(self._rng, C.byref(conf), out.ctypes.data),
              "structured_precision_mvnorm")
        return out


class _NumpyGenerator(HTNGenerator):
    """A NumPy bit generator behind the C `rng_t` vtable, exactly what the reference's wrapper builds
    (pyhtnorm/_htnorm.pyx:116-122: base = bitgen.state, next_uint64 / next_double = NumPy's functions).
    Such a FOREIGN generator cannot run on the device: its normals are drawn on the host through the
    callbacks (csrc/host/foreign_rng.c), the algebra runs on the GPU."""

    def __init__(self, random_state):
        gen = np.random.default_rng(random_state)   # same rule as the reference (_htnorm.pyx:201)
        self.gen = "numpy"
        self._np = gen
        self._bitgen = gen.bit_generator
        iface = self._bitgen.ctypes
        self._rng_struct = _lib.RngT(C.c_void_p(iface.state_address).value,
                                     C.cast(iface.next_uint64, C.c_void_p).value,
                                     C.cast(iface.next_double, C.c_void_p).value)
        self._rng = C.pointer(self._rng_struct)

    def __del__(self):
        self._rng = None

    def _locked(self, fn, *a, **kw):
        with self._bitgen.lock:                     # the reference holds the same lock (_htnorm.pyx:205)
            return fn(*a, **kw)

    def hyperplane_truncated_mvnorm(self, *a, **kw):
        return self._locked(HTNGenerator.hyperplane_truncated_mvnorm, self, *a, **kw)

    def structured_precision_mvnorm(self, *a, **kw):
        return self._locked(HTNGenerator.structured_precision_mvnorm, self, *a, **kw)


def _generator(random_state):
    """random_state as in the reference: anything numpy.random.default_rng accepts (None, int, SeedSequence,
    BitGenerator, Generator) -> NumPy's generator driven through the rng_t callbacks; an HTNGenerator -> one
    of the library's own generators, advanced on the GPU."""
    if isinstance(random_state, HTNGenerator):
        return random_state
    return _NumpyGenerator(random_state)


def hyperplane_truncated_mvnorm(mean, cov, g, r, diag=False, out=None, random_state=None):
    """One draw from N(mean, cov) truncated on {x : g x = r} (pyhtnorm/_htnorm.pyx:125-210)."""
    return _generator(random_state).hyperplane_truncated_mvnorm(mean, cov, g, r, diag, out)


def structured_precision_mvnorm(mean, a, phi, omega, mean_structured=False, a_type=0, o_type=0, out=None,
                                random_state=None):
    """One draw from N(mean, (a + phi^T omega phi)^-1) (pyhtnorm/_htnorm.pyx:213-308)."""
    return _generator(random_state).structured_precision_mvnorm(mean, a, phi, omega, mean_structured, a_type,
                                                                o_type, out)


# ---------------------------------------------------------------- batched (new) ----------------


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def _prep(arrays, dtype=np.float64):
    """-> (device_mode, arrays): all NumPy (host) or all torch CUDA tensors on one device."""
    torch_mode = any(_is_torch(a) for a in arrays if a is not None)
    outv = []
    if torch_mode:
        import torch
        tdtype = {np.float64: torch.float64, np.uint64: torch.int64}[dtype]
        dev = None
        for a in arrays:
            if a is None:
                outv.append(None)
                continue
            if not _is_torch(a) or not a.is_cuda:
                raise TypeError("mixing host arrays and CUDA tensors in one call is not supported")
            if a.dtype != tdtype and not (dtype is np.uint64 and a.dtype in (torch.int64, torch.uint64)):
                raise TypeError(f"CUDA tensors must be {tdtype}")
            if not a.is_contiguous():
                raise ValueError("CUDA tensors must be contiguous")
            dev = a.device.index if dev is None else dev
            if a.device.index != dev:
                raise ValueError("all tensors must live on one device")
            outv.append(a)
        return dev, outv
    for a in arrays:
        if a is None:
            outv.append(None)
            continue
        a = np.asarray(a)
        if a.dtype != dtype:
            a = a.astype(dtype)
        outv.append(np.ascontiguousarray(a))
    return None, outv


def _batch(nsamples, gen, seeds, seed0, independent, device, flags, stream=None):
    b = Batch()
    lib.htn_batch_init(C.byref(b), C.c_size_t(nsamples))
    b.gen = GEN[gen]
    b.seed0 = int(seed0) & (2**64 - 1)
    b.independent = int(bool(independent))
    b.flags = flags
    if device is not None:
        import torch
        b.mem = _lib.HTN_MEM_DEVICE
        b.device = device
        b.stream = stream if stream is not None else torch.cuda.current_stream(device).cuda_stream
    if seeds is not None:
        b.seeds = _ptr(seeds)
    return b


def _seeds_arg(seeds, nsamples, device):
    if seeds is None:
        return None
    if _is_torch(seeds):
        import torch
        if device is None:
            raise TypeError("`seeds` is a CUDA tensor but the inputs are host arrays")
        if not seeds.is_cuda or seeds.device.index != device:
            raise ValueError("`seeds` must live on the same CUDA device as the inputs")
        if seeds.dtype not in (torch.int64, torch.uint64):
            raise TypeError("`seeds` must be a 64-bit integer tensor (int64 or uint64)")
        if seeds.numel() != nsamples:
            raise ValueError("`seeds` must have one entry per sample")
        return seeds.contiguous()
    s = np.ascontiguousarray(seeds, dtype=np.uint64)
    if s.shape != (nsamples,):
        raise ValueError("`seeds` must have one entry per sample")
    if device is not None:
        import torch
        return torch.from_numpy(s.view(np.int64)).to(f"cuda:{device}")
    return s


def _alloc(shape, device, dtype="f8"):
    """Result buffers the wrapper allocates itself.  Device mode: `out` starts as NaN and `info` as 0, so rows a
    failed factorisation leaves untouched are recognisable (the device path reports failures through `info` only)."""
    if device is None:
        return np.empty(shape, dtype=dtype)
    import torch
    if dtype == "f8":
        return torch.full(shape, float("nan"), dtype=torch.float64, device=f"cuda:{device}")
    return torch.zeros(shape, dtype=torch.int32, device=f"cuda:{device}")


def _check_result(x, name, shape, device, dtype):
    """A caller-supplied `out` / `info` goes to the C ABI as a raw pointer: shape, dtype, contiguity and memory
    space (host arrays with host inputs, CUDA tensors on the same device with device inputs) are checked here."""
    want = {"f8": "float64", "i4": "int32"}[dtype]
    if device is None:
        if _is_torch(x):
            raise TypeError(f"`{name}` is a torch tensor but the inputs are host arrays")
        if not isinstance(x, np.ndarray):
            raise TypeError(f"`{name}` must be a NumPy array")
        if x.dtype != np.dtype(dtype) or tuple(x.shape) != tuple(shape) or not x.flags.c_contiguous:
            raise ValueError(f"`{name}` must be a C-contiguous {want} array of shape {tuple(shape)}")
        if not x.flags.writeable:
            raise ValueError(f"`{name}` must be writeable")
        return x
    import torch
    if not _is_torch(x) or not x.is_cuda:
        raise TypeError(f"`{name}` must be a CUDA tensor when the inputs are CUDA tensors")
    if x.device.index != device:
        raise ValueError(f"`{name}` must live on the same CUDA device as the inputs")
    if x.dtype != {"f8": torch.float64, "i4": torch.int32}[dtype]:
        raise TypeError(f"`{name}` must be a {want} tensor")
    if tuple(x.shape) != tuple(shape) or not x.is_contiguous():
        raise ValueError(f"`{name}` must be a contiguous tensor of shape {tuple(shape)}")
    return x


def _finish(rc, info, device, raise_on_error, what):
    """Return-code policy.  Host memory: the call is synchronous and returns the first non-zero per-sample code
    (raised as the reference's ValueError unless raise_on_error is False).  Device memory: the call returns 0 once
    enqueued and NEVER waits for the GPU by itself; the codes are in `info`.  With raise_on_error=True the wrapper
    synchronises once and raises on the first non-zero code; the default (None) leaves the device path asynchronous."""
    if raise_on_error is None:
        raise_on_error = device is None
    if rc < 0 or (rc > 0 and raise_on_error):
        check(rc, what)
    if device is not None and raise_on_error and info.numel():
        bad = info[info != 0]               # synchronises
        if bad.numel():
            check(int(bad[0].item()), what)


def hyperplane_truncated_mvnorm_batch(nsamples, mean, cov, g, r, diag=False, *, gen="pcg64", seeds=None,
                                      seed0=0, independent=False, out=None, info=None, flags=0, stream=None,
                                      raise_on_error=None):
    """``nsamples`` draws in one call.  Sample i == the reference's single draw with a generator freshly
    seeded with ``seeds[i]`` (``seed0 + i`` when seeds is None).  ``independent``: inputs carry one problem
    per sample (mean: B x k, cov: B x k x k, g: B x k2 x k, r: B x k2).  Returns ``(out, info)``.
    ``raise_on_error``: None = raise for host arrays, stay asynchronous for CUDA tensors (see ``_finish``)."""
    device, (mean, cov, g, r) = _prep([mean, cov, g, r])
    k2, k = int(g.shape[-2]), int(g.shape[-1])
    lead = (nsamples,) if independent else ()
    if tuple(cov.shape) != lead + (k, k) or tuple(mean.shape) != lead + (k,):
        raise ValueError(_SHAPE_ERR)
    if tuple(r.shape) != lead + (k2,) or tuple(g.shape) != lead + (k2, k):
        raise ValueError("`r` and `g` have incompatible shapes")
    seeds = _seeds_arg(seeds, nsamples, device)
    out = _alloc((nsamples, k), device) if out is None else _check_result(out, "out", (nsamples, k), device, "f8")
    info = _alloc((nsamples,), device, "i4") if info is None else _check_result(info, "info", (nsamples,), device, "i4")
    conf = HtConfig()
    lib.init_ht_config(C.byref(conf), C.c_size_t(k2), C.c_size_t(k), _ptr(mean), _ptr(cov), _ptr(g), _ptr(r),
                       C.c_bool(bool(diag)))
    b = _batch(nsamples, gen, seeds, seed0, independent, device, flags, stream)
    rc = lib.htn_hyperplane_truncated_mvn_batch(C.byref(b), C.byref(conf), _ptr(out), _ptr(info))
    _finish(rc, info, device, raise_on_error, "hyperplane_truncated_mvnorm_batch")
    return out, info


def structured_precision_mvnorm_batch(nsamples, mean, a, phi, omega, mean_structured=False, a_type=0,
                                      o_type=0, *, gen="pcg64", seeds=None, seed0=0, out=None, info=None,
                                      flags=0, stream=None, raise_on_error=None):
    """``nsamples`` draws from N(mean, (a + phi^T omega phi)^-1) sharing a, phi, omega."""
    device, (mean, a, phi, omega) = _prep([mean, a, phi, omega])
    n, p = int(phi.shape[0]), int(phi.shape[1])
    a_type, o_type = TYPES[a_type], TYPES[o_type]
    if tuple(a.shape) != (p, p) or tuple(omega.shape) != (n, n):
        raise ValueError("`omega` and `a` both need to be square matrices matching `phi`")
    if tuple(mean.shape) != ((n,) if mean_structured else (p,)):
        raise ValueError("`mean` has the wrong length")
    seeds = _seeds_arg(seeds, nsamples, device)
    out = _alloc((nsamples, p), device) if out is None else _check_result(out, "out", (nsamples, p), device, "f8")
    info = _alloc((nsamples,), device, "i4") if info is None else _check_result(info, "info", (nsamples,), device, "i4")
    conf = SpConfig()
    lib.init_sp_config(C.byref(conf), C.c_size_t(n), C.c_size_t(p), _ptr(mean), _ptr(a), _ptr(phi), _ptr(omega),
                       C.c_bool(bool(mean_structured)), C.c_int(a_type), C.c_int(o_type))
    b = _batch(nsamples, gen, seeds, seed0, False, device, flags, stream)
    rc = lib.htn_structured_precision_mvn_batch(C.byref(b), C.byref(conf), _ptr(out), _ptr(info))
    _finish(rc, info, device, raise_on_error, "structured_precision_mvnorm_batch")
    return out, info


def raw_streams(nstreams, nwords, gen="pcg64", seeds=None, seed0=0):
    """uint64[nstreams, nwords]: the first words of each stream, generated on the GPU."""
    seeds = _seeds_arg(seeds, nstreams, None)
    out = np.empty((nstreams, nwords), dtype=np.uint64)
    b = _batch(nstreams, gen, seeds, seed0, False, None, 0)
    check(lib.htn_rng_raw_fill(C.byref(b), C.c_size_t(nwords), _ptr(out)), "raw_streams")
    return out


def standard_normals(nstreams, n, gen="pcg64", seeds=None, seed0=0):
    """(float64[nstreams, n] in the reference's reverse fill order, uint32[nstreams] generator calls)."""
    seeds = _seeds_arg(seeds, nstreams, None)
    out = np.empty((nstreams, n), dtype=np.float64)
    nd = np.empty(nstreams, dtype=np.uint32)
    b = _batch(nstreams, gen, seeds, seed0, False, None, 0)
    check(lib.htn_std_normal_fill(C.byref(b), C.c_size_t(n), _ptr(out), _ptr(nd)), "standard_normals")
    return out, nd


def launch_count():
    return int(lib.htn_launch_count())


def fp64_tensor_peak_tflops(device=-1, ms=50.0):
    return float(lib.htn_fp64_tensor_peak_tflops(device, ms))


def fp64_fma_peak_tflops(device=-1, ms=50.0):
    return float(lib.htn_fp64_fma_peak_tflops(device, ms))


def int_issue_peak_gops(device=-1, ms=50.0):
    """integer issue ceiling (giga warp-instructions / s) with the generators' instruction mix"""
    return float(lib.htn_int_issue_peak_gops(device, ms))


def phase_timing(enable=None):
    """enable/disable the CUDA-event phase marks, or (enable=None) read the per-phase AVERAGES in ms over the
    calls made since it was enabled: dict(setup, normals, coeff, gemm, calls)."""
    if enable is not None:
        lib.htn_phase_timing_enable(int(bool(enable)))
        return None
    ms = (C.c_double * 8)()
    ncalls = lib.htn_phase_timing_get(ms, 8)
    return dict(setup=ms[0], normals=ms[1], coeff=ms[2], gemm=ms[3], calls=ncalls)
