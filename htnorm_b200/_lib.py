"""ctypes binding of the C-ABI library ``libhtnorm_b200.so`` (include/htnorm.h, htnorm_rng.h,
htnorm_batch.h).  There is no fallback of any kind: if the CUDA library is missing the import fails."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhtnorm_b200.so")

# every symbol the public headers declare; tests check the list against include/*.h
PUBLIC_SYMBOLS = [
    "init_ht_config", "init_sp_config", "htn_hyperplane_truncated_mvn", "htn_structured_precision_mvn",
    "rng_free", "rng_pcg64_new", "rng_pcg64_new_seeded", "rng_xrs128p_new", "rng_xrs128p_new_seeded",
    "htn_batch_init", "htn_hyperplane_truncated_mvn_batch", "htn_structured_precision_mvn_batch",
    "htn_rng_raw_fill", "htn_std_normal_fill", "htn_device_count", "htn_last_error", "htn_launch_count",
    "htn_version", "htn_fp64_tensor_peak_tflops", "htn_fp64_fma_peak_tflops", "htn_int_issue_peak_gops",
    "htn_phase_timing_enable", "htn_phase_timing_get",
]

HTNORM_ALLOC_ERROR = -100
HTN_E_CUDA, HTN_E_ARG, HTN_E_NO_DEVICE, HTN_E_FOREIGN_RNG = -201, -202, -203, -204
HTN_FLAG_PAPER_SIGN, HTN_FLAG_COLMAJOR = 0x1, 0x2
HTN_MEM_HOST, HTN_MEM_DEVICE = 0, 1
GEN = {"pcg64": 0, "xrs128p": 1, "xoroshiro128p": 1, 0: 0, 1: 1}
TYPES = {"regular": 0, "normal": 0, "diagonal": 1, "identity": 2, 0: 0, 1: 1, 2: 2}

_dblp = C.POINTER(C.c_double)


class RngT(C.Structure):  # include/htnorm_rng.h
    _fields_ = [("base", C.c_void_p), ("next_uint64", C.c_void_p), ("next_double", C.c_void_p)]


class HtConfig(C.Structure):  # include/htnorm.h
    _fields_ = [("gnrow", C.c_size_t), ("gncol", C.c_size_t), ("mean", C.c_void_p), ("cov", C.c_void_p),
                ("g", C.c_void_p), ("r", C.c_void_p), ("diag", C.c_bool)]


class SpConfig(C.Structure):  # include/htnorm.h
    _fields_ = [("a_id", C.c_int), ("o_id", C.c_int), ("pnrow", C.c_size_t), ("pncol", C.c_size_t),
                ("mean", C.c_void_p), ("a", C.c_void_p), ("phi", C.c_void_p), ("omega", C.c_void_p),
                ("struct_mean", C.c_bool)]


class Batch(C.Structure):  # include/htnorm_batch.h
    _fields_ = [("nsamples", C.c_size_t), ("gen", C.c_int), ("seeds", C.c_void_p), ("seed0", C.c_uint64),
                ("independent", C.c_int), ("mem", C.c_int), ("device", C.c_int), ("stream", C.c_void_p),
                ("flags", C.c_uint)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(htnorm_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.htn_last_error.restype = C.c_char_p
    lib.htn_version.restype = C.c_char_p
    lib.htn_launch_count.restype = C.c_ulonglong
    lib.htn_device_count.restype = C.c_int
    lib.htn_fp64_tensor_peak_tflops.restype = C.c_double
    lib.htn_fp64_tensor_peak_tflops.argtypes = [C.c_int, C.c_double]
    lib.htn_fp64_fma_peak_tflops.restype = C.c_double
    lib.htn_fp64_fma_peak_tflops.argtypes = [C.c_int, C.c_double]
    lib.htn_int_issue_peak_gops.restype = C.c_double
    lib.htn_int_issue_peak_gops.argtypes = [C.c_int, C.c_double]
    lib.rng_pcg64_new.restype = C.POINTER(RngT)
    lib.rng_pcg64_new_seeded.restype = C.POINTER(RngT)
    lib.rng_pcg64_new_seeded.argtypes = [C.c_uint64]
    lib.rng_xrs128p_new.restype = C.POINTER(RngT)
    lib.rng_xrs128p_new_seeded.restype = C.POINTER(RngT)
    lib.rng_xrs128p_new_seeded.argtypes = [C.c_uint64]
    lib.rng_free.argtypes = [C.POINTER(RngT)]
    lib.rng_free.restype = None
    lib.htn_batch_init.argtypes = [C.POINTER(Batch), C.c_size_t]
    lib.htn_batch_init.restype = None
    lib.htn_hyperplane_truncated_mvn.argtypes = [C.POINTER(RngT), C.POINTER(HtConfig), C.c_void_p]
    lib.htn_structured_precision_mvn.argtypes = [C.POINTER(RngT), C.POINTER(SpConfig), C.c_void_p]
    lib.htn_hyperplane_truncated_mvn_batch.argtypes = [C.POINTER(Batch), C.POINTER(HtConfig), C.c_void_p,
                                                       C.c_void_p]
    lib.htn_structured_precision_mvn_batch.argtypes = [C.POINTER(Batch), C.POINTER(SpConfig), C.c_void_p,
                                                       C.c_void_p]
    lib.htn_rng_raw_fill.argtypes = [C.POINTER(Batch), C.c_size_t, C.c_void_p]
    lib.htn_std_normal_fill.argtypes = [C.POINTER(Batch), C.c_size_t, C.c_void_p, C.c_void_p]
    return lib


lib = _load()


def load_dev():
    """tests/ and tools/ only: libhtnorm_b200_dev.so = the product objects + csrc/host/debug.c (htn_dbg_*: GEMM,
    Cholesky and solve hooks with host arrays, latency probes).  The product library does not export them."""
    path = os.path.join(_HERE, "libhtnorm_b200_dev.so")
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    dev = C.CDLL(path)
    dev.htn_last_error.restype = C.c_char_p
    return dev


def last_error():
    return lib.htn_last_error().decode()


def check(rc, what):
    """Map a return code to the exceptions the reference's Python wrapper raises
    (pyhtnorm/_htnorm.pyx:98-109) plus this library's own failures."""
    if rc == 0:
        return
    if rc == HTNORM_ALLOC_ERROR:
        raise MemoryError(f"{what}: not enough memory to allocate resources ({last_error()})")
    if rc > 0:
        raise ValueError("Either the leading minor of the {}th order is not positive definite "
                         "(meaning the covariance matrix is also not positive definite), or "
                         "factorization of one of the inputs returned a `U` with a zero in the {}th "
                         "diagonal.".format(rc, rc))
    if rc in (HTN_E_CUDA, HTN_E_NO_DEVICE):
        raise RuntimeError(f"{what}: {last_error()}")
    if rc == HTN_E_FOREIGN_RNG:
        raise TypeError(f"{what}: {last_error()}")
    raise ValueError(f"{what}: illegal argument ({rc}): {last_error()}")
