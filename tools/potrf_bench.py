"""Throughput of the blocked Cholesky alone (htn_potrf) against the measured FP64 tensor peak."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htnorm_b200 as h
DEV = h._lib.load_dev()   # htn_dbg_* hooks: dev build only
DEV.htn_dbg_potrf_time.restype = C.c_double
peak = h.fp64_tensor_peak_tflops(0, 100.0)
print(f"DMMA peak {peak:.1f} TFLOP/s")
sizes = [int(a) for a in sys.argv[1:]] or [1000, 2000, 4096, 8192, 16384]
for n in sizes:
    info = C.c_int(0)
    ms = DEV.htn_dbg_potrf_time(n, 3, C.byref(info))
    tf = n ** 3 / 3 / (ms * 1e-3) / 1e12
    print(f"potrf n={n:6d}: {ms:8.3f} ms  {tf:6.2f} TFLOP/s  {100 * tf / peak:5.1f}% of peak  info={info.value}")
