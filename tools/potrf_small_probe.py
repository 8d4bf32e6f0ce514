"""Time of htn_potrf for small orders: n = 128 is one fused panel launch alone, the increments show the cost of one
more panel step (panel kernel + trailing update + launch gaps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htnorm_b200 as h
DEV = h._lib.load_dev()   # htn_dbg_* hooks: dev build only
DEV.htn_dbg_potrf_time.restype = C.c_double
for n in (64, 104, 128, 256, 384, 512, 640, 768, 896, 1000, 1024):
    info = C.c_int(0)
    ms = min(DEV.htn_dbg_potrf_time(n, 20, C.byref(info)) for _ in range(3))
    print(f"potrf n={n:5d}: {ms * 1e3:8.1f} us  info={info.value}", flush=True)
