"""one htn_potrf of order n (default 8192) after a warm-up call, for launch lists"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htnorm_b200 as h
DEV = h._lib.load_dev()
DEV.htn_dbg_potrf_time.restype = C.c_double
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
info = C.c_int(0)
print(n, DEV.htn_dbg_potrf_time(n, 1, C.byref(info)), info.value)
