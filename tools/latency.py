import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htnorm_b200 as h
names = ["dfma", "dmma", "rsqrt()", "1/sqrt()", "lds+i2f", "bar", "shfl+dadd"]
for nt in (32, 512, 1024):
    out = (C.c_double * 8)()
    rc = h.lib.htn_dbg_latency(nt, out)
    print(nt, "threads:", rc, {n: round(out[i], 1) for i, n in enumerate(names)})
