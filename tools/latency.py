import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htnorm_b200 as h
DEV = h._lib.load_dev()   # htn_dbg_* hooks: dev build only
names = ["dfma", "dmma", "rsqrt()", "1/sqrt()", "lds+i2f", "bar", "shfl+dadd"]
for nt in (32, 512, 1024):
    out = (C.c_double * 8)()
    rc = DEV.htn_dbg_latency(nt, out)
    print(nt, "threads:", rc, {n: round(out[i], 1) for i, n in enumerate(names)})
DEV.htn_dbg_dmma_sweep.restype = C.c_double
for w in (4, 8, 16, 32):
    print("dmma sweep warps/SM", w, {ilp: round(DEV.htn_dbg_dmma_sweep(w, ilp), 2) for ilp in (8, 32)})
DEV.htn_dbg_dmma_tile_probe.restype = C.c_double
print("dmma tile probe (smem fragments + 32 DMMA / k4, no loads, no barriers): 1 CTA/SM",
      round(DEV.htn_dbg_dmma_tile_probe(1), 2), " 2 CTAs/SM", round(DEV.htn_dbg_dmma_tile_probe(2), 2))
