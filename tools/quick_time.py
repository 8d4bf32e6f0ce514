"""Quick device-side timing of the headline path and its pieces (CUDA events on the torch stream)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import htnorm_b200 as h
import cases

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

res = {}
res["dmma_peak_tflops"] = h.fp64_tensor_peak_tflops(0, 100.0)
res["dfma_peak_tflops"] = h.fp64_fma_peak_tflops(0, 100.0)
print(res, flush=True)
dev = "cuda:0"
for (k, k2, B) in [(1000, 1, 65536), (1000, 1, 8192), (4096, 64, 8192), (64, 4, 65536)]:
    rng = np.random.default_rng(0)
    cov = cases.spd(rng, k); g = rng.standard_normal((k2, k)); r = rng.standard_normal(k2); mu = rng.standard_normal(k)
    t = [torch.from_numpy(x).to(dev) for x in (mu, cov, g, r)]
    out = torch.empty((B, k), dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
    l0 = h.launch_count()
    f = lambda: h.hyperplane_truncated_mvnorm_batch(B, *t, seed0=1, out=out, info=info)
    f(); l1 = h.launch_count()
    best, med = timeit(f)
    flops = k**3/3 + B*(k*k + 4*k*k2)
    res[f"ht_k{k}_k2{k2}_B{B}"] = dict(ms_best=best, ms_med=med, samples_per_s=B/(best*1e-3), tflops=flops/(best*1e-3)/1e12, launches=l1-l0)
    print(k, k2, B, res[f"ht_k{k}_k2{k2}_B{B}"], flush=True)
# pieces: normals only
for (B, n) in [(65536, 1000), (8192, 4096)]:
    import ctypes as C
    z = torch.empty((B, n), dtype=torch.float64, device=dev)
    b = h.api._batch(B, "pcg64", None, 1, False, 0, 0)
    f = lambda: h.lib.htn_std_normal_fill(C.byref(b), C.c_size_t(n), C.c_void_p(z.data_ptr()), None)
    best, med = timeit(f)
    res[f"normals_B{B}_n{n}"] = dict(ms_best=best, gnormals_per_s=B*n/(best*1e-3)/1e9, gbytes_per_s=B*n*8/(best*1e-3)/1e9)
    print("normals", B, n, res[f"normals_B{B}_n{n}"], flush=True)
# independent small problems
for (k, k2, B) in [(64, 4, 1 << 16)]:
    rng = np.random.default_rng(1)
    cov1 = cases.spd(rng, k)
    cov = torch.from_numpy(cov1).to(dev).repeat(B, 1, 1).contiguous()
    g = torch.randn(B, k2, k, dtype=torch.float64, device=dev); r = torch.randn(B, k2, dtype=torch.float64, device=dev)
    mu = torch.randn(B, k, dtype=torch.float64, device=dev)
    out = torch.empty((B, k), dtype=torch.float64, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev)
    f = lambda: h.hyperplane_truncated_mvnorm_batch(B, mu, cov, g, r, seed0=1, independent=True, out=out, info=info)
    best, med = timeit(f)
    bytes_ = B * 8 * (k*k + k2*k + 2*k + k2)
    res[f"small_k{k}_B{B}"] = dict(ms_best=best, problems_per_s=B/(best*1e-3), gbytes_per_s=bytes_/(best*1e-3)/1e9)
    print("small", res[f"small_k{k}_B{B}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)
