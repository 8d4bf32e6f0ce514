"""single process, two devices: copy-engine peer copy vs an SM copy kernel (torch elementwise) between cuda:0 and cuda:1"""
import time, torch
n = 65536 * 1000
a = torch.randn(n, dtype=torch.float64, device="cuda:0")
b = torch.empty(n, dtype=torch.float64, device="cuda:1")
print("can access:", torch.cuda.can_device_access_peer(0, 1))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    return (time.perf_counter() - t0) / reps
t = timeit(lambda: b.copy_(a, non_blocking=True))
print(f"same-process copy_ cuda:0 -> cuda:1: {8*n/t/1e9:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print(subprocess.run(["nvidia-smi", "nvlink", "-s", "-i", "0"], capture_output=True, text=True).stdout[:800])
