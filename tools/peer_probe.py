"""Peer-copy bandwidth probes (torchrun, N ranks): torch copy_ into an IPC-mapped peer tensor, cudaMemcpyPeerAsync through
cuda-python / ctypes on the same pointers, and NCCL all_gather for comparison."""
import ctypes, os, time
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 65536 * 1000
src = torch.randn(n, dtype=torch.float64, device=dev)
full = torch.empty(world * n, dtype=torch.float64, device=dev)
handles = [None] * world
dist.all_gather_object(handles, reduce_tensor(full))
peers = [full if q == rank else handles[q][0](*handles[q][1]) for q in range(world)]
dist.barrier()
q = (rank + 1) % world
dst = peers[q][rank * n:(rank + 1) * n]
if rank == 0:
    print("can access peer:", torch.cuda.can_device_access_peer(local, q), "dst device", dst.device, flush=True)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = timeit(lambda: dst.copy_(src, non_blocking=True))
if rank == 0: print(f"torch copy_ into IPC peer tensor: {8*n/t/1e9:.1f} GB/s", flush=True)
# raw runtime call on the same pointers
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpyPeerAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
st = torch.cuda.Stream()
def raw():
    rc = rt.cudaMemcpyPeerAsync(dst.data_ptr(), q, src.data_ptr(), local, 8 * n, st.cuda_stream)
    assert rc == 0, rc
t = timeit(lambda: (raw(), st.synchronize()))
if rank == 0: print(f"cudaMemcpyPeerAsync same pointers: {8*n/t/1e9:.1f} GB/s", flush=True)
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
def raw2():
    rc = rt.cudaMemcpyAsync(dst.data_ptr(), src.data_ptr(), 8 * n, 4, st.cuda_stream)   # cudaMemcpyDefault
    assert rc == 0, rc
t = timeit(lambda: (raw2(), st.synchronize()))
if rank == 0: print(f"cudaMemcpyAsync(default) same pointers: {8*n/t/1e9:.1f} GB/s", flush=True)
t = timeit(lambda: dist.all_gather_into_tensor(full, src))
if rank == 0: print(f"nccl all_gather: {8*n*(world-1)/t/1e9:.1f} GB/s received per rank", flush=True)
dist.barrier()
dist.destroy_process_group()
