"""torchrun: torch symmetric memory (cuMem-based peer mappings) + copy-engine copies"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 65536 * 1000
src = torch.randn(n, dtype=torch.float64, device=dev)
full = symm_mem.empty(world * n, dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(full, dist.group.WORLD.group_name)
peers = [hdl.get_buffer(q, (world * n,), torch.float64) for q in range(world)]
if rank == 0: print("peer buffer devices:", [p.device for p in peers], type(hdl).__name__, flush=True)
q = (rank + 1) % world
dst = peers[q][rank * n:(rank + 1) * n]
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = timeit(lambda: dst.copy_(src, non_blocking=True))
if rank == 0: print(f"copy_ into symmetric-memory peer buffer: {8*n/t/1e9:.1f} GB/s", flush=True)
dist.barrier()
ok = torch.equal(full[q * n:(q + 1) * n] if False else full[((rank - 1) % world) * n:(((rank - 1) % world) + 1) * n], full[((rank - 1) % world) * n:(((rank - 1) % world) + 1) * n])
# the neighbour wrote its src into my buffer: compare checksums
s = torch.zeros(world, dtype=torch.float64, device=dev); s[rank] = src.sum(); dist.all_reduce(s)
prev = (rank - 1) % world
print(rank, "received slice checksum matches:", bool(full[prev * n:(prev + 1) * n].sum() == s[prev]), flush=True)
dist.barrier()
dist.destroy_process_group()
