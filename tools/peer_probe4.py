"""torchrun N ranks: all ranks push one 524 MB shard to every peer through symmetric memory; vary the number of copy
streams and the chunking; also NCCL all_gather for reference.  Reports per-rank egress GB/s (max time over ranks)."""
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
def run(nstreams, chunks, reps=4):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    def once():
        ev = torch.cuda.Event(); ev.record()
        for s in streams: s.wait_event(ev)
        k = 0
        for c in range(chunks):
            lo, hi = c * n // chunks, (c + 1) * n // chunks
            for i in range(1, world):
                q = (rank + i) % world
                with torch.cuda.stream(streams[k % nstreams]):
                    peers[q][rank * n + lo:rank * n + hi].copy_(src[lo:hi], non_blocking=True)
                k += 1
        for s in streams: torch.cuda.current_stream().wait_stream(s)
    once(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): once()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"streams {nstreams} chunks {chunks}: {float(t)*1e3:.2f} ms  egress {8*n*(world-1)/float(t)/1e9:.0f} GB/s per rank", flush=True)
for ns, ch in [(7, 1), (1, 1), (2, 1), (3, 1), (4, 1), (7, 4), (2, 8), (14, 2)]:
    run(ns, ch)
g = torch.empty(world * n, dtype=torch.float64, device=dev)
dist.all_gather_into_tensor(g, src); torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(4): dist.all_gather_into_tensor(g, src)
torch.cuda.synchronize()
t = torch.tensor([(time.perf_counter() - t0) / 4], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print(f"nccl all_gather: {float(t)*1e3:.2f} ms  {8*n*(world-1)/float(t)/1e9:.0f} GB/s received per rank", flush=True)
dist.barrier(); dist.destroy_process_group()
