"""Multi-GPU functional check (run under torchrun, one process per GPU, NCCL): every rank draws its shard of ONE
logical batch, the shards are all-gathered over NVLink, and rank 0 compares the result bit for bit with the same
batch drawn by a single call on its own GPU - the sharding rule of htnorm_b200.dist."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/htn_nccl_%h_%p.log")
import numpy as np, torch, torch.distributed as dist
import htnorm_b200 as h, cases

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
ok = True
for name, B in (("k7_dense_333", 10007), ("k1_dense_1000_spd", 4099)):
    c = cases.ht_case(name)
    t = [torch.from_numpy(np.ascontiguousarray(c[n])).to(dev) for n in ("mean", "cov", "g", "r")]
    sampler = lambda n, seeds, seed0: h.hyperplane_truncated_mvnorm_batch(n, *t, seeds=seeds, seed0=seed0)
    out, info = h.sample_sharded(sampler, B, rank, world, seed0=777)
    full = h.gather_rows(out, B)
    torch.cuda.synchronize()
    if rank == 0:
        ref, _ = h.hyperplane_truncated_mvnorm_batch(B, *t, seed0=777)
        torch.cuda.synchronize()
        same = bool(torch.equal(full, ref))
        ok &= same
        print(f"{name}: B={B} over {world} GPUs (shards {h.shard_sizes(B, world)[:3]}...), gathered == single call: {same}", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)
