"""Multi-GPU: shard independent samples over ranks, one process per GPU.

The path shards naturally (samples are i.i.d. given the inputs; per-sample seeds make the result
independent of the number of ranks), so there is NO data-path collective: rank q draws the contiguous
slice [lo_q, hi_q) with seeds ``seed0 + lo_q + i``, replicating the (small) shared inputs and their
factorisation.  The only collective is the optional gather of the output shards
(``torch.distributed``: NCCL over NVLink on the GPUs, gloo in the CPU tests of this host logic).
"""
import numpy as np


def shard_range(nsamples, rank, world):
    """Contiguous slice of rank `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(nsamples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(nsamples, world):
    return [shard_range(nsamples, q, world)[1] - shard_range(nsamples, q, world)[0] for q in range(world)]


def sample_sharded(sampler, nsamples, rank, world, *, seeds=None, seed0=0, **kw):
    """Run ``sampler(n_local, seeds=..., seed0=..., **kw) -> (out, info)`` on this rank's slice.
    `sampler` is e.g. a functools.partial of hyperplane_truncated_mvnorm_batch with the inputs bound."""
    lo, hi = shard_range(nsamples, rank, world)
    local_seeds = None if seeds is None else seeds[lo:hi]
    return sampler(hi - lo, seeds=local_seeds, seed0=(int(seed0) + lo) & (2**64 - 1), **kw)


def gather_rows(local, nsamples, group=None, dst=None):
    """All-gather (dst=None) or gather-to-dst of row shards laid out by shard_range.  `local` is a torch
    tensor [n_local, k] (CUDA for NCCL, CPU for gloo).  Returns the [nsamples, k] tensor (None on the
    ranks that are not `dst`)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(nsamples, world)
    k = local.shape[1]
    if local.shape[0] != sizes[rank]:
        raise ValueError("local shard has the wrong number of rows")
    if world == 1:
        return local
    offs = np.concatenate([[0], np.cumsum(sizes)])
    mx = max(sizes)
    # shards differ by at most one row: pad to the common size so that one equal-sized collective works
    # on both backends, then drop the padding rows
    if local.shape[0] < mx:
        local = torch.cat([local, local.new_zeros((mx - local.shape[0], k))])
    local = local.contiguous()
    if dst is None:
        flat = torch.empty((world * mx, k), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, local, group=group)
        buf = flat.view(world, mx, k)
    else:
        buf = torch.empty((world, mx, k), dtype=local.dtype, device=local.device) if rank == dst else None
        dist.gather(local, list(buf.unbind(0)) if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    if len(set(sizes)) == 1:
        return buf.reshape(nsamples, k)
    full = torch.empty((nsamples, k), dtype=local.dtype, device=local.device)
    for q in range(world):
        full[offs[q]:offs[q + 1]] = buf[q, :sizes[q]]
    return full


class PeerGather:
    """All-gather of equal-sized row shards by COPY-ENGINE pushes into peer-mapped receive buffers - no NCCL kernel,
    hence no CTA taken from the sampling GEMM that owns every SM while the previous step's shard is in flight.

    One process per GPU (torch.distributed for the plumbing only): every rank allocates its receive buffer
    ``full`` [world * rows, k], exports it as a CUDA IPC handle (torch.multiprocessing's reductions), and all ranks
    exchange the handles once (``all_gather_object``).  ``push(shard, event)`` then enqueues, on one dedicated stream per
    peer, a plain ``cudaMemcpyAsync`` (``Tensor.copy_``: device-to-device, peer access enabled) of this rank's shard
    into rows [rank * rows, (rank + 1) * rows) of EVERY rank's buffer, NVLink / NVSwitch carrying each copy at the
    copy engine's rate.  ``wait_sent()`` makes the compute stream wait until this rank's shard buffer may be reused;
    ``finish()`` (copy streams drained on every rank, then one barrier) makes ``full`` complete everywhere.
    Single-node only (CUDA IPC); callers fall back to ``gather_rows`` (NCCL) when construction fails."""

    def __init__(self, rows, k, device, group=None, dtype=None):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows, self.k, self.device = rows, k, torch.device(device)
        dtype = dtype or torch.float64
        self.full = torch.empty((self.world * rows, k), dtype=dtype, device=self.device)
        handle = reduce_tensor(self.full)                     # (rebuild function, picklable arguments)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.peers = []
        for q in range(self.world):
            if q == self.rank:
                self.peers.append(self.full)
            else:
                fn, args = handles[q]
                self.peers.append(fn(*args))                  # the peer's buffer, mapped into this process
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.world)]
        self.sent = None
        lo = self.rank * rows
        self.dst = [p[lo:lo + rows] for p in self.peers]
        dist.barrier(group=group)

    def push(self, shard):
        """enqueue the copies of `shard` (ready on the current stream at this point) to every rank, own buffer included"""
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        done = []
        order = [(self.rank + 1 + i) % self.world for i in range(self.world)]   # ring order: spread the targets
        for q in order:
            st = self.streams[q]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                self.dst[q].copy_(shard, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
            done.append(ev)
        return done

    def wait(self, done):
        """the current stream waits until the copies of an earlier push() have read their source"""
        cur = self.torch.cuda.current_stream(self.device)
        for ev in done:
            cur.wait_event(ev)

    def finish(self):
        """every push of every rank has landed: local copy streams drained, then a barrier"""
        for st in self.streams:
            st.synchronize()
        self.dist.barrier(group=self.group)

    def close(self):
        self.dist.barrier(group=self.group)                   # nobody may still be writing into a buffer about to go
        self.dst = self.peers = None
        self.full = None
