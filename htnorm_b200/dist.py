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
    ``full`` [nbuf, world * rows, k] as torch SYMMETRIC MEMORY (CUDA virtual-memory allocations exported to the peers and
    mapped into every process: ``torch.distributed._symmetric_memory``), so each rank holds a local view of every
    peer's buffer.  The sampler writes its shard straight into ``own(b)`` = rows [rank * rows, (rank + 1) * rows) of its
    own buffer b (no local copy); ``push(b)`` then enqueues a plain
    ``cudaMemcpyAsync`` (``Tensor.copy_``) of that slice into the same rows of every PEER's buffer b (one copy stream,
    ring order),
    NVLink / NVSwitch carrying each copy at the copy engine's rate (measured: 758 GB/s per direction between two
    B200s; the legacy cudaIpc mapping of a caching-allocator block reached 22 GB/s on the same box and is not used).
    ``wait(events)`` makes the compute stream wait until a shard buffer may be reused; ``finish()`` (copy streams
    drained on every rank, then one barrier) makes ``full`` complete everywhere.
    Single-node only; callers fall back to ``gather_rows`` (NCCL) when construction fails."""

    def __init__(self, rows, k, device, group=None, dtype=None, nbuf=2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows, self.k, self.device, self.nbuf = rows, k, torch.device(device), nbuf
        dtype = dtype or torch.float64
        n = nbuf * self.world * rows * k
        flat = symm_mem.empty(n, dtype=dtype, device=self.device)
        name = (group or dist.group.WORLD).group_name
        self.handle = symm_mem.rendezvous(flat, name)
        shape = (nbuf, self.world * rows, k)
        self.full = flat.view(shape)
        self.peers = [self.full if q == self.rank else self.handle.get_buffer(q, (n,), dtype).view(shape)
                      for q in range(self.world)]
        import os
        self.align = os.environ.get("HTN_GATHER_ALIGN", "1") != "0" and self.world > 2
        ncopy = max(1, int(os.environ.get("HTN_GATHER_STREAMS", "1")))   # experiments; one stream measured best
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(ncopy)]
        self.lo = self.rank * rows
        dist.barrier(group=group)

    def own(self, b):
        """this rank's slice of its own receive buffer b: the sampler writes its shard here"""
        return self.full[b, self.lo:self.lo + self.rows]

    def push(self, b):
        """enqueue the copies of own(b) (ready on the current stream at this point) into every peer's buffer b, ONE after
        the other on one copy stream, in ring order (rank + 1, rank + 2, ...): at any moment every GPU sends to exactly
        one peer and receives from exactly one - a permutation through the NVSwitch.  Measured on 8 B200s, 524 MB
        shards: 729 GB/s egress per rank this way, 481 GB/s with the seven copies on seven streams at once, 652 GB/s
        for ncclAllGather.  Returns the event that marks the end of the last copy."""
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        src = self.own(b)
        for st in self.streams:
            st.wait_event(ready)
        if self.align:
            # line the ranks' copy streams up (a device-side barrier through the symmetric-memory signal pads, one
            # tiny kernel): with every rank starting its ring at the same moment the permutation pattern holds
            with torch.cuda.stream(self.streams[0]):
                self.handle.barrier()
            for st in self.streams[1:]:
                st.wait_stream(self.streams[0])
        for i in range(1, self.world):
            q = (self.rank + i) % self.world
            with torch.cuda.stream(self.streams[(i - 1) % len(self.streams)]):
                self.peers[q][b, self.lo:self.lo + self.rows].copy_(src, non_blocking=True)
        done = []
        for st in self.streams:
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
        self.peers = None
        self.full = None
        self.handle = None
