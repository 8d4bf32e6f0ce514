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
