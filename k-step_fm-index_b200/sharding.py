"""Host-side multi-GPU plumbing: the batch is split into contiguous 32-aligned shards (one per GPU / rank),
the index is replicated, and only per-shard (L,R) intervals travel back.  The search itself needs no
collective; torch.distributed is used for the replica broadcast and for timing only.

Mirrors the C driver's partition (fm_host.c: transferCPUtoGPU) so one-process-many-GPUs and
one-process-per-GPU runs shard identically.
"""


def shard_bounds(nqueries, nshards):
    """first[g] for g in 0..nshards: shard g = reads [first[g], first[g+1]); every shard start is a multiple of 32."""
    per = ((nqueries + nshards - 1) // nshards + 31) & ~31
    return [min(per * g, nqueries) for g in range(nshards + 1)]


def shard_range(nqueries, nshards, rank):
    b = shard_bounds(nqueries, nshards)
    return b[rank], b[rank + 1]


def broadcast_meta(meta, src, dist, meta_type):
    """Sends the fmgpu_index_meta_t of the rank that built the index to every rank (object broadcast)."""
    holder = [bytes(meta) if dist.get_rank() == src else None]
    dist.broadcast_object_list(holder, src=src)
    return meta_type.from_buffer_copy(holder[0])


def gather_intervals(local_lr, nqueries, dist, torch):
    """Concatenates the per-rank (L,R) shards on every rank (test/debug helper; the bench never gathers)."""
    world = dist.get_world_size()
    bounds = shard_bounds(nqueries, world)
    per = max(bounds[g + 1] - bounds[g] for g in range(world))
    pad = torch.zeros(2 * per, dtype=torch.int64)
    pad[: local_lr.numel()] = local_lr.to(torch.int64)
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([parts[g][: 2 * (bounds[g + 1] - bounds[g])] for g in range(world)])
