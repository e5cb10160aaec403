"""Multi-GPU plumbing of the path: reads shard, the index is replicated, per-class counts are summed.

One process per GPU.  Rank g maps the contiguous range [g*N/G, (g+1)*N/G) of the reads (SURVEY.md
section 8(e)); there is no exchange step inside the path.  The one collective is the sum of the
per-class counts at the end: on GPUs it is psa_mapper_counts_allreduce (one ncclAllReduce over
NVLink, include/psa.h); `allreduce_counts` below is the same reduction for counts that already
sit in host memory, over whatever torch.distributed backend the job uses (gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_reads, world, rank):
    """[lo, hi) of the reads rank `rank` of `world` maps; the ranges tile [0, n_reads) exactly."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return (n_reads * rank) // world, (n_reads * (rank + 1)) // world


def shard_of_read(i, n_reads, world):
    """The rank whose range holds read i (inverse of shard_range)."""
    if not (0 <= i < n_reads):
        raise ValueError("read index out of range")
    r = (i * world) // n_reads
    while shard_range(n_reads, world, r)[1] <= i:
        r += 1
    while shard_range(n_reads, world, r)[0] > i:
        r -= 1
    return r


def allreduce_counts(counts, group=None):
    """Sum uint64 per-class counts (n_eq + 2 entries, include/psa.h) over the ranks, in place."""
    import torch
    import torch.distributed as dist
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    t = torch.from_numpy(c.view(np.int64))          # counts < 2^63: the int64 sum is the uint64 sum
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    counts[...] = t.numpy().view(np.uint64)
    return counts


def gather_hits(hits, tx, n_reads, world, rank, group=None):
    """Per-read results stay on their rank in the product; this gathers them to rank 0 in input
    order for checks (hits: psa_hit array of this rank's range, tx: its member buffer)."""
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object((hits, tx), parts, dst=0, group=group)
    if rank != 0:
        return None
    all_hits = np.concatenate([p[0] for p in parts])
    base = np.cumsum([0] + [len(p[1]) for p in parts[:-1]]).astype(np.uint64)
    o = 0
    for r, p in enumerate(parts):
        all_hits["tx_off"][o:o + len(p[0])] += base[r]
        o += len(p[0])
    assert len(all_hits) == n_reads
    return all_hits, np.concatenate([p[1] for p in parts])
