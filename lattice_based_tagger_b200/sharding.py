"""Sentence sharding across GPUs (host side).

The decode path shards by independent sentences: every rank holds the full (read-only)
tables and tags a contiguous, length-balanced slice of the batch; there is no collective on the data
path.  The only exchange is the final gather of the per-sentence results on the host, for which
`gather_results` uses `torch.distributed.all_gather_object` (NCCL or gloo process groups alike).
"""

import numpy as np


def shard_bounds(lengths, world_size):
    """Split sentences 0..n-1 into `world_size` contiguous slices of (nearly) equal total work.

    Work of a sentence is estimated by its length in code units (the beam does a bounded amount of
    work per syllable).  Returns `world_size + 1` boundaries; slice r is [bounds[r], bounds[r+1]).
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    n = lengths.size
    if world_size <= 1 or n == 0:
        return [0] + [n] * max(1, world_size)
    cum = np.cumsum(lengths + 1)                      # +1: per-sentence fixed cost, keeps empties moving
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        bounds.append(int(np.searchsorted(cum, target, side='left')))
    bounds.append(n)
    for r in range(1, len(bounds)):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds


def partition_by_work(lengths, world_size):
    """Index arrays, one per rank, of (nearly) equal estimated work AND equal sentence count.

    A sentence's decode work grows with its length (transitions ~ length x (8 k + k E / L), SURVEY §8e;
    beam and lattice density are common to the batch).  Sentences are sorted by length and dealt to the
    ranks in snake order (0 1 .. n-1 n-1 .. 1 0 ...), which balances total length and the length mix;
    each rank's indices are returned in ascending order, so that its slice of the text is read front to
    back.  Unlike `shard_bounds` the parts are not contiguous: results are put back with the index arrays.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    n = lengths.size
    if world_size <= 1:
        return [np.arange(n, dtype=np.int64)]
    order = np.argsort(-lengths, kind='stable')
    pos = np.arange(n, dtype=np.int64)
    lap, col = divmod(pos, world_size)
    owner = np.where(lap % 2 == 0, col, world_size - 1 - col)
    return [np.sort(order[owner == r]) for r in range(world_size)]


def my_shard(sents, rank, world_size):
    """(start, stop) of this rank's slice of `sents`."""
    bounds = shard_bounds([len(s) for s in sents], world_size)
    return bounds[rank], bounds[rank + 1]


def tag_sharded(tag_fn, sents, rank, world_size, gather=True):
    """Tag this rank's slice with `tag_fn(list[str]) -> list`, then gather all slices in order.

    With `gather=False` only the local slice is returned (as `(start, results)`).
    """
    start, stop = my_shard(sents, rank, world_size)
    local = tag_fn(sents[start:stop])
    if not gather:
        return start, local
    return gather_results(local, world_size)


def gather_results(local, world_size):
    """Concatenate per-rank result lists in rank order on every rank."""
    if world_size <= 1:
        return list(local)
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, list(local))
    out = []
    for p in parts:
        out.extend(p)
    return out
