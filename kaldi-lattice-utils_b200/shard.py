"""Multi-GPU sharding of a lattice batch.

Lattices are independent, so a batch is cut into one shard per GPU by total arc
count (greedy longest-first bin packing), every rank runs the whole hot path on
its own shard with its own engine/stream -- no collective on the data path -- and
rank 0 puts the per-lattice results back in INPUT order (the TaskSequencer
contract of the reference, P9).  The only exchange is the gather of result rows;
torch.distributed (NCCL on the GPU box, gloo in the CPU tests) carries it.
"""
import heapq

import numpy as np


def partition_by_arcs(arc_counts, nshards):
    """Greedy LPT: lattices by descending arc count, each to the lightest shard.
    Returns a list of nshards index lists (each ascending, i.e. input order)."""
    arc_counts = np.asarray(arc_counts, dtype=np.int64)
    heap = [(0, r) for r in range(nshards)]
    heapq.heapify(heap)
    shards = [[] for _ in range(nshards)]
    for i in np.argsort(-arc_counts, kind="stable"):
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + int(arc_counts[i]), r))
    return [sorted(s) for s in shards]


def run_sharded(lattices, run_shard, rank=0, world=1, gather=None):
    """Runs `run_shard(list_of_lattices) -> list of per-lattice results` on this rank's
    shard and returns, on rank 0, the results of ALL lattices in input order (None on
    the other ranks).  `gather(obj) -> list over ranks` defaults to
    torch.distributed.gather_object when world > 1."""
    counts = [lat.narcs for lat in lattices]
    shards = partition_by_arcs(counts, world)
    mine = shards[rank]
    local = run_shard([lattices[i] for i in mine])
    if len(local) != len(mine):
        raise RuntimeError("run_shard returned %d results for %d lattices" % (len(local), len(mine)))
    if world == 1:
        parts = [list(zip(mine, local))]
    else:
        if gather is None:
            import torch.distributed as dist

            def gather(obj):
                out = [None] * world if rank == 0 else None
                dist.gather_object(obj, out, dst=0)
                return out
        parts = gather(list(zip(mine, local)))
    if rank != 0:
        return None
    merged = [None] * len(lattices)
    for part in parts:
        for i, res in part:
            merged[i] = res
    return merged
