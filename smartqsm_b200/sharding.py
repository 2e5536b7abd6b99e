"""Tree-sharded multi-GPU driver (north_star: "work is partitioned across the 8 B200s by tree").

The reference processes clouds one after the other with a fresh algorithm object per file and no
shared state (``entrypoints/smartqsm.py:837`` loop, ``:918-921`` object creation), so trees are
independent units: one process per GPU, every rank takes a subset of the trees, and there is NO
collective on the data path - only a host-side gather of the per-tree summaries at the end.  Trees are
assigned by longest-processing-time-first bin packing on their point counts.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

__all__ = ["partition_trees", "run_sharded"]


def partition_trees(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """LPT bin packing: biggest tree first onto the least loaded rank. Deterministic (ties -> lowest rank,
    lowest tree index), identical on every rank, so no communication is needed to agree on it."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(sizes[i])
    for p in parts:
        p.sort()
    return parts


def run_sharded(sizes: Sequence[int], process_tree: Callable[[int], Dict[str, Any]], rank: Optional[int] = None,
                world_size: Optional[int] = None) -> Optional[List[Dict[str, Any]]]:
    """Run ``process_tree(i)`` for the trees of this rank and gather the summaries on rank 0.

    ``process_tree`` does the whole stage for tree ``i`` on this rank's GPU (e.g. drives a
    ``SpconvBasedContraction``) and returns a small picklable summary.  Uses the default
    ``torch.distributed`` process group when it is initialised (NCCL on the GPU box, gloo in the CPU
    tests); works without one for a single process.  Returns the list of summaries ordered by tree index
    on rank 0 and ``None`` elsewhere.
    """
    import torch.distributed as dist
    have_pg = dist.is_available() and dist.is_initialized()
    if rank is None:
        rank = dist.get_rank() if have_pg else 0
    if world_size is None:
        world_size = dist.get_world_size() if have_pg else 1
    mine = partition_trees(sizes, world_size)[rank]
    local = []
    for i in mine:
        s = dict(process_tree(i))
        s["tree"] = i
        s["rank"] = rank
        local.append(s)
    if not have_pg or world_size == 1:
        return sorted(local, key=lambda s: s["tree"])
    gathered: Optional[List[Any]] = [None] * world_size if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    flat = [s for part in gathered for s in part]
    flat.sort(key=lambda s: s["tree"])
    if [s["tree"] for s in flat] != list(range(len(sizes))):
        raise RuntimeError("tree sharding lost or duplicated a tree")
    return flat
