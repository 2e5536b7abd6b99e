"""Tree-sharded multi-GPU driver (north_star: "work is partitioned across the 8 B200s by tree").

The reference processes clouds one after the other with a fresh algorithm object per file and no
shared state (``entrypoints/smartqsm.py:837`` loop, ``:918-921`` object creation), so trees are
independent units: one process per GPU, every rank takes a subset of the trees, and there is NO
collective on the data path - only a host-side gather of the per-tree summaries at the end.  Trees are
assigned by longest-processing-time-first bin packing on their point counts.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

__all__ = ["partition_trees", "run_sharded", "ShardedTreeContraction", "emulate_sharded_step"]


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


# ------------------------------------------------------------------------------------------------------------------
# One oversized tree sharded over several GPUs
# ------------------------------------------------------------------------------------------------------------------
#
# Within an iteration the reference's cube samples are independent (each carries its own 0.4 m halo,
# external/SmartTreeXX/datasets/synthetic_tree.py:252-281) and are consumed 16 at a time (batch_size,
# core/core_algorithms/spconv_based_contraction.py:92-100).  Every rank holds the whole cloud (24 B per point), does
# the cheap tiling itself and runs voxelisation -> rulebooks -> U-Net only for a contiguous range of whole batches
# (same batch composition as on one GPU, so integer and float outputs are identical).  The exchange step is the
# all-gather of the moved points (24 B per point, NCCL over NVLink) plus an all-reduce of two float64 sums for the
# Chamfer monitor, whose nearest-neighbour queries are split by brick.

class ShardedTreeContraction:
    """Drive one engine per rank through the phased thunk of ``include/sqsm_b200.h`` (sqsm_step_*).

    ``engine`` is a ``smartqsm_b200._lib.Engine`` (or any object with the same ``step_*`` methods) that already
    holds the full cloud (``set_points``) on ``device``.  Uses the default ``torch.distributed`` group (NCCL on
    GPUs, gloo on CPU tensors in the tests).

    The exchange step is ONE all-gather per thunk: every rank contributes its segment (new points and accumulated
    displacements, 24 B per point, padded to the longest segment) and receives everybody's; the concatenation in rank
    order is the new cloud in canonical order.  On GPUs the collective is enqueued relative to the ENGINE's stream
    (``torch.cuda.ExternalStream``), so there is no device-wide synchronisation between the network phase, the exchange
    and the Chamfer phase - only the tiny host read of the segment sizes.
    """

    def __init__(self, engine, rank: int, world_size: int, device, monitored: bool) -> None:
        self.engine, self.rank, self.world, self.device, self.monitored = engine, rank, world_size, device, monitored
        self.history: List[Dict[str, Any]] = []
        self.exchange_bytes = 0
        self._stream = None
        if getattr(device, "type", "cpu") == "cuda" and hasattr(engine, "stream_ptr"):
            import torch
            self._stream = torch.cuda.ExternalStream(engine.stream_ptr(), device=device)

    def _ctx(self):
        import contextlib
        import torch
        return torch.cuda.stream(self._stream) if self._stream is not None else contextlib.nullcontext()

    def contract(self) -> Dict[str, Any]:
        """One ``_contract`` thunk over all ranks; returns this rank's stats with the global decision."""
        import time
        import torch
        import torch.distributed as dist
        eng = self.engine
        t0 = time.perf_counter()
        st = eng.step_begin(self.rank, self.world)
        t1 = time.perf_counter()
        if st["skipped"]:
            self.history.append(st)
            return st
        with self._ctx():
            sizes_t = torch.zeros(self.world, dtype=torch.int64, device=self.device)
            sizes_t[self.rank] = st["n_out"]
            if self.world > 1:
                dist.all_reduce(sizes_t)
            sizes = [int(x) for x in sizes_t.tolist()]
            total, longest = sum(sizes), max(sizes)
            if total == 0:
                raise RuntimeError("contraction produced no interior voxel")
            # send[0] = points, send[1] = displacements of this rank's segment, padded to the longest segment
            send = torch.empty((2, longest, 3), dtype=torch.float32, device=self.device)
            if sizes[self.rank]:
                eng.step_segment(send[0].data_ptr(), send[1].data_ptr())
            if self.world > 1:
                flat = torch.empty((self.world * 2, longest, 3), dtype=torch.float32, device=self.device)
                dist.all_gather_into_tensor(flat, send)           # concatenation along dim 0 (gloo accepts only this form)
                recv = flat.view(self.world, 2, longest, 3)
                self.exchange_bytes = flat.numel() * 4
                new_p = torch.cat([recv[r, 0, :sizes[r]] for r in range(self.world)])
                new_d = torch.cat([recv[r, 1, :sizes[r]] for r in range(self.world)])
            else:
                new_p, new_d = send[0, :total].contiguous(), send[1, :total].contiguous()
            if self._stream is None and getattr(self.device, "type", "cpu") == "cuda":
                torch.cuda.synchronize(self.device)           # no shared stream: order by the host
            t2 = time.perf_counter()
            eng.step_set_candidate(new_p.data_ptr(), new_d.data_ptr(), total)
            cd = float("nan")
            if self.monitored:
                sums, counts = eng.step_chamfer(self.rank, self.world)
                t = torch.tensor(sums, dtype=torch.float64, device=self.device)
                if self.world > 1:
                    dist.all_reduce(t)
                cd = float(t[0]) / counts[0] + float(t[1]) / counts[1]
            st["accepted"] = int(eng.step_finish(self.monitored, cd))
            if self._stream is not None:
                new_p.record_stream(self._stream)
                new_d.record_stream(self._stream)
        t3 = time.perf_counter()
        st["phase_s"] = {"begin": t1 - t0, "exchange": t2 - t1, "chamfer_finish": t3 - t2}
        st["chamfer"] = cd
        st["n_out_total"] = total
        st["segment_sizes"] = sizes
        self.history.append(st)
        return st


def emulate_sharded_step(engines, monitored: bool, device) -> Dict[str, Any]:
    """The same exchange with all "ranks" in ONE process (several engines on one GPU): used by the single-GPU
    parity test, because separate processes that wait on each other must not share a GPU."""
    import torch
    world = len(engines)
    stats = [e.step_begin(r, world) for r, e in enumerate(engines)]
    if stats[0]["skipped"]:
        return stats[0]
    sizes = [s["n_out"] for s in stats]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    total = offs[-1]
    new_p = torch.empty((total, 3), dtype=torch.float32, device=device)
    new_d = torch.empty((total, 3), dtype=torch.float32, device=device)
    for r, e in enumerate(engines):
        if sizes[r]:
            e.step_segment(new_p[offs[r]:offs[r + 1]].data_ptr(), new_d[offs[r]:offs[r + 1]].data_ptr())
    for e in engines:
        e.step_set_candidate(new_p.data_ptr(), new_d.data_ptr(), total)
    cd = float("nan")
    if monitored:
        tot = [0.0, 0.0]
        counts = None
        for r, e in enumerate(engines):
            s, counts = e.step_chamfer(r, world)
            tot[0] += s[0]
            tot[1] += s[1]
        cd = tot[0] / counts[0] + tot[1] / counts[1]
    acc = [e.step_finish(monitored, cd) for e in engines]
    assert len(set(acc)) == 1
    out = dict(stats[0])
    out.update(accepted=int(acc[0]), chamfer=cd, n_out_total=total, segment_sizes=sizes)
    return out
