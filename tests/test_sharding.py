"""Host-side logic of the tree-sharded multi-GPU path, on CPU with the gloo backend (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from smartqsm_b200.sharding import partition_trees, run_sharded


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1_000_000, 3_000_000, 64).tolist()
    for w in (1, 2, 4, 8):
        parts = partition_trees(sizes, w)
        assert sorted(i for p in parts for i in p) == list(range(64))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)            # LPT guarantee
        assert max(loads) <= 1.05 * sum(sizes) / w + max(sizes) / w
    assert partition_trees([], 4) == [[], [], [], []]
    assert partition_trees([5], 4) == [[0], [], [], []]
    with pytest.raises(ValueError):
        partition_trees([1], 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def process_tree(i):
            rng = np.random.default_rng(i)
            pts = rng.normal(size=(sizes[i], 3))
            return {"points": sizes[i], "checksum": float(pts.sum())}
        out = run_sharded(sizes, process_tree)
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_run_sharded_gloo_world2():
    sizes = [300, 100, 250, 50, 400, 120, 80]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] is None
    out = res[0]
    assert [s["tree"] for s in out] == list(range(len(sizes)))
    parts = partition_trees(sizes, 2)
    for s in out:
        assert s["tree"] in parts[s["rank"]]
        rng = np.random.default_rng(s["tree"])
        assert s["checksum"] == float(rng.normal(size=(sizes[s["tree"]], 3)).sum())


def test_run_sharded_single_process():
    out = run_sharded([3, 1, 2], lambda i: {"v": i * i})
    assert [s["v"] for s in out] == [0, 1, 4]


# ------------------------------------------------------------------ one tree sharded over ranks (driver logic, gloo)

class _FakeEngine:
    """CPU stand-in with the phased-thunk interface: the 'network' moves every point by a fixed vector, the
    segment of rank r is a contiguous slice of the cloud, the Chamfer partial sums are split by index."""

    def __init__(self, pts):
        import torch
        self.P = torch.tensor(pts, dtype=torch.float32)
        self.D = torch.zeros_like(self.P)
        self.cand = None

    def step_begin(self, shard, n_shards):
        n = len(self.P)
        lo, hi = n * shard // n_shards, n * (shard + 1) // n_shards
        self.seg = (self.P[lo:hi] + 0.01, self.D[lo:hi] + 0.01)
        return {"skipped": 0, "n_out": hi - lo, "accepted": 0, "chamfer": float("nan")}

    @staticmethod
    def _copy(dst_ptr, t):
        import ctypes
        ctypes.memmove(dst_ptr, t.contiguous().data_ptr(), t.numel() * 4)

    def step_segment(self, d_xyz=0, d_disp=0):
        if d_xyz:
            self._copy(d_xyz, self.seg[0]); self._copy(d_disp, self.seg[1])
        return len(self.seg[0])

    def step_set_candidate(self, d_xyz, d_disp, n):
        import ctypes, torch
        p = torch.empty((n, 3), dtype=torch.float32); d = torch.empty((n, 3), dtype=torch.float32)
        ctypes.memmove(p.data_ptr(), d_xyz, n * 12); ctypes.memmove(d.data_ptr(), d_disp, n * 12)
        self.cand = (p, d)

    def step_chamfer(self, shard, n_shards):
        p = self.cand[0]
        idx = range(shard, len(p), n_shards)
        s = float(sum(float(((p[i] - self.P[i]) ** 2).sum()) for i in idx))
        return [s, s], [len(p), len(p)]

    def step_finish(self, monitored, cd):
        self.P, self.D = self.cand
        return True


def _tree_worker(rank, world, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smartqsm_b200.sharding import ShardedTreeContraction
        pts = np.arange(30, dtype=np.float32).reshape(10, 3)
        drv = ShardedTreeContraction(_FakeEngine(pts), rank, world, torch.device("cpu"), monitored=True)
        st1 = drv.contract()
        st2 = drv.contract()
        q.put((rank, st1["segment_sizes"], st1["chamfer"], drv.engine.P.numpy().copy(), st2["n_out_total"]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_tree_driver_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tree_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(30, dtype=np.float32).reshape(10, 3) + 0.02
    for rank, sizes, cd, P, total in res:
        assert sizes == [5, 5] and total == 10
        assert np.allclose(P, want)                                   # both ranks hold the whole, identical cloud
        assert abs(cd - 2 * (3 * 0.01 ** 2)) < 1e-6                    # all-reduced partial sums / global counts (float32 data)
