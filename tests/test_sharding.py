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
