"""One tree sharded over the ranks of a torchrun job (config C5 at reduced size): checks bit-identity against the
single-GPU result computed by rank 0 and reports the strong-scaling time.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/sharded_tree_demo.py 5000000
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200 import checkpoints as weights
from smartqsm_b200._lib import Engine  # noqa: E402
from smartqsm_b200.sharding import ShardedTreeContraction  # noqa: E402
from smartqsm_b200.synthetic import make_tree  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
state = weights.load_state("leafon")
pts = make_tree(n, 5, True)                      # same seed on every rank: the same tree
kw = dict(device=local, monitored_by_chamfer_distance=True)
eng = Engine(**kw)
eng.load_state_dict(state)
eng.set_points(pts)
drv = ShardedTreeContraction(eng, rank, world, dev, monitored=True)
drv.contract()                                   # warm-up thunk (allocator, module load)
eng.set_points(pts)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(iters):
    st = drv.contract()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
P, D = eng.result()
ok = True
t1 = float("nan")
if rank == 0:
    ref = Engine(**kw)
    ref.load_state_dict(state)
    ref.set_points(pts)
    ref.contract_step()
    ref.set_points(pts)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        ref.contract_step()
    t1 = time.perf_counter() - t
    P1, D1 = ref.result()
    ok = np.array_equal(P, P1) and np.array_equal(D, D1)
    print({"points": n, "ranks": world, "thunks": iters, "sharded_s": round(dt, 4), "single_gpu_s": round(t1, 4),
           "speedup": round(t1 / dt, 3), "bit_identical_to_single_gpu": bool(ok), "segments": st["segment_sizes"],
           "accepted_last": st["accepted"], "phase_s_last": {k: round(v, 4) for k, v in st["phase_s"].items()}})
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
