"""Row f2 timing: the cloud preparation (smartqsm.py:882-913) on a synthetic raw scan, GPU against the oracle."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200._lib import Engine
from tests.helpers import make_raw_scan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
cpu_n = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
raw = make_raw_scan(n, 3, leaf_on=True, outliers=max(200, n // 5000))
eng = Engine(device=0)
eng.prepare_cloud(raw)                                  # warm-up (allocator pool, module load)
best = None
for _ in range(3):
    t0 = time.perf_counter()
    st = eng.prepare_cloud(raw)
    wall = (time.perf_counter() - t0) * 1e3
    if best is None or wall < best[0]:
        best = (wall, st)
wall, st = best
out = {"raw_points": len(raw), "n_voxel": st["n_voxel"], "n_kept": st["n_kept"], "scale": st["scale"],
       "wall_ms_incl_h2d": wall, "ms_device": st["ms_total"], "ms_downsample": st["ms_downsample"], "ms_filter": st["ms_filter"],
       "ms_scale": st["ms_scale"], "raw_points_per_s": len(raw) / (wall * 1e-3), "gpu_launches": st["gpu_launches"], "n_far_filter": st["n_far_filter"], "n_far_scale": st["n_far_scale"]}
if cpu_n:
    from oracle import preprocess as OP
    sample = make_raw_scan(cpu_n, 3, leaf_on=True)
    t0 = time.perf_counter()
    OP.prepare_cloud(sample)
    dt = time.perf_counter() - t0
    out["cpu_oracle"] = {"raw_points": len(sample), "seconds": dt, "raw_points_per_s": len(sample) / dt}
print(json.dumps(out))
