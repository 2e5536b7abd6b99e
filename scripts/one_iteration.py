"""One contraction iteration on a synthetic tree (profiling target for ncu)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200 import checkpoints as weights
from smartqsm_b200._lib import Engine
from smartqsm_b200.synthetic import make_tree

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
leaf = len(sys.argv) > 2 and sys.argv[2] == "leafon"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
state = weights.load_state("leafon" if leaf else "leafoff")
pts = make_tree(n, 3, leaf)
eng = Engine(monitored_by_chamfer_distance=True)
eng.load_state_dict(state)
eng.set_points(pts)
for _ in range(iters):
    st = eng.contract_step()
print({k: st[k] for k in ("n_in", "n_vox", "n_rows_l1", "n_rows_l2", "n_rows_l3", "n_out", "ms_network", "ms_chamfer", "ms_rulebook", "ms_voxel")})
