"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): two thunks with the Chamfer monitor on a
6 000-point tree (hash inserts, first-come winners, rulebooks, FFMA + tcgen05 convolutions, heads, Chamfer voxel sums and
nearest neighbours), skeletal sampling, radius graph and the graph / skeleton stage.

    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200 import checkpoints  # noqa: E402
from smartqsm_b200._lib import Engine  # noqa: E402
from smartqsm_b200.synthetic import make_tree  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
eng = Engine(monitored_by_chamfer_distance=True, batch_size=4)
eng.load_state_dict(checkpoints.load_state("leafoff"))
eng.set_points(make_tree(n, 9))
for _ in range(2):
    st = eng.contract_step()
ids, sk, rad = eng.skeletal_sample(0.01)
edges = eng.radius_graph(0.055)
from scipy.spatial import Delaunay  # noqa: E402
e2 = eng.graph_construct(Delaunay(sk).simplices, r=0.055)
skel = eng.graph_extract_skeleton(len(sk))
print("ok", st["n_out"], len(ids), len(edges), len(e2), int((skel["pred"] >= 0).sum()))
