"""Row f1 timing: graph construction + skeleton extraction on the skeletal points of a contracted synthetic tree.

    python scripts/graph_bench.py [n_points=2000000] [--reference]

Prints one JSON line: skeletal points, edges, seconds for Qhull (SciPy, host - the same call the reference makes), for the
device / C++ stage (`sqsm_graph_construct` + `sqsm_graph_extract_skeleton`) and - with --reference - for the restated
reference path (networkx + SciPy exactly as core/skeletonization_pipeline_base.py:75-238 uses them) on a bounded sub-sample.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200 import checkpoints  # noqa: E402
from smartqsm_b200._lib import Engine  # noqa: E402
from smartqsm_b200.synthetic import make_tree  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 2_000_000
eng = Engine(monitored_by_chamfer_distance=True, latch_early_stop=True)
eng.load_state_dict(checkpoints.load_state("leafoff"))
eng.set_points(make_tree(n, 2))
for _ in range(10):
    eng.contract_step()
ids, sk, rad = eng.skeletal_sample(0.01)
from scipy.spatial import Delaunay  # noqa: E402
t0 = time.perf_counter()
simplices = Delaunay(sk).simplices
t_qhull = time.perf_counter() - t0
t0 = time.perf_counter()
n_edges = eng.graph_construct(simplices, r=0.055, fetch=False)
t_construct = time.perf_counter() - t0
t0 = time.perf_counter()
skel = eng.graph_extract_skeleton(len(sk))
t_extract = time.perf_counter() - t0
out = {"input_points": n, "skeletal_points": int(len(sk)), "tetrahedra": int(len(simplices)), "edges": int(n_edges),
       "reached": int((skel["pred"] >= 0).sum()) + 1, "qhull_host_s": t_qhull, "construct_s": t_construct, "extract_s": t_extract,
       "info": eng.graph_info()}
if "--reference" in sys.argv:
    from oracle import graph as OG
    m = min(len(sk), 60000)
    sub = np.ascontiguousarray(sk[:m])
    t0 = time.perf_counter()
    cg = OG.construct_graph(sub)
    so = OG.extract_skeleton(sub, cg["edges"])
    t_ref = time.perf_counter() - t0
    e2 = Engine()
    t0 = time.perf_counter()
    sx = Delaunay(sub).simplices
    tq = time.perf_counter() - t0
    t0 = time.perf_counter()
    ed = e2.graph_construct(sx, r=0.055, points=sub)
    s2 = e2.graph_extract_skeleton(m)
    t_dev = time.perf_counter() - t0
    out["reference_subsample"] = {"skeletal_points": m, "networkx_scipy_s": t_ref, "device_stage_s": t_dev, "qhull_s": tq,
                                  "identical": bool(np.array_equal(ed, cg["edges"]) and np.array_equal(s2["pred"], so["pred"]))}
print(json.dumps(out))
