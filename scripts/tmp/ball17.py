import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from smartqsm_b200._lib import Engine
from smartqsm_b200.synthetic import make_tree
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ball_lists_bench import radii_for
rng = np.random.default_rng(0)
eng = Engine()
pts = make_tree(1_700_000, 4).astype(np.float32)
radii = radii_for(pts, rng)
print("ext", pts.max(0) - pts.min(0), "radii q", np.quantile(radii, [0, .25, .5, .75, .99, 1]), "mean", radii.mean(), flush=True)
t0 = time.perf_counter(); off, idx = eng.ball_lists(pts, radii); print("ms", (time.perf_counter() - t0) * 1e3)
c = np.diff(off); print("count q", np.quantile(c, [0, .5, .9, .99, 1]))
