"""One contraction thunk + skeletal sampling + radius graph on a synthetic tree (ncu target covering every kernel family)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200 import checkpoints as weights
from smartqsm_b200._lib import Engine
from smartqsm_b200.synthetic import make_tree

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
leaf = len(sys.argv) > 2 and sys.argv[2] == "leafon"
eng = Engine(monitored_by_chamfer_distance=True)
eng.load_state_dict(weights.load_state("leafon" if leaf else "leafoff"))
eng.set_points(make_tree(n, 3, leaf))
st = eng.contract_step()
m = eng.skeletal_sample(0.01, fetch=False)
e = eng.radius_graph(0.055, fetch=False)
print({"n_in": st["n_in"], "n_vox": st["n_vox"], "n_out": st["n_out"], "skeletal": m, "edges": e})
