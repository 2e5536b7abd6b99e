import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from smartqsm_b200._lib import Engine
from smartqsm_b200.checkpoints import load_state
from smartqsm_b200.synthetic import make_tree
state = load_state("leafon")
pts = make_tree(5_000_000, 3, True)
eng = Engine(monitored_by_chamfer_distance=True)
eng.load_state_dict(state)
eng.set_points(pts)
for i in range(3):
    st = eng.contract_step()
    print(i, st["n_out"], st["accepted"], st["chamfer"], st["ms_chamfer"], flush=True)
