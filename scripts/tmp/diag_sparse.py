import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import contraction as OC
from oracle.weights import load_state
from smartqsm_b200.core_algorithms import SpconvBasedContraction
from smartqsm_b200.synthetic import make_tree

state = load_state("leafoff")
pts = make_tree(20000, 3)
o = OC.SpconvBasedContractionOracle(state=state, monitored_by_chamfer_distance=True, max_iteration=10)
o.input(points=pts)
for f in o.get_pipeline():
    f()
res = {}
for name, env in (("sparse", None), ("dense", "0")):
    if env is None:
        os.environ.pop("SQSM_SPARSE_LOOP", None)
    else:
        os.environ["SQSM_SPARSE_LOOP"] = env
    g = SpconvBasedContraction(state_dict=state, batch_size=16, max_iteration=10, device="cuda", monitored_by_chamfer_distance=True)
    g.input(points=pts)
    for f in g.get_pipeline():
        f()
    res[name] = (g.history, g.output())
for i, ho in enumerate(o.history):
    print(i, "oracle", ho["n_out"], ho["accepted"], "%.12e" % ho["chamfer"],
          "| sparse", res["sparse"][0][i]["n_out"], "%.12e" % res["sparse"][0][i]["chamfer"], "rel %.2e" % (abs(res["sparse"][0][i]["chamfer"] - ho["chamfer"]) / ho["chamfer"]),
          "| dense", res["dense"][0][i]["n_out"], "%.12e" % res["dense"][0][i]["chamfer"], "rel %.2e" % (abs(res["dense"][0][i]["chamfer"] - ho["chamfer"]) / ho["chamfer"]))
oo = o.output()
for name in ("sparse", "dense"):
    out = res[name][1]
    if out["contracted_points"].shape == oo["contracted_points"].shape:
        d = np.abs(out["contracted_points"] - oo["contracted_points"]).max(axis=1)
        print(name, "max|dP| vs oracle %.3e" % d.max(), "p99.9 %.3e" % np.quantile(d, 0.999), "n>1e-5:", int((d > 1e-5).sum()))
    else:
        print(name, "shape differs", out["contracted_points"].shape, oo["contracted_points"].shape)
a, b = res["sparse"][1]["contracted_points"], res["dense"][1]["contracted_points"]
if a.shape == b.shape:
    d = np.abs(a - b).max(axis=1)
    print("sparse vs dense max %.3e" % d.max(), "n>1e-5:", int((d > 1e-5).sum()))
