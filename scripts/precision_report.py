"""Per-stage error of the CUDA path against the fp32 and fp64 oracle for every conv_mode (GPU box)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import contraction as OC, network as ON, weights
from smartqsm_b200._lib import Engine
from smartqsm_b200.synthetic import make_tree
from tests import helpers as H

state = weights.load_state("leafoff")
pts = make_tree(int(sys.argv[1]) if len(sys.argv) > 1 else 30000, 11)
p32 = pts.astype(np.float32)
ref = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    P, D, tr = OC.contract_once(p32, np.zeros_like(p32), ON.Weights(state, dt), keep=True)
    ref[name] = (P, D, tr)
olv = H.oracle_level_tables(ref["f32"][2], 16)
rep = {}
for mode, label in ((1, "ffma"), (2, "tc3"), (4, "f16x3"), (5, "f16x3_tma"), (3, "tf32")):
    eng = Engine(conv_mode=mode)
    eng.load_state_dict(state)
    eng.debug_capture(True)
    eng.set_points(pts)
    eng.contract_step()
    g2o = []
    for l in range(len(olv)):
        gc, _, _, _ = eng.debug_level_tables(l, False)
        g2o.append(H.match_rows(gc, olv[l]["coords"]))
    row = {}
    for rn in ("f32", "f64"):
        tr = ref[rn][2]
        for name in ["input_conv", "enc0", "enc1", "enc2", "enc3", "dec2", "dec1", "dec0"]:
            g, lvl = eng.debug_features(name)
            o = H.oracle_features(tr, name)
            row[f"{name}/{rn}"] = H.rel_err(g, o[g2o[lvl]])
        Pg, Dg = eng.result()
        row[f"D/{rn}"] = H.rel_err(Dg, ref[rn][1])
        row[f"P_abs/{rn}"] = float(np.max(np.abs(Pg.astype(np.float64) - ref[rn][0])))
    rep[label] = row
rep["oracle_f32_vs_f64"] = {"D": H.rel_err(ref["f32"][1], ref["f64"][1]),
                            **{n: H.rel_err(H.oracle_features(ref["f32"][2], n), H.oracle_features(ref["f64"][2], n))
                               for n in ["input_conv", "enc0", "enc1", "enc2", "enc3", "dec2", "dec1", "dec0"]}}
for k, v in rep.items():
    print(k, {a: float(f"{b:.2e}") for a, b in v.items()})
