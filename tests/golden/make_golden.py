"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own Python.

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden.py

The reference's files (``core/core_algorithms/spconv_based_contraction.py``,
``external/SmartTreeXX/datasets/synthetic_tree.py``, ``external/SmartTreeXX/models/*.py``) are imported
unmodified through ``oracle/refshim.py``, which supplies stand-ins for the un-installable third-party
wheels (spconv, open3d).  Every array written here therefore comes out of the reference's own driver,
tiling, collation, network wiring and decode code.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import network as ON, refshim, weights  # noqa: E402
from smartqsm_b200.synthetic import make_tree  # noqa: E402


def run_reference(state, pts, n_thunks, **kw):
    ref = refshim.reference_contraction(state, max_iteration=n_thunks, **kw)
    ref.input(points=pts)
    P, D, cd = [], [], []
    for f in ref.get_pipeline():
        assert f() is None                      # non-verbose thunks return None
        P.append(ref._contracted_points.copy())
        D.append(ref._displacements.copy())
        cd.append(np.nan if ref._chamfer_distance is None else ref._chamfer_distance)
    out = ref.output()
    assert np.array_equal(out["contracted_points"], P[-1])
    return P, D, np.array(cd)


def face_cloud():
    rng = np.random.default_rng(0)
    a = rng.uniform(-1.0, 1.0, (3000, 3)) * 3.0 + np.array([0.0, 0.0, 3.0])
    a[:150, 0] = 0.0
    a[150:220, 1] = 4.0
    a[220:250, 2] = 4.0
    far = np.array([[40.0, 40.0, 40.0], [80.5, 0.25, 0.25], [80.5, 0.25, 0.26]])
    return np.concatenate([a, far])


def main():
    assert refshim.available(), "needs /root/reference"
    weights.export_reference_checkpoints()
    refshim.install()
    # 1/2: the whole driver, 4 thunks, Chamfer monitor on
    cases = [("random", ON.random_state(seed=7), make_tree(6000, 31)),
             ("leafoff", weights.load_state("leafoff", allow_random=False), make_tree(6000, 32)),
             ("faces", ON.random_state(seed=7), face_cloud())]
    for name, state, pts in cases:
        bs = 2 if name == "faces" else 16
        P, D, cd = run_reference(state, pts, 4, batch_size=bs, monitored_by_chamfer_distance=True)
        sizes = np.array([len(p) for p in P])
        np.savez_compressed(os.path.join(HERE, f"contraction_{name}.npz"), points=pts, batch_size=bs,
                            sizes=sizes, chamfer=cd, P=np.concatenate(P), D=np.concatenate(D))
        print(name, sizes, cd)
    # 3: tiling of SingleTreeDataset on the face cloud
    from external.SmartTreeXX.datasets.synthetic_tree import SingleTreeDataset
    pts = face_cloud().astype(np.float32)
    ds = SingleTreeDataset(np.concatenate([pts, np.zeros_like(pts)], axis=1), device="cpu")
    n_s = np.array([len(d[0]) for d in ds._data])
    n_in = np.array([int(d[4].sum()) for d in ds._data])
    lo = np.stack([d[2].numpy() for d in ds._data])
    hi = np.stack([d[3].numpy() for d in ds._data])
    first = np.array([float(d[0][0].sum()) for d in ds._data])
    np.savez_compressed(os.path.join(HERE, "tiling_faces.npz"), points=pts, n_sample=n_s, n_interior=n_in, slo=lo, shi=hi,
                        first_point_sum=first)
    print("tiling", n_s, n_in)
    # 4: one forward of the reference SmartTreeXX on one collated batch
    from external.SmartTreeXX.datasets.synthetic_tree import collate_synthetic_tree
    state = ON.random_state(seed=7)
    model = refshim.reference_model(state)
    tree = make_tree(4000, 33).astype(np.float32)
    ds = SingleTreeDataset(np.concatenate([tree, np.zeros_like(tree)], axis=1), device="cpu")
    batch = collate_synthetic_tree([ds[i] for i in range(min(len(ds), 16))])
    with torch.no_grad():
        out = model(batch[0])
    np.savez_compressed(os.path.join(HERE, "forward_random.npz"), points=tree, feats=batch[0].features.numpy(),
                        indices=batch[0].indices.numpy(), spatial_shape=np.array([int(s) for s in batch[0].spatial_shape]),
                        direction=out["direction"].numpy(), distance=out["distance"].numpy(), mask=batch[2].numpy())
    print("forward", out["direction"].shape)


if __name__ == "__main__":
    main()
