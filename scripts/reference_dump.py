#!/usr/bin/env python
"""Export what the GENUINE reference computes on this path, for comparison with the oracle / the B200 engine.

Run inside a working SmartQSM environment (real ``spconv`` + ``open3d`` installed; CPU or GPU), from the SmartQSM
checkout or with ``--smartqsm-root``:

    python reference_dump.py --points cloud.npy --ckpt ckpts/for_leafoff_seasons.pth --out spconv_dump_mytree.npz \
        [--device cpu] [--thunks 3] [--batch-size 16]

``cloud.npy`` is the N x 3 float64 array that would reach ``Skeletonization.run`` (after the entry point's
pre-processing).  Without ``--points`` a small deterministic cloud is generated (``--n``, ``--seed``).

What is recorded (the reference's own call sites, nothing is patched except for observation):

* per cube sample of the FIRST iteration: the voxel indices ``PointToVoxel.generate_voxel_with_id`` produced, as they
  leave ``SingleTreeDataset.__getitem__`` (external/SmartTreeXX/datasets/synthetic_tree.py:308-333);
* per batch of the first iteration: the output (features, indices) of EVERY ``SubMConv3d`` / ``SparseConv3d`` /
  ``SparseInverseConv3d`` of ``SmartTreeXX`` (models/smart_tree_xx.py:21-38, sparse_module.py) through forward hooks,
  plus ``direction`` and ``distance`` (:75-76);
* per thunk: ``_contracted_points``, ``_displacements`` and the Chamfer value
  (core/core_algorithms/spconv_based_contraction.py:112-131).

``oracle/compare_dump.py`` ingests the file: ``python -m oracle.compare_dump spconv_dump_mytree.npz``; dropped into
``tests/golden/`` as ``spconv_dump_*.npz`` it is picked up by the test suite (CPU: oracle vs dump, GPU: engine vs dump).
This closes the "parity unpinned against the third-party wheels" gap of DESIGN.md section 2 for whoever owns such an
environment.  ``--shim`` runs the same script over the oracle's stand-in modules (authoring container: tests the tooling).
"""
import argparse
import os
import sys

import numpy as np


def default_cloud(n: int, seed: int) -> np.ndarray:
    """A small deterministic branch-like cloud (no dependency on this repository): noisy cylinders on a 1 cm lattice."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 1, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    r = 0.06 * (1.2 - t)
    trunk = np.stack([r * np.cos(ang), r * np.sin(ang), 3.0 * t], 1)
    k = n // 3
    br = np.stack([0.5 * t[:k] + 0.02 * np.cos(ang[:k]), 0.02 * np.sin(ang[:k]) + 0.1 * t[:k], 1.5 + 0.8 * t[:k]], 1)
    pts = np.concatenate([trunk, br]) + rng.normal(0, 0.002, (n + k, 3))
    ijk = np.floor((pts - pts.min(0)) / 0.01).astype(np.int64)
    _, first = np.unique(ijk[:, 0] * 1_000_000 + ijk[:, 1] * 1000 + ijk[:, 2], return_index=True)
    pts = pts[np.sort(first)]
    pts[:, 2] -= pts[:, 2].min()
    return np.ascontiguousarray(pts, dtype=np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="")
    ap.add_argument("--n", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ckpt", default="ckpts/for_leafoff_seasons.pth")
    ap.add_argument("--out", default="spconv_dump.npz")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--thunks", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--smartqsm-root", default=os.environ.get("SMARTQSM_ROOT", ""))
    ap.add_argument("--shim", action="store_true", help="authoring container: run over oracle/refshim stand-ins")
    ap.add_argument("--state-npz", default="", help="with --shim: checkpoint blob (.npz) instead of --ckpt")
    ap.add_argument("--no-weights", default="", metavar="NAME",
                    help="do not embed the weights; record this checkpoint name (leafoff / leafon) instead")
    args = ap.parse_args()

    pts = np.load(args.points) if args.points else default_cloud(args.n, args.seed)
    pts = np.ascontiguousarray(pts, dtype=np.float64)

    import torch
    if args.shim:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import refshim
        refshim.install()
        with np.load(args.state_npz) as z:
            state = {k: z[k] for k in z.files}
        algo = refshim.reference_contraction(state, max_iteration=args.thunks, batch_size=args.batch_size,
                                             monitored_by_chamfer_distance=True)
        kind = "shim (oracle stand-ins for spconv / open3d): NOT a genuine run"
    else:
        if args.smartqsm_root:
            sys.path.insert(0, args.smartqsm_root)
        import spconv  # noqa: F401  (must be the real wheel)
        from core.core_algorithms.spconv_based_contraction import SpconvBasedContraction
        algo = SpconvBasedContraction(ckpt_path=args.ckpt, batch_size=args.batch_size, max_iteration=args.thunks,
                                      device=args.device, monitored_by_chamfer_distance=True)
        kind = f"genuine: spconv {getattr(spconv, '__version__', '?')}, torch {torch.__version__}, device {args.device}"
    model = algo._model
    state_out = {"w:" + k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if not k.endswith("num_batches_tracked")}

    # ---- observation hooks
    rec = {"layers": [], "batch_of_layer": [], "heads": []}
    conv_names = {"SubMConv3d", "SparseConv3d", "SparseInverseConv3d"}
    live = {"on": True, "batch": -1}

    def conv_hook(name):
        def hook(mod, inp, out):
            if live["on"]:
                rec["layers"].append((live["batch"], name, out.features.detach().cpu().numpy().astype(np.float32),
                                      out.indices.detach().cpu().numpy().astype(np.int32)))
        return hook

    for name, mod in model.named_modules():
        if type(mod).__name__ in conv_names:
            mod.register_forward_hook(conv_hook(name))

    def model_pre_hook(mod, inp):
        if live["on"]:
            live["batch"] += 1
            x = inp[0]
            rec["heads"].append({"indices": x.indices.detach().cpu().numpy().astype(np.int32),
                                 "features": x.features.detach().cpu().numpy().astype(np.float32),
                                 "spatial_shape": np.asarray([int(s) for s in x.spatial_shape], np.int32)})

    def model_hook(mod, inp, out):
        if live["on"]:
            rec["heads"][-1]["direction"] = out["direction"].detach().cpu().numpy().astype(np.float32)
            rec["heads"][-1]["distance"] = out["distance"].detach().cpu().numpy().astype(np.float32)

    model.register_forward_pre_hook(model_pre_hook)
    model.register_forward_hook(model_hook)

    algo.input(points=pts)
    out = {"points": pts, "kind": np.array(kind), "batch_size": np.int32(args.batch_size), "thunks": np.int32(args.thunks)}
    if args.no_weights:
        out["ckpt_name"] = np.array(args.no_weights)
    else:
        out.update(state_out)
    for t, thunk in enumerate(algo.get_pipeline()):
        thunk()
        live["on"] = False                                   # layer outputs of the first iteration only
        out[f"P{t}"] = np.asarray(algo._contracted_points, np.float32)
        out[f"D{t}"] = np.asarray(algo._displacements, np.float32)
        out[f"chamfer{t}"] = np.float64(np.nan if algo._chamfer_distance is None else algo._chamfer_distance)
    for b, h in enumerate(rec["heads"]):
        for k, v in h.items():
            out[f"b{b}:{k}"] = v
    for i, (b, name, feats, idx) in enumerate(rec["layers"]):
        out[f"b{b}:layer:{name}:features"] = feats
        out[f"b{b}:layer:{name}:indices"] = idx
    out["n_batches"] = np.int32(len(rec["heads"]))
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {len(pts)} points, {len(rec['heads'])} batches, {len(rec['layers'])} layer outputs, {args.thunks} thunks; {kind}")


if __name__ == "__main__":
    main()
