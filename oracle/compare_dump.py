"""Oracle (test infrastructure): compare a dump of the GENUINE reference (``scripts/reference_dump.py``: real spconv +
Open3D) with this restatement - the hook that closes "parity unpinned against the third-party wheels" (SURVEY.md App. H
consequence 2).

``python -m oracle.compare_dump spconv_dump_x.npz`` prints, per recorded item, whether the restatement reproduces it:

* voxel sets per batch (coordinates as a set, and the ORDER: the CPU wheel creates voxels in first-appearance order, the GPU
  wheel in atomic order - App. B item 4 - so only the set is required to match);
* every sparse convolution output, rows matched by coordinate, for BOTH settings of ``DOWN_KERNEL_FLIP`` (the one
  convention the survey could not discriminate);
* direction / distance, the state after every thunk and the Chamfer values.
"""
from __future__ import annotations

import sys
from typing import Dict, List

import numpy as np

from . import contraction as OC
from . import network as ON
from . import tiling as OT


def _key(idx: np.ndarray) -> np.ndarray:
    i = idx.astype(np.int64)
    return ((i[:, 0] * 1024 + i[:, 1]) * 1024 + i[:, 2]) * 1024 + i[:, 3]


def _match(a_idx: np.ndarray, b_idx: np.ndarray):
    """rows of b in the order of a (by coordinate); None when the voxel sets differ"""
    ka, kb = _key(a_idx), _key(b_idx)
    if len(ka) != len(kb):
        return None
    oa, ob = np.argsort(ka, kind="stable"), np.argsort(kb, kind="stable")
    if not np.array_equal(ka[oa], kb[ob]):
        return None
    m = np.empty(len(ka), np.int64)
    m[oa] = ob
    return m


def dump_state(z) -> Dict[str, np.ndarray]:
    """The weights embedded in the dump, or the named shipped checkpoint (``--no-weights NAME``)."""
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    if not state:
        from . import weights
        state = weights.load_state(str(z["ckpt_name"]), allow_random=False)
    return state


def compare(path: str, verbose: bool = True) -> Dict[str, object]:
    z = np.load(path, allow_pickle=False)
    state = dump_state(z)
    pts = z["points"]
    bs, thunks, nb = int(z["batch_size"]), int(z["thunks"]), int(z["n_batches"])
    report: Dict[str, object] = {"kind": str(z["kind"]), "items": [], "ok": True}

    def item(name, ok, detail=""):
        report["items"].append((name, bool(ok), detail))
        report["ok"] = bool(report["ok"] and ok)
        if verbose:
            print(f"{'ok  ' if ok else 'DIFF'} {name} {detail}")

    p32 = pts.astype(np.float32)
    d32 = np.zeros_like(p32)
    results = {}
    for flip in (False, True):
        ON.DOWN_KERNEL_FLIP = flip
        try:
            Po, Do, tr = OC.contract_once(p32, d32, ON.Weights(state), batch_size=bs, keep=True)
        finally:
            ON.DOWN_KERNEL_FLIP = False
        results[flip] = (Po, Do, tr)
    tr = results[False][2]
    # ---- voxels per batch
    for b in range(nb):
        d_idx = z[f"b{b}:indices"]
        o_idx = tr.batches[b].indices
        m = _match(d_idx, o_idx)
        item(f"batch {b}: voxel set", m is not None, f"({len(d_idx)} vs {len(o_idx)} voxels)")
        if m is None:
            continue
        item(f"batch {b}: voxel order (first appearance, CPU wheel only)", np.array_equal(m, np.arange(len(m))))
        item(f"batch {b}: voxel features (kept point per voxel)", np.array_equal(z[f"b{b}:features"], tr.batches[b].feats[m]))
        item(f"batch {b}: spatial_shape", np.array_equal(z[f"b{b}:spatial_shape"], np.asarray(tr.batches[b].spatial_shape, np.int32)))
    # ---- layers, both kernel-index conventions of the strided / inverse convolutions
    layer_keys = sorted({k.split(":")[2] for k in z.files if ":layer:" in k})
    worst = {False: 0.0, True: 0.0}
    for flip in (False, True):
        trf = results[flip][2]
        for b in range(nb):
            taps, tidx = trf.fwd[b].taps, trf.fwd[b].tap_indices
            for name in layer_keys:
                fk = f"b{b}:layer:{name}:features"
                if fk not in z.files:
                    continue
                if name not in taps:
                    item(f"batch {b}: layer {name} [flip={flip}]", False, "(the restatement did not execute this layer)")
                    continue
                m = _match(z[f"b{b}:layer:{name}:indices"], tidx[name])
                if m is None:
                    item(f"batch {b}: layer {name} indices [flip={flip}]", False)
                    continue
                d, o = z[fk], taps[name].numpy()[m]
                err = float(np.max(np.abs(d - o)) / max(float(np.max(np.abs(o))), 1e-30))
                worst[flip] = max(worst[flip], err)
                if not flip:
                    item(f"batch {b}: layer {name}", err <= 1e-4, f"rel err {err:.2e}")
    report["layer_rel_err"] = {"DOWN_KERNEL_FLIP=False": worst[False], "DOWN_KERNEL_FLIP=True": worst[True]}
    if verbose:
        print(f"worst layer error: flip=False {worst[False]:.2e}, flip=True {worst[True]:.2e} -> "
              f"{'the default convention fits' if worst[False] <= worst[True] else 'DOWN_KERNEL_FLIP=True fits better: flip it'}")
    # ---- heads and the state per thunk
    for b in range(nb):
        dk = f"b{b}:direction"
        if dk in z.files:
            m = _match(z[f"b{b}:indices"], tr.batches[b].indices)
            if m is not None:
                direction, distance, _ = ON.forward(tr.batches[b].feats, tr.batches[b].indices, tr.batches[b].spatial_shape, ON.Weights(state))
                e1 = float(np.max(np.abs(z[dk] - direction.numpy()[m])))
                e2 = float(np.max(np.abs(z[f"b{b}:distance"] - distance.numpy()[m])))
                item(f"batch {b}: direction / distance", e1 <= 1e-4 and e2 <= 1e-5 + 1e-4 * float(np.max(np.abs(z[f'b{b}:distance']))), f"abs err {e1:.2e} / {e2:.2e}")
    o = OC.SpconvBasedContractionOracle(state=state, monitored_by_chamfer_distance=True, max_iteration=thunks, batch_size=bs)
    o.input(points=pts)
    for t, f in enumerate(o.get_pipeline()):
        f()
        P = o.output()["contracted_points"]
        dP = z[f"P{t}"]
        same_shape = P.shape == dP.shape
        item(f"thunk {t}: state size", same_shape, f"({len(dP)} vs {len(P)})")
        if same_shape:
            # the GPU wheel orders voxels differently: compare as multisets of rows
            a = dP[np.lexsort(dP.T[::-1])]
            b_ = P[np.lexsort(P.T[::-1])]
            item(f"thunk {t}: contracted points", float(np.max(np.abs(a - b_))) <= 1e-4, f"max abs {float(np.max(np.abs(a - b_))):.2e}")
        cd_d = float(z[f"chamfer{t}"])
        cd_o = o._chamfer_distance
        item(f"thunk {t}: Chamfer value", abs(cd_d - cd_o) <= 1e-4 * abs(cd_o), f"{cd_d:.6e} vs {cd_o:.6e}")
    return report


if __name__ == "__main__":
    r = compare(sys.argv[1])
    print("ALL ITEMS REPRODUCED" if r["ok"] else "DIFFERENCES FOUND (see DIFF lines)")
    sys.exit(0 if r["ok"] else 1)
