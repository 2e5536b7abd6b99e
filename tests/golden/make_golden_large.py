"""Golden fixtures at BASELINE-config sizes, computed once by the CPU oracle (minutes of CPU time) so that the GPU box
only has to compare:

* ``c1_iteration.npz``  - ONE ``_contract`` iteration of C1 (1 000 000-point leaf-off tree, seed 1, leaf-off checkpoint,
  reference hyper-parameters): SHA-256 digests of every integer output in a canonical form both sides can compute
  (level-0 voxels / winners / masks in canonical order, batch shapes, per level the coordinate-sorted voxel set with the
  27-bit SubM neighbour mask, parent flag and kernel offset of every row), the row counts, and every 64th row of the
  float32 outputs plus float64 column sums.
* ``run250k.npz``       - the whole stage (10 thunks, Chamfer monitor on) on a 250 000-point leaf-off tree, seed 6:
  per-thunk sizes, Chamfer values and decisions, the final state sampled every 16th row, skeletal sample / edge counts.

Run in the authoring container:  python tests/golden/make_golden_large.py   (needs baseline/_ref/*.npz checkpoints)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import contraction as OC, network as ON, weights  # noqa: E402
from smartqsm_b200.synthetic import make_tree  # noqa: E402
from tests import helpers as H  # noqa: E402


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def level_digest(coords, nbr, parent, koff) -> str:
    """Canonical form of one level: rows sorted by (sample, z, y, x); SubM table as a 27-bit presence mask (the rows the
    entries point to are implied by the voxel set), parent as a flag, kernel offset where a parent exists."""
    order = np.argsort(H.coord_key(coords), kind="stable")
    mask = np.zeros(len(coords), np.uint32)
    for k in range(27):
        mask |= (nbr[k] >= 0).astype(np.uint32) << np.uint32(k)
    has_p = (parent >= 0)
    ko = np.where(has_p, koff, -1).astype(np.int8)
    return digest(coords[order].astype(np.int32), mask[order], has_p[order].astype(np.uint8), ko[order])


def c1_iteration():
    state = weights.load_state("leafoff", allow_random=False)
    pts = make_tree(1_000_000, 1)
    p32 = pts.astype(np.float32)
    Po, Do, tr = OC.contract_once(p32, np.zeros_like(p32), ON.Weights(state), batch_size=16, keep=True)
    o_idx = np.concatenate([np.concatenate([np.full((len(v.coords_zyx), 1), si, np.int32), v.coords_zyx], axis=1)
                            for si, v in enumerate(tr.vox)])
    o_win = np.concatenate([v.winner_global for v in tr.vox]).astype(np.int64)
    o_mask = np.concatenate([v.mask for v in tr.vox]).astype(np.uint8)
    shapes = np.stack([b.spatial_shape for b in tr.batches]).astype(np.int32)
    olv = H.oracle_level_tables(tr, 16)
    lv_dig = [level_digest(L["coords"], L["nbr"], L["parent"], L["koff"]) for L in olv]
    desc = np.array([[int(x) for x in L["descended"]] for L in olv[:-1]], np.int32)
    np.savez_compressed(os.path.join(HERE, "c1_iteration.npz"),
                        n_in=len(pts), n_vox=len(o_idx), n_out=len(Po), n_cubes=len(tr.samples), far_face_rows=tr.far_face_rows,
                        level_rows=np.array([len(L["coords"]) for L in olv], np.int64),
                        voxel_digest=digest(o_idx), winner_digest=digest(o_win), mask_digest=digest(o_mask),
                        shapes=shapes, descended=desc, level_digests=np.array(lv_dig),
                        P_sample=Po[::64], D_sample=Do[::64], P_sum=Po.astype(np.float64).sum(0), D_sum=Do.astype(np.float64).sum(0),
                        D_absmax=np.abs(Do).max())
    print("c1_iteration:", len(o_idx), "voxels,", len(Po), "rows out, levels", [len(L["coords"]) for L in olv])


def run250k():
    state = weights.load_state("leafoff", allow_random=False)
    pts = make_tree(250_000, 6)
    o = OC.SpconvBasedContractionOracle(state=state, monitored_by_chamfer_distance=True, max_iteration=10, batch_size=16)
    o.input(points=pts)
    for f in o.get_pipeline():
        f()
    out = o.output()
    P, D = out["contracted_points"], out["displacements"]
    ids, sk, rad = OC.skeletal_points_and_radii(P, D, 0.01)
    edges = OC.radius_edges(sk, 0.055)
    np.savez_compressed(os.path.join(HERE, "run250k.npz"),
                        n_out=np.array([h["n_out"] for h in o.history], np.int64),
                        accepted=np.array([int(h["accepted"]) for h in o.history], np.int32),
                        chamfer=np.array([h["chamfer"] for h in o.history], np.float64),
                        P_sample=P[::16], D_sample=D[::16], n_final=len(P), n_skeletal=len(ids), n_edges=len(edges),
                        P_sum=P.astype(np.float64).sum(0), D_absmax=np.abs(D).max())
    print("run250k:", [h["n_out"] for h in o.history], [int(h["accepted"]) for h in o.history], len(ids), len(edges))


if __name__ == "__main__":
    which = sys.argv[1:] or ["run250k", "c1"]
    if "run250k" in which:
        run250k()
    if "c1" in which:
        c1_iteration()
