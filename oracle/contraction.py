"""Oracle (test infrastructure): the iterative contraction driver, Chamfer monitor,
skeletal sampling, radii and the radius-neighbour edge set.

Restates ``core/core_algorithms/spconv_based_contraction.py:75-151`` (``_contract``),
``core/thinning_based_skeletonization_pipeline.py:68-80`` and
``core/skeletonization_pipeline_base.py:132-144,158`` with Open3D's
``voxel_down_sample_and_trace`` / ``compute_point_cloud_distance`` restated from their
published algorithm.  See ``oracle/__init__.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from scipy.spatial import cKDTree

from . import network as net
from . import tiling


# ----------------------------------------------------------------------- Open3D restatements

def voxel_trace(points_f64: np.ndarray, voxel: float, min_bound: np.ndarray):
    """Open3D ``PointCloud.voxel_down_sample_and_trace(voxel, min_bound, max_bound)`` (App. C).

    voxel index = ``floor((p - min_bound) / voxel)`` in float64 (no half-voxel shift); returns
    ``(mean point per voxel f64[V,3], first original index per voxel int64[V])`` with voxels in
    ascending first-index order (Open3D's own order is its unordered_map iteration order and is
    not reproducible; App. D.2 compares sets / sorted arrays).
    """
    p = np.asarray(points_f64, dtype=np.float64)
    ref = (p - np.asarray(min_bound, np.float64)[None, :]) / float(voxel)
    ijk = np.floor(ref).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    lin = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
    uniq, first, inv, cnt = np.unique(lin, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    vid = rank[inv]
    mean = np.stack([np.bincount(vid, weights=p[:, a], minlength=len(uniq)) for a in range(3)], axis=1)
    mean /= cnt[order][:, None].astype(np.float64)
    return mean, first[order]


def chamfer(P_f32: np.ndarray, Q_f32: np.ndarray, voxel: float) -> float:
    """Chamfer monitor value (``spconv_based_contraction.py:112-122``)."""
    P = P_f32.astype(np.float64)
    Q = Q_f32.astype(np.float64)
    min_bound = np.minimum(P.min(axis=0), Q.min(axis=0))                           # :116
    Pv, _ = voxel_trace(P, voxel, min_bound)                                       # :120
    Qv, _ = voxel_trace(Q, voxel, min_bound)                                       # :121
    dpq = cKDTree(Qv).query(Pv, k=1)[0]
    dqp = cKDTree(Pv).query(Qv, k=1)[0]
    return float(np.mean(np.square(dpq)) + np.mean(np.square(dqp)))               # :122


# ----------------------------------------------------------------------- one iteration

@dataclass
class IterationTrace:
    """Integer by-products of one ``_contract`` call for parity tests."""
    samples: List[tiling.CubeSample] = field(default_factory=list)
    vox: List[tiling.VoxelisedSample] = field(default_factory=list)
    batches: List[tiling.Batch] = field(default_factory=list)
    fwd: List[net.ForwardTrace] = field(default_factory=list)
    n_in: int = 0
    n_vox: int = 0
    n_out: int = 0
    far_face_rows: int = 0


def contract_once(points_f32: np.ndarray, disp_f32: np.ndarray, w: net.Weights, *, batch_size: int = 16,
                  voxel_size: float = 0.01, cube_size: float = 4.0, buffer_size: float = 0.4,
                  keep: bool = False) -> Tuple[np.ndarray, np.ndarray, IterationTrace]:
    """The batch loop of ``_contract`` (``spconv_based_contraction.py:85-110``)."""
    tr = IterationTrace(n_in=len(points_f32))
    samples = tiling.tile_cubes(points_f32, cube_size, buffer_size)                # :85-91
    vox = [tiling.voxelise_sample(points_f32, s, voxel_size) for s in samples]
    out_p: List[np.ndarray] = []
    out_d: List[np.ndarray] = []
    for b0 in range(0, len(samples), batch_size):                                  # :92-100
        batch = tiling.collate(points_f32, disp_f32, samples[b0:b0 + batch_size], b0, vox[b0:b0 + batch_size])
        direction, distance, ft = net.forward(batch.feats, batch.indices, batch.spatial_shape, w, keep=keep)
        pred = (direction * distance).to(torch.float32).numpy()                    # :103
        m = batch.mask
        out_p.append((batch.feats[m] + pred[m]).astype(np.float32))                # :105-107
        out_d.append((pred[m] + batch.disp[m]).astype(np.float32))                 # :108
        tr.n_vox += len(batch.feats)
        zyx = batch.indices[:, 1:4].astype(np.int64)
        tr.far_face_rows += int(np.any(zyx >= batch.spatial_shape[None, :], axis=1).sum())
        if keep:
            tr.batches.append(batch)
            tr.fwd.append(ft)
    if keep:
        tr.samples, tr.vox = samples, vox
    P = np.concatenate(out_p, axis=0) if out_p else np.zeros((0, 3), np.float32)   # :109
    D = np.concatenate(out_d, axis=0) if out_d else np.zeros((0, 3), np.float32)   # :110
    tr.n_out = len(P)
    return P, D, tr


class SpconvBasedContractionOracle:
    """Restated ``SpconvBasedContraction`` (``spconv_based_contraction.py:34-151``), non-verbose."""

    def __init__(self, *, state: Dict[str, np.ndarray], batch_size: int = 16, max_iteration: int = 10,
                 voxel_size: float = 0.01, cube_size: float = 4.0, buffer_size: float = 0.4,
                 monitored_by_chamfer_distance: bool = False, dtype=torch.float32, latch: bool = False) -> None:
        self._w = net.Weights(state, dtype)
        self._batch_size, self._max_iteration = batch_size, max_iteration
        self._voxel_size, self._cube_size, self._buffer_size = voxel_size, cube_size, buffer_size
        self._chamfer_distance: Optional[float] = float("inf") if monitored_by_chamfer_distance else None
        self._epoch = 0
        self._latch = latch                 # F10: the reference never latches in non-verbose mode
        self._early_stopped = False
        self.history: List[dict] = []
        self._points = None
        self._contracted_points = None
        self._displacements = None

    def input(self, *, points: np.ndarray) -> None:
        self._points = points

    def _contract(self) -> None:
        self._epoch += 1
        if self._early_stopped:                                                    # :77-80
            self.history.append({"epoch": self._epoch, "skipped": True})
            return
        if self._epoch == 1:                                                       # :81-83
            self._contracted_points = np.copy(self._points).astype(np.float32)
            self._displacements = np.zeros_like(self._points).astype(np.float32)
        P, D, tr = contract_once(self._contracted_points, self._displacements, self._w,
                                 batch_size=self._batch_size, voxel_size=self._voxel_size,
                                 cube_size=self._cube_size, buffer_size=self._buffer_size)
        rec = {"epoch": self._epoch, "n_in": tr.n_in, "n_vox": tr.n_vox, "n_out": tr.n_out,
               "far_face_rows": tr.far_face_rows, "accepted": True, "chamfer": None}
        if self._chamfer_distance is not None:                                     # :112-128
            cd = chamfer(self._contracted_points, P, self._voxel_size)
            rec["chamfer"] = cd
            if cd > self._chamfer_distance:
                rec["accepted"] = False
                if self._latch:
                    self._early_stopped = True
                self.history.append(rec)
                return
            self._chamfer_distance = cd
        self._contracted_points, self._displacements = P, D                        # :130-131
        self.history.append(rec)

    def get_pipeline(self):
        return [self._contract] * self._max_iteration                              # :142-145

    def output(self) -> Dict[str, np.ndarray]:
        return {"contracted_points": self._contracted_points, "displacements": self._displacements}


# ----------------------------------------------------------------------- after the last iteration

def skeletal_points_and_radii(contracted_f32: np.ndarray, disp_f32: np.ndarray, voxel: float = 0.01):
    """``_produce_skeletal_points_and_radii`` (``thinning_based_skeletonization_pipeline.py:68-80``).

    Returns ``(sample_ids int64[M] ascending, skeletal_points f32[M,3], radii f32[M])``; node
    numbering = ascending ``sample_id`` (App. D.2).
    """
    P = contracted_f32.astype(np.float64)
    _, first = voxel_trace(P, voxel, P.min(axis=0))                                # :76
    ids = np.sort(first)                                                           # :77
    sk = contracted_f32[ids]                                                       # :78
    radii = np.linalg.norm(disp_f32[ids], axis=1)                                  # :79-80
    return ids, sk, radii


def radius_edges(skeletal_f32: np.ndarray, r: float = 0.055) -> np.ndarray:
    """Radius-neighbour edge set (``skeletonization_pipeline_base.py:132-144`` then ``:158``):
    all ``i < j`` with ``||p_i - p_j|| <= r``, lexicographically sorted ``int64[E,2]``.
    Uses SciPy's own ``query_ball_point`` exactly like the reference does."""
    tree = cKDTree(skeletal_f32)                                                   # :132 (float64 inside)
    nb = tree.query_ball_point(skeletal_f32, r=r)                                  # :136-139
    edges = [(i, j) for i, js in enumerate(nb) for j in js if j != i]
    if not edges:
        return np.zeros((0, 2), np.int64)
    return np.unique(np.sort(np.array(edges, dtype=np.int64), axis=1), axis=0)     # :158
