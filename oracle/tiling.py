"""Oracle (test infrastructure): cube tiling, per-cube voxelisation and batch collation.

Restates ``external/SmartTreeXX/datasets/synthetic_tree.py:213-333`` (``SingleTreeDataset``)
and ``:176-208`` (``collate_synthetic_tree``) plus the CPU path of spconv's ``PointToVoxel``
(un-vendored).  See ``oracle/__init__.py`` for the parity statement.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch


@dataclass
class CubeSample:
    """One 4.8 m cube sample (``synthetic_tree.py:252-281``)."""
    cube: np.ndarray          # int64[3] cube index (x, y, z)
    point_ids: np.ndarray     # int64[n_s] rows of the input cloud inside the halo box, ascending
    lo: np.ndarray            # f32[3] cube min corner
    hi: np.ndarray            # f32[3] cube max corner
    slo: np.ndarray           # f32[3] halo box min corner
    shi: np.ndarray           # f32[3] halo box max corner
    mask: np.ndarray          # bool[n_s] interior mask (inside [lo, hi], both ends inclusive)


def tile_cubes(points_f32: np.ndarray, cube_size: float = 4.0, buffer_size: float = 0.4,
               min_num_points: int = 1) -> List[CubeSample]:
    """``SingleTreeDataset.__init__`` (``synthetic_tree.py:243-281``).

    Uses the same torch CPU operators as the reference so that every float32 rounding
    (``div(floor)``, ``idx*size + size/2``, ``centre -/+ size/2``, ``-/+ buffer``) is identical.
    """
    pts = torch.from_numpy(np.ascontiguousarray(points_f32, dtype=np.float32))
    cube_idx = torch.div(pts, cube_size, rounding_mode="floor")                    # :246
    cube_idx, counts = torch.unique(cube_idx, return_counts=True, dim=0)          # :248 (lexicographic x,y,z)
    centres = cube_idx * cube_size + cube_size / 2                                # :249
    out: List[CubeSample] = []
    for i in range(len(centres)):
        if counts[i] < min_num_points:                                            # :253
            continue
        c = centres[i]
        lo = c - cube_size / 2                                                    # :259
        hi = c + cube_size / 2                                                    # :260
        slo = lo - buffer_size                                                    # :261
        shi = hi + buffer_size                                                    # :262
        m = torch.logical_and(pts >= slo, pts <= shi).all(dim=1)                  # :263-266
        ids = torch.nonzero(m).squeeze(1)
        sp = pts[ids]
        inner = torch.logical_and(sp >= lo, sp <= hi).all(dim=1)                  # :270-272
        out.append(CubeSample(cube_idx[i].numpy().astype(np.int64), ids.numpy(), lo.numpy(), hi.numpy(),
                              slo.numpy(), shi.numpy(), inner.numpy()))
    return out


def grid_size_of(slo: np.ndarray, shi: np.ndarray, vsize: float) -> np.ndarray:
    """spconv ``Point2VoxelCPU.calc_meta_data``: ``round((hi - lo) / vsize)`` in float32 (App. C)."""
    v = np.float32(vsize)
    return np.round((shi.astype(np.float32) - slo.astype(np.float32)) / v).astype(np.int64)


@dataclass
class VoxelisedSample:
    """Result of ``SingleTreeDataset.__getitem__`` (``synthetic_tree.py:290-333``)."""
    winner_local: np.ndarray  # int64[n_v] sample-local index of the kept point, ascending (= voxel order)
    winner_global: np.ndarray # int64[n_v] row of the input cloud
    coords_zyx: np.ndarray    # int32[n_v,3] voxel index (z, y, x)
    mask: np.ndarray          # bool[n_v]  interior mask gathered by float32(index) (F7)
    gather_global: np.ndarray # int64[n_v] input row whose displacement/mask is gathered (== winner_global below 2**24)


def voxelise_sample(points_f32: np.ndarray, s: CubeSample, vsize: float = 0.01) -> VoxelisedSample:
    """CPU ``PointToVoxel(...).generate_voxel_with_id`` with ``max_num_points_per_voxel=1``
    (``synthetic_tree.py:304-324``).

    ``c_a = floor((p_a - slo_a) / vsize)`` in float32 with a true division; points with
    ``c_a`` outside ``[0, grid)`` are dropped; first point to hit a voxel creates it and is the
    only one kept; voxel order = first-appearance order; coordinates stored as ``(z, y, x)``.
    """
    p = points_f32[s.point_ids].astype(np.float32, copy=False)
    v = np.float32(vsize)
    grid = grid_size_of(s.slo, s.shi, vsize)
    c = np.floor((p - s.slo.astype(np.float32)[None, :]) / v)                      # float32 arithmetic
    ok = np.all((c >= 0) & (c < grid[None, :].astype(np.float32)), axis=1)
    local = np.nonzero(ok)[0]
    ci = c[ok].astype(np.int64)
    lin = (ci[:, 2] * grid[1] + ci[:, 1]) * grid[0] + ci[:, 0]                     # (z, y, x) row-major
    _, first = np.unique(lin, return_index=True)
    first.sort()                                                                   # first-appearance order
    win_local = local[first]
    coords = ci[first][:, ::-1].astype(np.int32)                                   # (z, y, x)
    # the reference carries the index as a float32 feature column (:304-305, :320-324)
    gather_local = win_local.astype(np.float32).astype(np.int64)
    return VoxelisedSample(win_local, s.point_ids[win_local], coords, s.mask[gather_local],
                           s.point_ids[gather_local])


@dataclass
class Batch:
    """``collate_synthetic_tree`` output (``synthetic_tree.py:176-208``) for ≤ batch_size cubes."""
    feats: np.ndarray         # f32[N0,3] xyz of the kept points (network input features)
    disp: np.ndarray          # f32[N0,3] accumulated displacement gathered per voxel
    indices: np.ndarray       # int32[N0,4] (batch id, z, y, x)
    mask: np.ndarray          # bool[N0]
    spatial_shape: np.ndarray # int64[3] = max(indices)[1:] (F6: max index, not max+1)
    winner_global: np.ndarray # int64[N0]
    sample_of_row: np.ndarray # int32[N0] global cube-sample number


def collate(points_f32: np.ndarray, disp_f32: np.ndarray, samples: List[CubeSample], first_sample: int,
            vox: List[VoxelisedSample]) -> Batch:
    feats, disp, idx, mask, wg, sid = [], [], [], [], [], []
    for b, (s, v) in enumerate(zip(samples, vox)):
        feats.append(points_f32[v.winner_global])
        disp.append(disp_f32[v.gather_global])
        idx.append(np.concatenate([np.full((len(v.coords_zyx), 1), b, np.int32), v.coords_zyx], axis=1))
        mask.append(v.mask)
        wg.append(v.winner_global)
        sid.append(np.full(len(v.coords_zyx), first_sample + b, np.int32))
    indices = np.concatenate(idx)
    return Batch(np.concatenate(feats), np.concatenate(disp), indices, np.concatenate(mask),
                 indices.max(axis=0)[1:].astype(np.int64), np.concatenate(wg), np.concatenate(sid))
