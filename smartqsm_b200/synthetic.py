"""Seeded synthetic tree point clouds for tests and bench.

The reference ships no sample clouds, so every input here is
generated.  The generator emulates what reaches ``Skeletonization.run`` after the
entry point's pre-processing (reference ``entrypoints/smartqsm.py:882-922``):

* xy bounding-box centre moved to 0, ``z_min`` moved to 0 (``:882-888``);
* one point per occupied 1 cm voxel (``:889``; the reference keeps the voxel mean,
  we keep the first sample of every voxel — only the density matters here);
* density-adaptation scale ≈ 1 because the spacing already is ≈ 1 cm (``:911-913``);
* float64 ``N x 3`` array, point order shuffled with the same seed (the first-come
  voxelisation rule makes results order dependent).

A tree is a trunk plus recursively generated branches sampled as cylinder-surface
points ("leaf-off"); "leaf-on" adds small planar leaf patches around the outer
branches.  ``target_points`` is reached by scaling the tree until the voxelised
surface holds about that many points.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

__all__ = ["make_tree", "make_plot"]


def _segments(rng: np.random.Generator, height: float, trunk_radius: float, depth: int) -> np.ndarray:
    """Return an array [S, 8] of (x0,y0,z0, x1,y1,z1, r0, r1) tapered cylinders."""
    segs: List[Tuple[float, ...]] = []

    def grow(p0: np.ndarray, direction: np.ndarray, length: float, r0: float, level: int) -> None:
        n_piece = max(2, int(length / 0.5))
        p = p0.copy()
        d = direction / np.linalg.norm(direction)
        r = r0
        taper = 0.55 if level > 0 else 0.35
        for i in range(n_piece):
            step = length / n_piece
            d = d + rng.normal(0.0, 0.06, 3)
            d /= np.linalg.norm(d)
            q = p + d * step
            r1 = r0 * (1.0 - taper * (i + 1) / n_piece)
            segs.append((*p, *q, r, r1))
            # side branches
            if level < depth and i >= (1 if level > 0 else n_piece // 3):
                n_child = rng.poisson(1.1 if level == 0 else 0.8)
                for _ in range(n_child):
                    az = rng.uniform(0, 2 * math.pi)
                    el = rng.uniform(0.35, 1.1)
                    # build a direction tilted from d
                    a = np.cross(d, [0.0, 0.0, 1.0])
                    if np.linalg.norm(a) < 1e-6:
                        a = np.array([1.0, 0.0, 0.0])
                    a /= np.linalg.norm(a)
                    b = np.cross(d, a)
                    cd = math.cos(el) * d + math.sin(el) * (math.cos(az) * a + math.sin(az) * b)
                    cl = length * rng.uniform(0.45, 0.7) * (1.0 - 0.5 * i / n_piece)
                    cr = max(0.004, r1 * rng.uniform(0.35, 0.6))
                    if cl > 0.25:
                        grow(q, cd, cl, cr, level + 1)
            p, r = q, r1

    grow(np.zeros(3), np.array([0.0, 0.0, 1.0]), height, trunk_radius, 0)
    return np.asarray(segs, dtype=np.float64)


def _sample_cylinders(rng: np.random.Generator, segs: np.ndarray, density: float) -> np.ndarray:
    """Sample ``density`` points per m^2 of lateral surface on every tapered cylinder."""
    p0, p1 = segs[:, 0:3], segs[:, 3:6]
    r0, r1 = segs[:, 6], segs[:, 7]
    axis = p1 - p0
    length = np.linalg.norm(axis, axis=1)
    area = math.pi * (r0 + r1) * length
    counts = rng.poisson(area * density).astype(np.int64)
    total = int(counts.sum())
    seg_id = np.repeat(np.arange(len(segs)), counts)
    t = rng.random(total)
    theta = rng.random(total) * (2 * math.pi)
    d = axis[seg_id] / length[seg_id, None]
    helper = np.where(np.abs(d[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    u = np.cross(d, helper)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = np.cross(d, u)
    rad = r0[seg_id] + (r1[seg_id] - r0[seg_id]) * t
    rad = rad + rng.normal(0.0, 0.002, total)
    pts = p0[seg_id] + axis[seg_id] * t[:, None] + rad[:, None] * (np.cos(theta)[:, None] * u + np.sin(theta)[:, None] * v)
    return pts


def _sample_leaves(rng: np.random.Generator, segs: np.ndarray, n_points: int, leaf_size: float = 0.06) -> np.ndarray:
    """Planar square leaf patches scattered around thin branches (leaf-on crowns)."""
    if n_points <= 0:
        return np.zeros((0, 3))
    thin = segs[segs[:, 6] < np.quantile(segs[:, 6], 0.6)]
    per_leaf = max(8, int(leaf_size * leaf_size * 1.2e4))
    n_leaf = max(1, n_points // per_leaf)
    sid = rng.integers(0, len(thin), n_leaf)
    t = rng.random(n_leaf)[:, None]
    centre = thin[sid, 0:3] * (1 - t) + thin[sid, 3:6] * t + rng.normal(0.0, 0.12, (n_leaf, 3))
    nrm = rng.normal(size=(n_leaf, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    helper = np.where(np.abs(nrm[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    u = np.cross(nrm, helper)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = np.cross(nrm, u)
    lid = np.repeat(np.arange(n_leaf), per_leaf)
    a = (rng.random(len(lid)) - 0.5) * leaf_size
    b = (rng.random(len(lid)) - 0.5) * leaf_size
    return centre[lid] + a[:, None] * u[lid] + b[:, None] * v[lid]


def _voxel_first(points: np.ndarray, voxel: float) -> np.ndarray:
    """Keep the first point of every occupied voxel (one point per 1 cm voxel)."""
    ijk = np.floor((points - points.min(axis=0)) / voxel).astype(np.int64)
    dims = ijk.max(axis=0) + 1
    lin = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
    _, first = np.unique(lin, return_index=True)
    first.sort()
    return points[first]


def _rotate_z(segs: np.ndarray, angle: float) -> np.ndarray:
    c, s = math.cos(angle), math.sin(angle)
    out = segs.copy()
    for o in (0, 3):
        x, y = segs[:, o].copy(), segs[:, o + 1].copy()
        out[:, o], out[:, o + 1] = c * x - s * y, s * x + c * y
    return out


def _area(segs: np.ndarray) -> float:
    return float((math.pi * (segs[:, 6] + segs[:, 7]) * np.linalg.norm(segs[:, 3:6] - segs[:, 0:3], axis=1)).sum())


def make_tree(target_points: int, seed: int, leaf_on: bool = False, voxel: float = 0.01) -> np.ndarray:
    """Synthetic single-tree cloud with ``target_points`` points, float64 ``[N,3]``.

    The result is pre-processed like ``entrypoints/smartqsm.py:882-913`` would leave it
    and shuffled with ``seed``.  A nominal 12 m tree is generated, scaled (≤ 31 m),
    thickened and, for very large targets, given extra rotated copies of its crown until
    the voxelised surface holds a little more than ``target_points`` points; surplus rows
    are dropped after the shuffle.
    """
    rng = np.random.default_rng(seed)
    yield_per_m2 = 0.62e4          # occupied 1 cm voxels per m^2 at the sampling density below
    want_area = 1.25 * target_points / yield_per_m2
    leaf_frac = 0.7 if leaf_on else 0.0
    wood_area = want_area * (1.0 - leaf_frac)
    depth = 3 if target_points < 200_000 else 4
    segs = _segments(rng, 12.0, 0.16, depth)
    a0 = _area(segs)
    s = float(np.clip(math.sqrt(wood_area / a0), 0.12, 2.6))
    segs = segs * s                                     # positions and radii scale together
    k = wood_area / (a0 * s * s)
    if k > 1.0:
        thick = min(k, 2.5)
        segs[:, 6:8] *= thick
        copies = int(math.ceil(k / thick))
        if copies > 1:
            segs = np.concatenate([_rotate_z(segs, rng.uniform(0, 2 * math.pi)) if i else segs
                                   for i in range(copies)], axis=0)
    pts = [_sample_cylinders(rng, segs, 1.0e4)]
    if leaf_on:
        pts.append(_sample_leaves(rng, segs, int(1.25 * target_points * leaf_frac * 1.35)))
    cloud = _voxel_first(np.concatenate(pts, axis=0), voxel)
    tries = 0
    while len(cloud) < target_points and tries < 4:     # rarely needed safety net
        extra = cloud[rng.integers(0, len(cloud), (target_points - len(cloud)) * 2 + 1024)]
        extra = extra + rng.normal(0.0, 0.02, extra.shape)
        cloud = _voxel_first(np.concatenate([cloud, extra], axis=0), voxel)
        tries += 1
    lo, hi = cloud.min(axis=0), cloud.max(axis=0)
    cloud = cloud - np.array([(lo[0] + hi[0]) / 2.0, (lo[1] + hi[1]) / 2.0, lo[2]])
    cloud = cloud[rng.permutation(len(cloud))]
    if len(cloud) > target_points:
        cloud = cloud[:target_points]
        cloud[:, 2] -= cloud[:, 2].min()
    return np.ascontiguousarray(cloud, dtype=np.float64)


def make_plot(n_trees: int, points_per_tree: int, seed0: int = 100, leaf_on: bool = False) -> List[np.ndarray]:
    """Config C4: ``n_trees`` independent trees (seeds ``seed0 ...``), each its own cloud."""
    return [make_tree(points_per_tree, seed0 + i, leaf_on) for i in range(n_trees)]
