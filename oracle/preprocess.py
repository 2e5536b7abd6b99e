"""CPU oracle for row f2 of SURVEY.md §8 - the cloud preparation in front of the contraction path
(``entrypoints/smartqsm.py:882-913``).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

PARITY UNPINNED for the Open3D calls (``translate``, ``voxel_down_sample``,
``remove_statistical_outlier``, ``remove_radius_outlier``): Open3D is neither vendored nor installable
here, so their published algorithms are restated in float64 NumPy; the density scale uses SciPy's
``KDTree`` exactly like the reference (``smartqsm.py:911-912``).

Canonical order.  Open3D emits down-sampled voxels in the iteration order of a ``std::unordered_map``
(implementation defined, not reproducible).  The contract here: voxels ascend by
``(bz, by, bx, lz, ly, lx)`` with ``b = index >> 3`` and ``l = index & 7`` of the non-negative voxel
index - 8^3-voxel bricks in key order, then z, y, x inside the brick - which keeps neighbours close
in memory for everything downstream.  The filters preserve the order of their input
(``select_by_index``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
from scipy.spatial import cKDTree

__all__ = ["global_shift", "voxel_down_sample", "remove_statistical_outlier", "remove_radius_outlier",
           "density_scale", "prepare_cloud", "PreparedCloud"]


def global_shift(points: np.ndarray) -> np.ndarray:
    """``smartqsm.py:882-887``: xy centre of the bounding box to the origin, lowest point to z = 0."""
    p = np.asarray(points, np.float64)
    mn, mx = p.min(axis=0), p.max(axis=0)
    return np.array([-(mn[0] + mx[0]) / 2.0, -(mn[1] + mx[1]) / 2.0, -mn[2]], dtype=np.float64)


def voxel_down_sample(points: np.ndarray, voxel: float) -> Tuple[np.ndarray, np.ndarray]:
    """Open3D ``PointCloud.voxel_down_sample`` (``smartqsm.py:889``).

    ``index = floor((p - (min_bound - voxel/2)) / voxel)``; every voxel yields the float64 mean of its
    points, summed in input order and divided by the count.  Returns ``(means f64[V,3], index i64[V,3])``
    in the canonical order of the module docstring.
    """
    p = np.asarray(points, np.float64)
    if len(p) == 0:
        return np.zeros((0, 3)), np.zeros((0, 3), np.int64)
    vmin = p.min(axis=0) - float(voxel) * 0.5
    idx = np.floor((p - vmin[None, :]) / float(voxel)).astype(np.int64)
    b, l = idx >> 3, idx & 7
    key = np.stack([b[:, 2], b[:, 1], b[:, 0], l[:, 2], l[:, 1], l[:, 0]], axis=1)
    uniq, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)   # lexicographic = canonical
    inv = inv.reshape(-1)
    mean = np.stack([np.bincount(inv, weights=p[:, a], minlength=len(uniq)) for a in range(3)], axis=1)
    mean /= cnt[:, None].astype(np.float64)
    vidx = np.stack([uniq[:, 2] * 8 + uniq[:, 5], uniq[:, 1] * 8 + uniq[:, 4], uniq[:, 0] * 8 + uniq[:, 3]], axis=1)
    return mean, vidx


def remove_statistical_outlier(points: np.ndarray, nb_neighbors: int, std_ratio: float):
    """Open3D ``remove_statistical_outlier`` (``smartqsm.py:895-898``).

    Per point: mean of the distances to its ``nb_neighbors`` nearest points, the point itself (distance 0)
    included; cloud mean and sample standard deviation of those values; keep ``0 < d_i < mean + ratio*std``.
    Returns ``(kept index i64[K], avg f64[N], (mean, std, threshold))``.
    """
    p = np.asarray(points, np.float64)
    n = len(p)
    k = min(int(nb_neighbors), n)
    d = cKDTree(p).query(p, k=k)[0].reshape(n, k)
    avg = d.sum(axis=1) / float(k)
    valid = avg > 0
    nv = n                                   # every point finds at least itself
    cloud_mean = float(avg[valid].sum() / nv)
    sq = float(np.square(avg[valid] - cloud_mean).sum())
    std = float(np.sqrt(sq / (nv - 1))) if nv > 1 else 0.0
    thr = cloud_mean + float(std_ratio) * std
    keep = np.nonzero(valid & (avg < thr))[0].astype(np.int64)
    return keep, avg, (cloud_mean, std, thr)


def remove_radius_outlier(points: np.ndarray, nb_points: int, radius: float):
    """Open3D ``remove_radius_outlier`` (``smartqsm.py:899-902``): ``SearchRadius`` hands ``radius**2`` to nanoflann,
    whose ``RadiusResultSet`` keeps ``dist2 < radius2`` (open ball, the point itself included); a point stays when it has
    MORE than ``nb_points`` such neighbours.  Returns ``(kept index i64[K], count i64[N])``."""
    p = np.asarray(points, np.float64)
    r2 = float(radius) * float(radius)
    lists = cKDTree(p).query_ball_point(p, float(radius) * (1.0 + 1e-9))          # superset; exact test below
    cnt = np.zeros(len(p), np.int64)
    for i, nb in enumerate(lists):
        d = p[nb] - p[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        cnt[i] = int(np.count_nonzero(d2 < r2))
    return np.nonzero(cnt > int(nb_points))[0].astype(np.int64), cnt


def density_scale(points: np.ndarray, voxel: float) -> float:
    """``smartqsm.py:911-912``: ``voxel_down_size / mean(distance to the nearest other point)``."""
    p = np.asarray(points, np.float64)
    return float(voxel) / float(cKDTree(p).query(p, k=2)[0][:, 1].mean())


@dataclass
class PreparedCloud:
    points: np.ndarray            # what the reference holds in `points` (shifted, down-sampled, filtered), float64
    global_shift: np.ndarray
    scale: float
    n_in: int
    n_voxel: int
    n_kept: int
    filter_stats: Optional[tuple] = None


def prepare_cloud(points: np.ndarray, voxel_down_size: float = 0.01, filter: Optional[str] = "statistical_outlier_removal",
                  nb_neighbors: int = 10, std_ratio: float = 4.0, nb_points: int = 0, radius: float = 0.0,
                  density_adaptation: bool = True) -> PreparedCloud:
    """``smartqsm.py:882-913`` in one call; the contraction receives ``points * scale`` (``:922``)."""
    p = np.asarray(points, np.float64)
    shift = global_shift(p)
    p = p + shift[None, :]                                                           # :888
    v, _ = voxel_down_sample(p, voxel_down_size)                                     # :889
    stats = None
    if filter is None:
        kept = v
    elif filter == "statistical_outlier_removal":
        keep, _, stats = remove_statistical_outlier(v, nb_neighbors, std_ratio)
        kept = v[keep]
    elif filter == "radius_outlier_removal":
        kept = v[remove_radius_outlier(v, nb_points, radius)[0]]
    else:
        raise NotImplementedError(f"Filter {filter} is not implemented.")            # :903-904
    scale = density_scale(kept, voxel_down_size) if density_adaptation else 1.0      # :908-913
    return PreparedCloud(kept, shift, scale, len(points), len(v), len(kept), stats)
