"""TEST INFRASTRUCTURE ONLY - CPU oracle of row f3 (refinement's ball occupancy queries).

The reference calls SciPy directly (``core/refinement.py:420,448-455``: ``KDTree(points)``,
``kdtree.query_ball_point(points[path], r=radii)``, ``np.unique(np.concatenate(...))``); SciPy is importable here, so the
oracle IS the reference's own dependency - parity of this row is pinned, not restated.
"""
import numpy as np
from scipy.spatial import KDTree


def ball_lists(points: np.ndarray, radii: np.ndarray):
    """(offsets, indices) of ``KDTree(points).query_ball_point(points, r=radii)`` with every list sorted ascending."""
    pts = np.asarray(points)
    lists = KDTree(pts).query_ball_point(pts, r=np.asarray(radii, dtype=np.float64))
    lists = [np.sort(np.asarray(l, dtype=np.int64)) for l in lists]
    off = np.zeros(len(lists) + 1, np.int64)
    off[1:] = np.cumsum([len(l) for l in lists])
    return off, (np.concatenate(lists) if len(lists) and off[-1] else np.empty(0, np.int64))


def inlier_nodes(points: np.ndarray, radii: np.ndarray, path) -> np.ndarray:
    """core/refinement.py:448-455 verbatim."""
    pts = np.asarray(points)
    kdtree = KDTree(pts)
    neighbor_ids_per_node = kdtree.query_ball_point(pts[path], r=np.asarray(radii, dtype=np.float64)[path])
    return np.unique(np.concatenate(neighbor_ids_per_node)).astype(np.int64)
