"""Row f3: the spatial queries of SmartQSM's refinement stage on the device.

``core/refinement.py:_identify_reasonable_branches`` (:405-470) walks the paths of the skeleton in order and, for every
path, occupies the nodes inside a series of balls::

    neighbor_ids_per_node = kdtree.query_ball_point(points[path], r=reference_radii[path] * factor)   # :448-451
    inlier_nodes = np.unique(np.concatenate(neighbor_ids_per_node))                                   # :453-455
    after_check[inlier_nodes] = True

The loop is sequential (an occupied node is removed from every later path), but the radius of EVERY node is known before it
starts (``_calculate_reference_radii``, :336-373).  ``BallOccupancy`` therefore asks the device once for the ball of every
node with its own radius (``sqsm_ball_lists``: a radius graph with per-node radii, cKDTree semantics, CSR) and serves the
loop from the precomputed lists; ``install_into`` shows the two lines a maintainer changes.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def occupancy_radii(reference_radii: np.ndarray, occupancy_factor_or_buffer: float) -> np.ndarray:
    """The radius expression of core/refinement.py:450 for all nodes at once."""
    r = np.asarray(reference_radii, dtype=np.float64)
    return r * occupancy_factor_or_buffer if occupancy_factor_or_buffer >= 1.0 else r + occupancy_factor_or_buffer


class BallOccupancy:
    """Precomputed ``query_ball_point`` lists of all nodes (device), united per path on the host."""

    def __init__(self, engine, skeletal_points: np.ndarray, radii: np.ndarray) -> None:
        self.offsets, self.indices = engine.ball_lists(skeletal_points, radii)

    def query_ball_point(self, node: int) -> np.ndarray:
        """``kdtree.query_ball_point(points[node], r=radii[node])`` (ascending indices)."""
        return self.indices[self.offsets[node]:self.offsets[node + 1]]

    def inlier_nodes(self, path: Sequence[int]) -> np.ndarray:
        """``np.unique(np.concatenate(kdtree.query_ball_point(points[path], r=radii[path])))`` (:448-455)."""
        if len(path) == 0:
            return np.empty(0, np.int64)
        parts = [self.indices[self.offsets[n]:self.offsets[n + 1]] for n in path]
        return np.unique(np.concatenate(parts)).astype(np.int64)


def install_into(refinement, engine) -> BallOccupancy:
    """What the reference would do in ``_identify_reasonable_branches`` just before the loop over the paths::

        occ = BallOccupancy(engine, self._skeletal_points, occupancy_radii(self._reference_radii, self._occupancy_factor_or_buffer))
        ...
        inlier_nodes = occ.inlier_nodes(path)          # replaces :448-455
    """
    radii = occupancy_radii(refinement._reference_radii, refinement._occupancy_factor_or_buffer)
    return BallOccupancy(engine, np.asarray(refinement._skeletal_points), radii)
