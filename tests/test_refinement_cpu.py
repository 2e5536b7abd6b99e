"""Row f3, host side (no GPU): the ``BallOccupancy`` mirror of refinement's occupancy loop (core/refinement.py:405-470)
served from an engine stand-in whose ``ball_lists`` is the SciPy oracle - the loop logic, not the device kernel, is under
test here (the kernel is tested against the same oracle in ``tests/test_gpu_refinement.py``)."""
import numpy as np

from oracle import refinement as OR
from smartqsm_b200.refinement_queries import BallOccupancy, occupancy_radii


class _OracleEngine:
    def ball_lists(self, pts, radii):
        off, idx = OR.ball_lists(np.asarray(pts, np.float32), radii)
        return off, idx.astype(np.int32)


def _skeleton(n_paths=12, seed=0):
    """A few polylines with 1 cm node spacing (a trunk and branches), radii tapering along each path."""
    rng = np.random.default_rng(seed)
    pts, radii, paths = [], [], []
    for p in range(n_paths):
        start = rng.uniform(0, 1.0, 3)
        direction = rng.normal(0, 1, 3)
        direction /= np.linalg.norm(direction)
        n = int(rng.integers(20, 120))
        base = len(pts)
        r0 = rng.uniform(0.005, 0.08)
        for i in range(n):
            pts.append(start + direction * 0.01 * i + rng.normal(0, 0.001, 3))
            radii.append(r0 * (1.0 - 0.7 * i / n))
        paths.append(list(range(base, base + n)))
    return np.asarray(pts, np.float32), np.asarray(radii), paths


def test_occupancy_radii_expression():
    r = np.array([0.01, 0.05, 0.0])
    assert np.array_equal(occupancy_radii(r, 1.5), r * 1.5)                  # factor (>= 1), core/refinement.py:450
    assert np.array_equal(occupancy_radii(r, 0.02), r + 0.02)                # additive buffer (< 1)


def test_occupancy_loop_equals_the_reference_expressions():
    """The sequential loop of ``_identify_reasonable_branches`` (:421-458, reduced to its occupancy logic: trim the nodes
    already occupied, query the balls of what is left, occupy the inliers) run twice - with the reference's KDTree
    expressions and with ``BallOccupancy`` - must leave the same ``after_check`` and the same trimmed paths."""
    from scipy.spatial import KDTree
    pts, ref_r, paths = _skeleton()
    for fac in (1.5, 0.01):
        radii = occupancy_radii(ref_r, fac)
        occ = BallOccupancy(_OracleEngine(), pts, radii)
        kdtree = KDTree(pts)
        after_a = np.zeros(len(pts), bool)
        after_b = np.zeros(len(pts), bool)
        trimmed_a, trimmed_b = [], []
        for path in paths:
            pa = [n for n in path if not after_a[n]]                         # :426-439 (occupied nodes leave the path)
            pb = [n for n in path if not after_b[n]]
            trimmed_a.append(pa)
            trimmed_b.append(pb)
            if len(pa) > 1:
                lists = kdtree.query_ball_point(pts[pa], r=radii[pa])       # :448-451
                after_a[np.unique(np.concatenate(lists))] = True            # :453-456
            if len(pb) > 1:
                after_b[occ.inlier_nodes(pb)] = True
        assert trimmed_a == trimmed_b
        assert np.array_equal(after_a, after_b)
        for n in (0, len(pts) // 2, len(pts) - 1):
            assert np.array_equal(occ.query_ball_point(n), np.sort(kdtree.query_ball_point(pts[n], r=radii[n])))
