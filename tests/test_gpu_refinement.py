"""Row f3: variable-radius ball lists (``sqsm_ball_lists``) against SciPy's KDTree - the reference's own dependency
(core/refinement.py:420,448-455).  ``-m gpu``; integer outputs, bit-exact."""
import numpy as np
import pytest

from oracle import refinement as OR

pytestmark = pytest.mark.gpu


def _engine():
    from smartqsm_b200._lib import Engine
    return Engine()


def _check(pts, radii):
    off_g, idx_g = _engine().ball_lists(pts, radii)
    off_o, idx_o = OR.ball_lists(pts, radii)
    assert np.array_equal(off_g, off_o)
    assert np.array_equal(idx_g.astype(np.int64), idx_o)


def test_ball_lists_on_a_skeleton_like_cloud():
    """Contracted tree points as nodes, radii over three orders of magnitude (twigs 2 mm ... trunk 0.4 m)."""
    from smartqsm_b200.synthetic import make_tree
    rng = np.random.default_rng(3)
    pts = make_tree(40000, 9).astype(np.float32)
    radii = np.exp(rng.uniform(np.log(0.002), np.log(0.4), len(pts)))
    radii[rng.random(len(pts)) < 0.9] *= 0.05                       # most nodes are thin
    _check(pts, radii)


def test_ball_lists_boundaries_and_degenerate_radii():
    """Closed ball in float64 (d2 <= r * r) with pairs at EXACTLY the radius, zero radii (the node itself only), negative
    radii (cKDTree squares them), NaN radii (empty lists), one radius larger than the whole cloud."""
    rng = np.random.default_rng(4)
    pts = rng.uniform(0, 0.5, (5000, 3)).astype(np.float32)
    r = 0.0625
    pts[:100] = np.round(pts[:100] * 64) / 64
    pts[100:200] = pts[:100] + np.array([r, 0, 0], np.float32)      # exact in float32
    radii = np.full(len(pts), r)
    radii[200:300] = 0.0
    radii[300:310] = -1.0
    radii[310:320] = np.nan
    radii[320] = 10.0
    _check(pts, radii)


def test_inlier_nodes_of_a_path_equal_the_reference_expression():
    """BallOccupancy.inlier_nodes == np.unique(np.concatenate(kdtree.query_ball_point(points[path], r=radii[path])))."""
    from smartqsm_b200.refinement_queries import BallOccupancy, occupancy_radii
    from smartqsm_b200.synthetic import make_tree
    rng = np.random.default_rng(5)
    pts = make_tree(20000, 10).astype(np.float32)
    ref_r = rng.uniform(0.001, 0.05, len(pts))
    for fac in (1.5, 0.01):                                          # factor (>= 1) and additive buffer (< 1), :450
        radii = occupancy_radii(ref_r, fac)
        occ = BallOccupancy(_engine(), pts, radii)
        for _ in range(5):
            path = rng.choice(len(pts), size=int(rng.integers(1, 200)), replace=False).tolist()
            assert np.array_equal(occ.inlier_nodes(path), OR.inlier_nodes(pts, radii, path))
        assert len(occ.inlier_nodes([])) == 0
