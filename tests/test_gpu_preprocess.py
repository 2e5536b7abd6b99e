"""GPU parity tests of the cloud preparation through the C ABI against ``oracle/preprocess.py``:
bit-exact shift, voxel means, kept set and float32 hand-over; float64 statistics within 1e-12 relative (the device
reduces global sums with a fixed tree, NumPy pairwise)."""
import numpy as np
import pytest

from tests.helpers import make_raw_scan

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from smartqsm_b200._lib import Engine
    return Engine(device=0)


def _check(engine, raw, **cfg):
    from oracle import preprocess as OP
    filt = cfg.get("filter", "statistical_outlier_removal")
    dens = cfg.get("density_adaptation", True)
    nb, ratio = cfg.get("nb_neighbors", 10), cfg.get("std_ratio", 4.0)
    nbp, rad = cfg.get("nb_points", 0), cfg.get("radius", 0.0)
    code = {None: 0, "statistical_outlier_removal": 1, "radius_outlier_removal": 2}[filt]
    ref = OP.prepare_cloud(raw, 0.01, filt, nb, ratio, nbp, rad, density_adaptation=dens)
    st = engine.prepare_cloud(raw, voxel_down_size=0.01, filter=code, nb_neighbors=nb, std_ratio=ratio, nb_points=nbp, radius=rad,
                              density_adaptation=dens, keep_debug=True)
    assert np.array_equal(st["global_shift"], ref.global_shift)
    assert (st["n_in"], st["n_voxel"], st["n_kept"]) == (ref.n_in, ref.n_voxel, ref.n_kept)
    vox, avg, keep = engine.debug_prepared(with_filter=bool(filt))
    v_ref, _ = OP.voxel_down_sample(raw + ref.global_shift[None, :], 0.01)
    assert np.array_equal(vox, v_ref), "voxel means must be bit-identical (input-order sums)"
    if filt == "radius_outlier_removal":
        keep_ref, cnt_ref = OP.remove_radius_outlier(v_ref, nbp, rad)
        assert np.array_equal(avg.astype(np.int64), cnt_ref), "neighbour counts (open ball, float64) must be identical"
        assert np.array_equal(np.nonzero(keep)[0], keep_ref)
    elif filt:
        keep_ref, avg_ref, (mean, std, thr) = OP.remove_statistical_outlier(v_ref, nb, ratio)
        np.testing.assert_allclose(avg, avg_ref, rtol=1e-12, atol=0)
        np.testing.assert_allclose([st["cloud_mean"], st["std_dev"], st["threshold"]], [mean, std, thr], rtol=1e-12)
        assert np.array_equal(np.nonzero(keep)[0], keep_ref)
    got = engine.prepared_points()
    assert np.array_equal(got, ref.points)
    np.testing.assert_allclose(st["scale"], ref.scale, rtol=1e-12)
    return st, ref


def test_prepare_cloud_matches_oracle(engine):
    st, ref = _check(engine, make_raw_scan(60000, 7))
    assert st["n_kept"] < st["n_voxel"], "the far outliers must be filtered"
    assert st["n_far_filter"] > 0, "the sprinkled outliers must exercise the far search"
    assert st["gpu_launches"] > 0


def test_prepare_cloud_leaf_on_and_other_parameters(engine):
    _check(engine, make_raw_scan(40000, 8, leaf_on=True), nb_neighbors=6, std_ratio=1.5)


def test_prepare_cloud_without_filter_or_scale(engine):
    st, _ = _check(engine, make_raw_scan(20000, 9), filter=None, density_adaptation=False)
    assert st["scale"] == 1.0 and st["n_kept"] == st["n_voxel"]
    _check(engine, make_raw_scan(20000, 9), filter=None, density_adaptation=True)


def test_radius_outlier_removal(engine):
    """``filter: radius_outlier_removal`` (smartqsm.py:899-902): counts inside one ring of bricks (r < 8 cm) and beyond it
    (r = 10 cm probes the hash for the second ring)."""
    raw = make_raw_scan(30000, 14)
    st, _ = _check(engine, raw, filter="radius_outlier_removal", nb_points=6, radius=0.03)
    assert st["n_kept"] < st["n_voxel"]
    _check(engine, make_raw_scan(8000, 15, leaf_on=True), filter="radius_outlier_removal", nb_points=40, radius=0.1,
           density_adaptation=False)


def test_isolated_points_use_the_ring_search(engine):
    # two tight clusters 3 m apart plus single points in between: k-NN must look far beyond the 27 bricks
    rng = np.random.default_rng(5)
    a = rng.normal(0, 0.02, (300, 3))
    b = rng.normal(0, 0.02, (300, 3)) + [3.0, 0.5, 1.0]
    lone = np.array([[1.5, 0.2, 0.4], [0.7, -0.9, 2.0], [2.2, 1.4, 0.1]])
    _check(engine, np.concatenate([a, b, lone]), nb_neighbors=10, std_ratio=4.0)


def test_far_search_without_the_occupancy_bitmap(engine, monkeypatch):
    monkeypatch.setenv("SQSM_PREP_NO_BITMAP", "1")          # lattices too large for the bitmap probe the hash instead
    st, _ = _check(engine, make_raw_scan(20000, 10))
    assert st["n_far_filter"] > 0


def test_hand_over_to_contraction_is_the_scaled_float32_cloud(engine):
    raw = make_raw_scan(30000, 11)
    _, ref = _check(engine, raw)
    engine.set_points_prepared()
    P, D = engine.result()
    assert np.array_equal(P, (ref.points * ref.scale).astype(np.float32))       # smartqsm.py:922 + contraction :82
    assert not D.any()


def test_prepare_input_on_the_plugin_equals_host_round_trip():
    from oracle import weights
    from smartqsm_b200.core_algorithms import SpconvBasedContraction
    state = weights.load_state("leafoff")
    raw = make_raw_scan(30000, 12)
    kw = dict(batch_size=16, max_iteration=2, device="cuda", monitored_by_chamfer_distance=True, state_dict=state)
    cfg = dict(voxel_down_size=0.01, filter="statistical_outlier_removal", kwargs_for_filter={"nb_neighbors": 10, "std_ratio": 4.0},
               using_density_adaptation_in_skeletonization=True)
    a = SpconvBasedContraction(**kw)
    prepared = a.prepare_input(raw, **cfg)
    a.input(points=prepared)                                                   # what `skeletonization.run(points=prepared)` does
    for thunk in a.get_pipeline():
        thunk()
    b = SpconvBasedContraction(**kw)
    b.input(points=prepared.points * prepared.scale)                           # what the reference entry point does
    for thunk in b.get_pipeline():
        thunk()
    oa, ob = a.output(), b.output()
    assert np.array_equal(oa["contracted_points"], ob["contracted_points"])
    assert np.array_equal(oa["displacements"], ob["displacements"])


def test_errors_and_determinism(engine):
    with pytest.raises(RuntimeError, match="removed every point"):
        engine.prepare_cloud(np.array([[1.0, 2.0, 3.0]]))
    with pytest.raises(RuntimeError, match="NaN"):
        engine.prepare_cloud(np.array([[1.0, np.nan, 3.0], [0.0, 0.0, 0.0]]))
    with pytest.raises(NotImplementedError):
        from smartqsm_b200.preprocess import prepare_cloud
        prepare_cloud(np.zeros((4, 3)), filter="bilateral", engine=engine)
    raw = make_raw_scan(50000, 13, leaf_on=True)
    engine.prepare_cloud(raw)
    p1 = engine.prepared_points()
    s1 = engine.prepare_cloud(raw)
    assert np.array_equal(p1, engine.prepared_points())
    s2 = engine.prepare_cloud(raw)
    assert s1["scale"] == s2["scale"] and s1["threshold"] == s2["threshold"]
