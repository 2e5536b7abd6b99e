"""CPU tests of the row-f2 oracle (``oracle/preprocess.py``): the restated Open3D / SciPy operators of
``entrypoints/smartqsm.py:882-913`` against brute-force definitions and their invariants."""
import numpy as np
import pytest

from oracle import preprocess as OP
from tests.helpers import make_raw_scan


def test_global_shift_centres_xy_and_grounds_z():
    raw = make_raw_scan(3000, 1)
    s = OP.global_shift(raw)
    p = raw + s
    assert p[:, 2].min() == 0.0
    assert abs(p[:, 0].min() + p[:, 0].max()) < 1e-9 and abs(p[:, 1].min() + p[:, 1].max()) < 1e-9


def test_voxel_down_sample_means_and_order():
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 0.3, (5000, 3))
    v, idx = OP.voxel_down_sample(p, 0.01)
    vmin = p.min(axis=0) - 0.005
    # every mean lies in its voxel, voxels are unique, order is (brick z,y,x ; local z,y,x)
    assert np.array_equal(np.floor((v - vmin) / 0.01).astype(np.int64), idx)
    assert len(np.unique(idx, axis=0)) == len(idx)
    key = np.stack([idx[:, 2] >> 3, idx[:, 1] >> 3, idx[:, 0] >> 3, idx[:, 2] & 7, idx[:, 1] & 7, idx[:, 0] & 7], axis=1)
    assert np.array_equal(np.lexsort(key.T[::-1]), np.arange(len(key)))
    # brute force mean of one voxel, summed in input order
    pid = np.floor((p - vmin) / 0.01).astype(np.int64)
    for r in (0, len(idx) // 2, len(idx) - 1):
        members = np.nonzero((pid == idx[r]).all(axis=1))[0]
        s = np.zeros(3)
        for m in members:
            s = s + p[m]
        assert np.array_equal(s / len(members), v[r])
    assert OP.voxel_down_sample(np.zeros((0, 3)), 0.01)[0].shape == (0, 3)


def test_statistical_outlier_matches_brute_force():
    rng = np.random.default_rng(1)
    p = np.concatenate([rng.normal(0, 0.05, (600, 3)), rng.uniform(-2, 2, (12, 3))])
    keep, avg, (mean, std, thr) = OP.remove_statistical_outlier(p, 10, 2.0)
    d = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
    d.sort(axis=1)
    avg_b = d[:, :10].sum(axis=1) / 10.0                       # the point itself (0) is one of the 10
    np.testing.assert_allclose(avg, avg_b, rtol=1e-12)
    mean_b = avg_b.sum() / len(p)
    std_b = np.sqrt(((avg_b - mean_b) ** 2).sum() / (len(p) - 1))
    np.testing.assert_allclose([mean, std, thr], [mean_b, std_b, mean_b + 2.0 * std_b], rtol=1e-12)
    assert np.array_equal(keep, np.nonzero((avg_b > 0) & (avg_b < mean_b + 2.0 * std_b))[0])
    assert 0 < len(keep) < len(p)


def test_radius_outlier_matches_brute_force():
    rng = np.random.default_rng(4)
    p = np.concatenate([rng.normal(0, 0.05, (500, 3)), rng.uniform(-2, 2, (20, 3))])
    p[1] = p[0] + np.array([0.125, 0.0, 0.0])                  # a pair at exactly r: the ball is open, it does not count
    keep, cnt = OP.remove_radius_outlier(p, 5, 0.125)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    cnt_b = (d2 < 0.125 * 0.125).sum(axis=1)                   # itself included
    assert np.array_equal(cnt, cnt_b)
    assert np.array_equal(keep, np.nonzero(cnt_b > 5)[0]) and 0 < len(keep) < len(p)


def test_density_scale_matches_brute_force():
    rng = np.random.default_rng(2)
    p = rng.uniform(0, 1, (700, 3))
    d = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
    np.fill_diagonal(d, np.inf)
    np.testing.assert_allclose(OP.density_scale(p, 0.01), 0.01 / d.min(axis=1).mean(), rtol=1e-12)


def test_prepare_cloud_pipeline_and_unknown_filter():
    raw = make_raw_scan(8000, 3)
    pc = OP.prepare_cloud(raw)
    assert pc.n_in == len(raw) and pc.n_kept <= pc.n_voxel <= pc.n_in
    assert pc.points.shape == (pc.n_kept, 3) and pc.scale > 0
    assert pc.points[:, 2].min() >= 0.0
    nof = OP.prepare_cloud(raw, filter=None, density_adaptation=False)
    assert nof.n_kept == nof.n_voxel and nof.scale == 1.0
    with pytest.raises(NotImplementedError):
        OP.prepare_cloud(raw, filter="bilateral")
