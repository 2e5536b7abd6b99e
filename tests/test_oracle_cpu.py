"""CPU tests (``-m "not gpu"``): the oracle against the reference's golden vectors, against the live
reference where the checkout exists, and against independent brute-force / dense restatements."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import contraction as OC
from oracle import network as ON
from oracle import refshim, tiling as OT, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _state_for(name):
    if name == "leafoff":
        if weights.blob_path("leafoff") is None:
            pytest.skip("reference checkpoint blob not present (baseline/_ref)")
        return weights.load_state("leafoff", allow_random=False)
    return ON.random_state(seed=7)


@pytest.mark.parametrize("name", ["random", "leafoff", "faces"])
def test_oracle_reproduces_reference_golden(name):
    """Golden vectors were produced by the reference's own driver/tiling/network code
    (tests/golden/make_golden.py); the restatement must reproduce them bit for bit."""
    g = np.load(os.path.join(GOLDEN, f"contraction_{name}.npz"))
    o = OC.SpconvBasedContractionOracle(state=_state_for(name), batch_size=int(g["batch_size"]), max_iteration=4,
                                        monitored_by_chamfer_distance=True)
    o.input(points=g["points"])
    off = 0
    for i, f in enumerate(o.get_pipeline()):
        f()
        n = int(g["sizes"][i])
        assert len(o._contracted_points) == n
        assert np.array_equal(o._contracted_points, g["P"][off:off + n])
        assert np.array_equal(o._displacements, g["D"][off:off + n])
        assert o._chamfer_distance == g["chamfer"][i]
        off += n


def test_tiling_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "tiling_faces.npz"))
    s = OT.tile_cubes(g["points"])
    assert np.array_equal([len(x.point_ids) for x in s], g["n_sample"])
    assert np.array_equal([int(x.mask.sum()) for x in s], g["n_interior"])
    assert np.array_equal(np.stack([x.slo for x in s]), g["slo"])
    assert np.array_equal(np.stack([x.shi for x in s]), g["shi"])
    # face points are members (and interior) of two cubes (F11)
    assert sum(g["n_interior"]) > len(g["points"])


def test_forward_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "forward_random.npz"))
    d, r, _ = ON.forward(g["feats"], g["indices"], g["spatial_shape"], ON.Weights(ON.random_state(seed=7)))
    assert np.array_equal(d.numpy(), g["direction"])
    assert np.array_equal(r.numpy(), g["distance"])


@pytest.mark.skipif(not refshim.available(), reason="needs the reference checkout")
def test_live_reference_driver_equals_oracle():
    """The reference's SpconvBasedContraction (unmodified, via the shim) against the restatement."""
    from smartqsm_b200.synthetic import make_tree
    st = ON.random_state(seed=3)
    pts = make_tree(5000, 41)
    ref = refshim.reference_contraction(st, batch_size=3, max_iteration=3, monitored_by_chamfer_distance=True)
    ref.input(points=pts)
    o = OC.SpconvBasedContractionOracle(state=st, batch_size=3, max_iteration=3, monitored_by_chamfer_distance=True)
    o.input(points=pts)
    for fr, fo in zip(ref.get_pipeline(), o.get_pipeline()):
        fr(); fo()
        assert np.array_equal(ref.output()["contracted_points"], o.output()["contracted_points"])
        assert np.array_equal(ref.output()["displacements"], o.output()["displacements"])


@pytest.mark.skipif(not refshim.available(), reason="needs the reference checkout")
def test_safe_to_downsample_matches_reference():
    refshim.install()
    from external.SmartTreeXX.models.sparse_module import UBlock
    import spconv.pytorch as sp
    conv = sp.SparseSequential(torch.nn.Identity(), sp.SparseConv3d(8, 16, kernel_size=2, stride=2, bias=False))
    rng = np.random.default_rng(0)
    for shape in ([5, 5, 5], [3, 9, 4], [2, 7, 7], [479, 3, 200], [1, 1, 1], [4, 4, 4]):
        for _ in range(5):
            n = int(rng.integers(1, 20))
            idx = np.concatenate([np.zeros((n, 1), np.int32),
                                  rng.integers(0, np.array(shape) + 1, (n, 3)).astype(np.int32)], axis=1)
            t = sp.SparseConvTensor(torch.zeros(n, 8), torch.from_numpy(idx), torch.tensor(shape), n)
            want = all(d > 2 for d in shape) and bool(UBlock.safe_to_downsample(conv, t))
            assert ON.safe_to_downsample(idx, shape) == want


def _dense_check(c_in, c_out, seed):
    """Sparse restatement == dense torch convolution on a small crop with inactive sites zeroed."""
    rng = np.random.default_rng(seed)
    S = 12
    occ = rng.random((S, S, S)) < 0.25
    zyx = np.argwhere(occ).astype(np.int32)
    idx = np.concatenate([np.zeros((len(zyx), 1), np.int32), zyx], axis=1)
    x = torch.from_numpy(rng.normal(size=(len(zyx), c_in))).double()
    dense = torch.zeros(1, c_in, S, S, S, dtype=torch.float64)
    dense[0, :, zyx[:, 0], zyx[:, 1], zyx[:, 2]] = x.T
    return S, zyx, idx, x, dense


def test_subm_conv_equals_dense_conv3d():
    S, zyx, idx, x, dense = _dense_check(5, 7, 0)
    w = torch.randn(7, 3, 3, 3, 5, dtype=torch.float64)
    nbr = ON.subm_rulebook(idx, [S, S, S])
    y = ON.subm_conv(x, w, nbr)
    yd = F.conv3d(dense, w.permute(0, 4, 1, 2, 3), padding=1)[0][:, zyx[:, 0], zyx[:, 1], zyx[:, 2]].T
    assert torch.allclose(y, yd, atol=1e-12)


def test_down_and_inverse_conv_equal_dense():
    S, zyx, idx, x, dense = _dense_check(4, 6, 1)
    w = torch.randn(6, 2, 2, 2, 4, dtype=torch.float64)
    rb = ON.down_rulebook(idx, [S, S, S])
    y = ON.down_conv(x, w, rb)
    yd = F.conv3d(dense, w.permute(0, 4, 1, 2, 3), stride=2)[0]
    c = rb.coarse_indices
    assert torch.allclose(y, yd[:, c[:, 1], c[:, 2], c[:, 3]].T, atol=1e-12)
    # coarse set = every 2x2x2 cell that holds an active voxel, in first-appearance order
    want = {tuple(int(a) // 2 for a in v) for v in zyx.tolist()}
    assert {tuple(v) for v in c[:, 1:].tolist()} == want
    first = [tuple(int(a) // 2 for a in v) for v in zyx.tolist()]
    seen, order = set(), []
    for f in first:
        if f not in seen:
            seen.add(f); order.append(f)
    assert [tuple(v) for v in c[:, 1:].tolist()] == order
    # inverse conv == transposed dense conv restricted to the fine active set
    wi = torch.randn(4, 2, 2, 2, 6, dtype=torch.float64)
    z = ON.inverse_conv(y, wi, rb)
    coarse_dense = torch.zeros(1, 6, S // 2, S // 2, S // 2, dtype=torch.float64)
    coarse_dense[0, :, c[:, 1], c[:, 2], c[:, 3]] = y.T
    zd = F.conv_transpose3d(coarse_dense, wi.permute(4, 0, 1, 2, 3), stride=2)[0][:, zyx[:, 0], zyx[:, 1], zyx[:, 2]].T
    assert torch.allclose(z, zd, atol=1e-12)


def test_far_face_rows_keep_centre_only():
    """App. D.3 clean bounds rule: spatial_shape = max index, rows on the far faces only see themselves."""
    idx = np.array([[0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 1], [0, 1, 1, 1]], np.int32)
    shape = idx.max(axis=0)[1:]
    nbr = ON.subm_rulebook(idx, shape)
    assert (nbr[13] == np.arange(4)).all()
    far = np.any(idx[:, 1:] >= shape[None, :], axis=1)
    others = np.delete(nbr, 13, axis=0)
    assert (others[:, far] == -1).all()
    assert (others[:, ~far] >= 0).any()


def test_first_come_voxelisation():
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 0.2, (3000, 3)).astype(np.float32)
    s = OT.tile_cubes(pts)[0]
    v = OT.voxelise_sample(pts, s)
    # brute force: iterate in order, first point creates the voxel
    seen, order = {}, []
    for i in s.point_ids:
        c = tuple(np.floor((pts[i] - s.slo) / np.float32(0.01)).astype(int)[::-1])
        if c not in seen:
            seen[c] = i; order.append(c)
    assert [tuple(c) for c in v.coords_zyx.tolist()] == order
    assert np.array_equal(v.winner_global, [seen[c] for c in order])


def test_chamfer_and_sampling_and_edges_bruteforce():
    rng = np.random.default_rng(9)
    P = rng.uniform(0, 0.3, (400, 3)).astype(np.float32)
    Q = (P + rng.normal(0, 0.01, P.shape)).astype(np.float32)[:350]
    cd = OC.chamfer(P, Q, 0.01)
    mn = np.minimum(P.astype(np.float64).min(0), Q.astype(np.float64).min(0))

    def vox(X):
        k = np.floor((X.astype(np.float64) - mn) / 0.01).astype(np.int64)
        d = {}
        for i, key in enumerate(map(tuple, k)):
            d.setdefault(key, []).append(i)
        return np.array([X[v].astype(np.float64).mean(0) for v in d.values()])
    Pv, Qv = vox(P), vox(Q)
    d2 = ((Pv[:, None, :] - Qv[None, :, :]) ** 2).sum(-1)
    want = d2.min(1).mean() + d2.min(0).mean()
    assert abs(cd - want) <= 1e-12
    ids, sk, rad = OC.skeletal_points_and_radii(P, Q[:1].repeat(len(P), 0), 0.01)
    k = np.floor((P.astype(np.float64) - P.astype(np.float64).min(0)) / 0.01).astype(np.int64)
    first = {}
    for i, key in enumerate(map(tuple, k)):
        first.setdefault(key, i)
    assert np.array_equal(ids, sorted(first.values()))
    e = OC.radius_edges(sk, 0.055)
    D = np.sqrt(((sk[:, None, :].astype(np.float64) - sk[None, :, :].astype(np.float64)) ** 2).sum(-1))
    i, j = np.nonzero(np.triu(D <= 0.055, 1))
    assert np.array_equal(e, np.stack([i, j], 1))


def test_weights_table_matches_checkpoint_shapes():
    shapes = ON.layer_shapes()
    assert len(shapes) == 161 - 25          # 161 tensors in the checkpoint, 25 of them num_batches_tracked
    assert sum(int(np.prod(s)) for s in shapes.values()) == 451661 - 25
    if weights.blob_path("leafoff"):
        st = weights.load_state("leafoff")
        for k, s in shapes.items():
            assert tuple(st[k].shape) == tuple(s), k
