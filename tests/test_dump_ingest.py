"""Ingestion of reference dumps (``scripts/reference_dump.py``): every ``tests/golden/spconv_dump_*.npz`` is compared with
the restatement (CPU) and with the B200 engine (``-m gpu``).  The committed ``spconv_dump_shim_tooling.npz`` was produced
by the same script over the oracle's stand-in modules (its ``kind`` says so): it proves the tooling, not the third-party
semantics.  A dump from a genuine spconv + Open3D environment dropped next to it is picked up automatically."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMPS = sorted(glob.glob(os.path.join(GOLD, "spconv_dump_*.npz")))


def _have_ckpt(z) -> bool:
    from oracle import weights
    return any(k.startswith("w:") for k in z.files) or weights.blob_path(str(z["ckpt_name"])) is not None


@pytest.mark.parametrize("path", DUMPS, ids=[os.path.basename(p) for p in DUMPS])
def test_restatement_reproduces_dump(path):
    from oracle import compare_dump
    z = np.load(path)
    if not _have_ckpt(z):
        pytest.skip("checkpoint blob not present")
    rep = compare_dump.compare(path, verbose=False)
    bad = [i for i in rep["items"] if not i[1] and "voxel order" not in i[0]]      # order is a CPU-wheel property only
    assert not bad, f"{rep['kind']}: {bad[:5]}"
    assert rep["layer_rel_err"]["DOWN_KERNEL_FLIP=False"] <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("path", DUMPS, ids=[os.path.basename(p) for p in DUMPS])
def test_engine_reproduces_dump(path):
    """The B200 engine on the dump's input: voxel sets of the first iteration, the state after every thunk (rows as a
    multiset: the GPU wheel orders voxels differently) and the Chamfer values."""
    from oracle import compare_dump
    from smartqsm_b200._lib import Engine
    z = np.load(path)
    if not _have_ckpt(z):
        pytest.skip("checkpoint blob not present")
    eng = Engine(batch_size=int(z["batch_size"]), monitored_by_chamfer_distance=True)
    eng.load_state_dict(compare_dump.dump_state(z))
    eng.debug_capture(True)
    eng.set_points(z["points"])
    prev = np.inf
    for t in range(int(z["thunks"])):
        st = eng.contract_step()
        if t == 0:
            idx, _, _ = eng.debug_voxels()
            bs = int(z["batch_size"])
            got = np.concatenate([idx[:, :1] // bs, idx[:, :1] % bs, idx[:, 1:]], axis=1)       # (batch, sample in batch, z, y, x)
            want = np.concatenate([np.concatenate([np.full((len(z[f"b{b}:indices"]), 1), b), z[f"b{b}:indices"]], axis=1)
                                   for b in range(int(z["n_batches"]))])
            assert sorted(map(tuple, got)) == sorted(map(tuple, want)), "voxel sets of the first iteration differ"
            eng.debug_capture(False)
        P, D = eng.result()
        dP = z[f"P{t}"]
        assert P.shape == dP.shape
        a, b = P[np.lexsort(P.T[::-1])], dP[np.lexsort(dP.T[::-1])]
        assert np.max(np.abs(a - b)) <= 1e-4
        cd = float(z[f"chamfer{t}"])
        if st["accepted"]:
            assert abs(st["chamfer"] - cd) <= 1e-4 * cd
            prev = cd
        else:
            assert st["chamfer"] > prev * (1 - 1e-9)
