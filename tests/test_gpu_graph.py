"""Row f1 on the device (``csrc/graph.cu`` through the C ABI) against golden vectors of the reference's OWN
``_construct_graph`` / ``_extract_skeleton`` and against the restatement on further seeded inputs.  Everything here is
integer / float32-bit-pattern work: bit-exact.  ``-m gpu``."""
import os

import numpy as np
import pytest

from oracle import graph as OG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_graph(sp, r=0.055, weight="l1"):
    from scipy.spatial import Delaunay
    from smartqsm_b200._lib import Engine
    eng = Engine()
    edges = eng.graph_construct(Delaunay(sp).simplices, r=r, points=sp)
    sk = eng.graph_extract_skeleton(len(sp), weight)
    return eng, edges, sk


@pytest.mark.parametrize("name", ["tree_leafoff", "tree_leafon", "lattice"])
def test_graph_stage_equals_reference_golden(name):
    from tests.test_graph_cpu import digraph_node_order
    g = np.load(os.path.join(GOLD, f"graph_{name}.npz"))
    sp = g["skeletal_points"]
    eng, edges, sk = gpu_graph(sp)
    assert np.array_equal(edges, g["edges"].astype(np.int64)), "merged edge list (:158) differs"
    assert sk["root"] == int(g["node_order"][0])
    assert np.array_equal(sk["pred"], g["pred"].astype(np.int64)), "predecessors[0] differ"
    assert np.array_equal(digraph_node_order(sk["pred"], sk["order"]), g["node_order"].astype(np.int64))
    reached = sk["pred"] >= 0
    assert np.array_equal(sk["weight"][reached], g["weight"][reached]), "float32 skeleton edge weights differ"


@pytest.mark.parametrize("seed,n,leaf_on", [(51, 25000, False), (52, 40000, True), (53, 15000, False)])
def test_graph_stage_equals_restatement_on_seeded_trees(seed, n, leaf_on):
    """>= 3 further seeded trees: the voxel-sampled surface of a synthetic tree as skeletal points (no contraction needed:
    the stage only sees float32 points on a 1 cm lattice-like sampling, ties included)."""
    from oracle import contraction as OC
    from smartqsm_b200.synthetic import make_tree
    pts = make_tree(n, seed, leaf_on).astype(np.float32)
    _, sp, _ = OC.skeletal_points_and_radii(pts, np.zeros_like(pts), 0.02)
    eng, edges, sk = gpu_graph(sp)
    cg = OG.construct_graph(sp)
    assert np.array_equal(edges, cg["edges"])
    info = eng.graph_info()
    assert info["delaunay_root"] == int(cg["delaunay_root"])
    assert info["n_spt_edges"] == len(cg["spt_edges"]) and info["n_mst_edges"] == len(cg["mst_edges"])
    so = OG.extract_skeleton(sp, cg["edges"])
    assert sk["root"] == int(so["root"])
    assert np.array_equal(sk["pred"], so["pred"]) and np.array_equal(sk["order"], so["order"])
    assert np.array_equal(sk["weight"][so["order"]], so["weight"])


def test_l2_weights_and_digraph_order():
    from smartqsm_b200.skeleton_graph import skeleton_to_digraph
    rng = np.random.default_rng(3)
    sp = (np.round(rng.uniform(0, 0.4, (3000, 3)) / 0.01) * 0.01).astype(np.float32)
    sp = np.unique(sp, axis=0)
    sp = sp[rng.permutation(len(sp))]
    eng, edges, sk = gpu_graph(sp, weight="l2")
    cg = OG.construct_graph(sp)
    assert np.array_equal(edges, cg["edges"])
    so = OG.extract_skeleton(sp, cg["edges"], "l2")
    assert np.array_equal(sk["pred"], so["pred"]) and np.array_equal(sk["order"], so["order"])
    g, go = skeleton_to_digraph(sk), OG.skeleton_as_digraph(sp, so)
    assert list(g.nodes) == list(go.nodes) and list(g.edges) == list(go.edges)
    assert all(g[u][v]["weight"] == go[u][v]["weight"] for u, v in g.edges)


def test_graph_from_engine_state(leafoff_state):
    """The drop-in flow: contraction -> skeletal sample -> graph -> skeleton, all on one engine, no host round trip of the
    skeletal points; equals the restatement on the same skeletal points."""
    from smartqsm_b200._lib import Engine
    from smartqsm_b200.skeleton_graph import construct_graph, extract_skeleton
    from smartqsm_b200.synthetic import make_tree
    eng = Engine(monitored_by_chamfer_distance=True)
    eng.load_state_dict(leafoff_state)
    eng.set_points(make_tree(20000, 61))
    for _ in range(3):
        eng.contract_step()
    ids, sp, rad = eng.skeletal_sample(0.01)
    edges = construct_graph(eng, None, {"r": 0.055})
    sk = extract_skeleton(eng, len(sp))
    cg = OG.construct_graph(sp)
    assert np.array_equal(edges, cg["edges"])
    so = OG.extract_skeleton(sp, cg["edges"])
    assert np.array_equal(sk["pred"], so["pred"]) and np.array_equal(sk["order"], so["order"])
