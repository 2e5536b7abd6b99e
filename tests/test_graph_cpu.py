"""Row f1 (graph + skeleton extraction): the CPU restatement against golden vectors that came out of the reference's
OWN ``_construct_graph`` / ``_extract_skeleton`` (tests/golden/make_golden_graph.py) - this row's parity is pinned:
SciPy and networkx, its third-party dependencies, are the real libraries."""
import os

import numpy as np
import pytest

from oracle import graph as OG

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tree_leafoff", "tree_leafon", "lattice"]


def digraph_node_order(pred, order):
    """Node insertion order of the reference's skeleton DiGraph (:205-213): edges (predecessors[0], node) are added in the
    order networkx first SAW the nodes; a predecessor found later than its child enters the graph with the child's edge."""
    seen, out = set(), []
    for n in order:
        for x in (int(pred[n]), int(n)):
            if x not in seen:
                seen.add(x)
                out.append(x)
    return np.array(out, np.int64)


def load(name):
    g = np.load(os.path.join(GOLD, f"graph_{name}.npz"))
    return {k: g[k] for k in g.files}


@pytest.mark.parametrize("name", CASES)
def test_restatement_equals_reference_golden(name):
    g = load(name)
    sp = g["skeletal_points"]
    cg = OG.construct_graph(sp, r=0.055)
    assert np.array_equal(cg["edges"], g["edges"].astype(np.int64)), "merged edge list (:158) differs"
    sk = OG.extract_skeleton(sp, cg["edges"])
    assert int(sk["root"]) == int(g["node_order"][0])
    assert np.array_equal(digraph_node_order(sk["pred"], sk["order"]), g["node_order"].astype(np.int64))
    assert np.array_equal(sk["pred"], g["pred"].astype(np.int64)), "predecessors[0] differ"
    assert np.array_equal(sk["weight"], g["weight"][sk["order"]]), "float32 skeleton weights differ"


def test_live_reference_graph_stage_equals_restatement():
    """Where the reference checkout exists: run its own classes on a fresh seeded cloud and compare."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("/root/reference not present")
    rng = np.random.default_rng(9)
    t = rng.uniform(0, 1, 1500)
    sp = np.stack([0.05 * np.cos(9 * t), 0.05 * np.sin(9 * t), 2.0 * t], 1) + rng.normal(0, 0.004, (1500, 3))
    sp = np.round(sp / 0.01).astype(np.float32) * np.float32(0.01)            # lattice-snapped: many exact ties
    sp = np.unique(sp, axis=0)[rng.permutation(len(np.unique(sp, axis=0)))]
    p = refshim.reference_graph_stage(sp)
    cg = OG.construct_graph(sp)
    assert np.array_equal(cg["edges"], p._merged_edges)
    sk = OG.extract_skeleton(sp, cg["edges"])
    g = OG.skeleton_as_digraph(sp, sk)
    assert list(g.nodes) == list(p._skeleton.nodes) and list(g.edges) == list(p._skeleton.edges)
    assert all(g[u][v]["weight"] == p._skeleton[u][v]["weight"] for u, v in g.edges)


def test_l1_weight_convention():
    """np.linalg.norm on ONE float32 3-vector = sqrt(sdot): float32 products accumulated in a C double (OpenBLAS tail
    loop), rounded to float32, float32 sqrt - the formula the device kernel uses."""
    rng = np.random.default_rng(0)
    v = rng.normal(0, 0.03, (20000, 3)).astype(np.float32)
    ref = np.array([np.linalg.norm(x) for x in v], np.float32)
    sq = v * v
    acc = (sq[:, 0].astype(np.float64) + sq[:, 1].astype(np.float64)) + sq[:, 2].astype(np.float64)
    assert np.array_equal(ref, np.sqrt(acc.astype(np.float32)))
