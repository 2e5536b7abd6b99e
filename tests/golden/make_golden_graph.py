"""Golden vectors of the graph / skeleton stage (SURVEY §8 row a16 = f1), from the REFERENCE's own code.

Run in the authoring container (needs /root/reference):  python tests/golden/make_golden_graph.py

``core/skeletonization_pipeline_base.py`` (``_construct_graph`` :75-188, ``_extract_skeleton`` :190-238) is imported
unmodified through ``oracle/refshim.py``; for this stage the third-party arithmetic is SciPy (Qhull, cKDTree) and
networkx, both REAL here - only ``open3d`` (touched in verbose branches) is a stand-in.  The skeletal points of the
"tree" cases also come out of the reference's own thinning pipeline (``ThinningBasedSkeletonizationPipeline``, contraction on
the stub spconv / Open3D of refshim).  Written per case: skeletal points, radii, the merged edge list (``_edges``, :158),
the skeleton's node insertion order, ``predecessors[0]`` per node and the float32 edge weights.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import refshim, weights  # noqa: E402
from smartqsm_b200.synthetic import make_tree  # noqa: E402

CFG = dict(ckpt_path="/nonexistent.pth", batch_size=16, max_iteration=10, device="cpu", voxel_size=0.01, cube_size=4.0,
           buffer_size=0.4, monitored_by_chamfer_distance=True)


def thinning_skeletal_points(n, seed, leaf_on, which):
    """points -> reference thinning pipeline (its own _contract thunks + _produce_skeletal_points_and_radii)."""
    import torch
    state = weights.load_state(which, allow_random=False)
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state.items()}
    ck = {"model_state_dict": sd, "hparams": {"u_channels": [8, 16, 32, 64], "mlp_hidden_layers": [8, 4], "conv_block_reps": 1}}
    real_load = torch.load
    torch.load = lambda *a, **k: ck
    try:
        p = refshim.reference_thinning_pipeline(core_algorithm="spconv_based_contraction", core_algorithm_kwargs=CFG,
                                                voxel_size_for_downsampling=0.01, edge_search_kwargs={"r": 0.055},
                                                edge_weight_function="l1", topology_extractor="spt")
    finally:
        torch.load = real_load
    pts = make_tree(n, seed, leaf_on)
    gen = p.run(points=pts)
    try:
        while True:
            next(gen)
    except StopIteration as stop:
        sk, skeleton, radii = stop.value
    return pts, sk, radii, skeleton, p


def lattice_cloud():
    """Worst case for tie-breaking: a 1 cm lattice shell (equal weights, equal path lengths, cospherical tetrahedra),
    a far satellite cluster and two exact duplicates."""
    g = np.stack(np.meshgrid(np.arange(14), np.arange(14), np.arange(30), indexing="ij"), -1).reshape(-1, 3)
    shell = g[(np.abs(g[:, 0] - 6.5) > 4) | (np.abs(g[:, 1] - 6.5) > 4)]
    rng = np.random.default_rng(4)
    pts = np.concatenate([shell * 0.01, rng.uniform(0.5, 0.6, (300, 3)), shell[:2] * 0.01]).astype(np.float32)
    return pts[rng.permutation(len(pts))]


def dump(name, sk, radii, p):
    skel = p._skeleton
    nodes = np.array([int(n) for n in skel.nodes], np.int64)
    pred = np.full(len(sk), -1, np.int64)
    w = np.zeros(len(sk), np.float32)
    for u, v, d in skel.edges(data=True):
        pred[int(v)] = int(u)
        w[int(v)] = d["weight"]
        assert isinstance(d["weight"], np.float32)
    edges = np.array(p._merged_edges, np.int64)
    np.savez_compressed(os.path.join(HERE, f"graph_{name}.npz"), skeletal_points=np.asarray(sk, np.float32),
                        radii=np.zeros(0, np.float32) if radii is None else np.asarray(radii, np.float32),
                        edges=edges.astype(np.int32), node_order=nodes.astype(np.int32), pred=pred.astype(np.int32), weight=w)
    print(f"graph_{name}: {len(sk)} nodes, {len(edges)} edges, skeleton {skel.number_of_edges()} edges, root {nodes[0]}")


def main():
    for name, (n, seed, leaf_on, which) in {"tree_leafoff": (12000, 41, False, "leafoff"),
                                            "tree_leafon": (9000, 42, True, "leafon")}.items():
        pts, sk, radii, skeleton, p = thinning_skeletal_points(n, seed, leaf_on, which)
        # _construct_graph ran inside the pipeline; redo the graph stage alone to capture _edges (the pipeline clears them)
        q = refshim.reference_graph_stage(sk, radii)
        assert list(q._skeleton.edges) == list(skeleton.edges)
        np.savez_compressed(os.path.join(HERE, f"graph_{name}_input.npz"), points=pts)
        dump(name, sk, radii, q)
    lat = lattice_cloud()
    dump("lattice", lat, None, refshim.reference_graph_stage(lat))


if __name__ == "__main__":
    main()
