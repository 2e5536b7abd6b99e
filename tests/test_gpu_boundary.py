"""Drop-in boundary through the REFERENCE's own code (``-m gpu``).

``ThinningBasedSkeletonizationPipeline`` (core/thinning_based_skeletonization_pipeline.py:47-102) is imported unmodified
(``oracle/refshim``: stand-ins only for the wheels that cannot be installed) and resolves ``core_algorithm`` through the
reference's registry (``:55-58``), into which ``smartqsm_b200.register_into`` has put the B200 plug-in.  The pipeline then
drives it exactly like the application does (``smartqsm.py:928-938``): ``input(points=)``, the ``max_iteration`` thunks,
``output()``, its own ``_produce_skeletal_points_and_radii``, ``_construct_graph``, ``_extract_skeleton``.
"""
import os

import numpy as np
import pytest

from oracle import contraction as OC
from oracle import graph as OG
from oracle import refshim, weights

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(pipeline, pts):
    gen = pipeline.run(points=pts)
    steps = 0
    try:
        while True:
            assert next(gen) is None                     # non-verbose thunks yield None (drives the GUI progress bar)
            steps += 1
    except StopIteration as stop:
        return stop.value, steps


def _pipeline(use_device_graph: bool):
    from smartqsm_b200.core_algorithms import SpconvBasedContraction, register_into
    blob = weights.blob_path("leafoff")
    kwargs = dict(ckpt_path=blob, batch_size=16, max_iteration=10, device="cuda", voxel_size=0.01, cube_size=4.0,
                  buffer_size=0.4, monitored_by_chamfer_distance=True)          # configs/spconv-contraction-LEAFOFF-GPU.yaml:10-18
    p = refshim.reference_thinning_pipeline(registry_patch=register_into, core_algorithm="spconv_based_contraction",
                                            core_algorithm_kwargs=kwargs, voxel_size_for_downsampling=0.01,
                                            edge_search_kwargs={"r": 0.055}, edge_weight_function="l1", topology_extractor="spt")
    assert isinstance(p._core_algorithm, SpconvBasedContraction), "the registry must resolve to the B200 plug-in"
    assert len(p) == 10 + 3                                # progress-bar length: thunks + sampling + graph + skeleton
    if use_device_graph:
        from smartqsm_b200.skeleton_graph import install_into
        install_into(p, p._core_algorithm.engine)
    return p


def test_reference_pipeline_drives_the_plugin():
    if not refshim.available() or weights.blob_path("leafoff") is None:
        pytest.skip("reference Python / checkpoint blob not present (baseline/_ref; run build() in the authoring container)")
    g_in = np.load(os.path.join(GOLD, "graph_tree_leafoff_input.npz"))["points"]
    p = _pipeline(False)
    algo = p._core_algorithm
    (sk, skeleton, radii), steps = _run(p, g_in)
    assert steps == 13
    out = algo.output()
    # a14 ran in the reference's own Python on the plug-in's output: equals the restatement and the device operator
    ids_o, sk_o, rad_o = OC.skeletal_points_and_radii(out["contracted_points"], out["displacements"], 0.01)
    assert len(sk) == len(sk_o)
    assert np.array_equal(np.sort(radii), np.sort(rad_o))
    ids_d, sk_d, rad_d = algo.produce_skeletal_points_and_radii(0.01)
    assert np.array_equal(ids_d, ids_o) and np.array_equal(sk_d, sk_o) and np.array_equal(rad_d, rad_o)
    # a15 + a16: the skeleton the reference built on these skeletal points equals the restatement on the same points
    cg = OG.construct_graph(np.asarray(sk, np.float32))
    so = OG.extract_skeleton(np.asarray(sk, np.float32), cg["edges"])
    go = OG.skeleton_as_digraph(np.asarray(sk, np.float32), so)
    assert list(skeleton.edges) == list(go.edges) and skeleton.number_of_nodes() == go.number_of_nodes()
    # against the golden vectors of the reference's CPU run on the same input (stub spconv): same scale of result
    g = np.load(os.path.join(GOLD, "graph_tree_leafoff.npz"))
    assert abs(len(sk) - len(g["skeletal_points"])) <= max(3, len(g["skeletal_points"]) // 500)


def test_reference_pipeline_with_device_graph_stage():
    """The same pipeline object with ``_construct_graph`` / ``_extract_skeleton`` swapped for the device stage (row f1):
    identical skeleton, node for node and edge for edge, weights included."""
    if not refshim.available() or weights.blob_path("leafoff") is None:
        pytest.skip("reference Python / checkpoint blob not present")
    g_in = np.load(os.path.join(GOLD, "graph_tree_leafoff_input.npz"))["points"]
    (sk_a, skel_a, rad_a), _ = _run(_pipeline(False), g_in)
    (sk_b, skel_b, rad_b), _ = _run(_pipeline(True), g_in)
    assert np.array_equal(sk_a, sk_b) and np.array_equal(rad_a, rad_b)
    assert list(skel_a.nodes) == list(skel_b.nodes) and list(skel_a.edges) == list(skel_b.edges)
    assert all(skel_a[u][v]["weight"] == skel_b[u][v]["weight"] for u, v in skel_a.edges)
