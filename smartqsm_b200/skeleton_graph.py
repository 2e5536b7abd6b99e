"""Host-side mirror of the graph half of ``SkeletonizationPipelineBase`` (SURVEY.md §8 row a16 = f1).

``construct_graph`` / ``extract_skeleton`` have the inputs, outputs and error behaviour of
``core/skeletonization_pipeline_base.py:75-174`` and ``:190-238`` (``topology_extractor="spt"``); the work runs in
``libsqsm_b200.so`` (``csrc/graph.cu``): device radix sorts / uniques / weights / CSR, host C++ for the two priority-queue
loops and the union-find whose networkx semantics are sequential by definition.  The Delaunay tetrahedra come from Qhull
through SciPy on the host, exactly like the reference obtains them (``:77``).

``install_into(pipeline)`` swaps the two thunks of a reference pipeline object for these, so that
``ThinningBasedSkeletonizationPipeline.run`` hands ``(skeletal_points, skeleton: nx.DiGraph, radii)`` to refinement
unchanged (``:283``).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from ._lib import Engine

__all__ = ["construct_graph", "extract_skeleton", "skeleton_to_digraph", "install_into"]


def construct_graph(engine: Engine, skeletal_points: Optional[np.ndarray], edge_search_kwargs: Dict[str, Any],
                    preset_edges: Optional[np.ndarray] = None, voxel_size_for_downsampling: float = 0.01) -> np.ndarray:
    """``_construct_graph`` (``:75-158``): merged edge list ``int64[E,2]``, ``i < j``, lexicographic.

    ``skeletal_points=None`` uses the engine's skeletal sample at ``voxel_size_for_downsampling`` (device resident; its
    radius graph is computed once and reused)."""
    from scipy.spatial import Delaunay
    if "k" in edge_search_kwargs:
        raise NotImplementedError("k-nearest-neighbour edges (edge_search_kwargs['k'], :146-156) are not on the spconv-contraction "
                                  "path (configs/spconv-contraction-*.yaml use 'r' only)")
    pts = skeletal_points
    if pts is None:
        _, pts, _ = engine.skeletal_sample(voxel_size_for_downsampling)
    simplices = Delaunay(pts).simplices                      # :77 - raises QhullError on too few points like the reference
    return engine.graph_construct(simplices, r=edge_search_kwargs.get("r", 0.0),
                                  points=None if skeletal_points is None else np.asarray(skeletal_points, np.float32),
                                  preset_edges=preset_edges)


def extract_skeleton(engine: Engine, n_nodes: int, edge_weight_function: str = "l1", topology_extractor: str = "spt") -> dict:
    """``_extract_skeleton`` (``:190-215``): ``{"root", "pred", "order", "weight"}`` (see ``sqsm_graph_extract_skeleton``)."""
    if topology_extractor != "spt":
        raise NotImplementedError(f"Topology extractor {topology_extractor} is not implemented.")        # :222
    return engine.graph_extract_skeleton(n_nodes, edge_weight_function)


def skeleton_to_digraph(sk: dict):
    """The ``nx.DiGraph`` of ``:199-213`` with the reference's node / edge insertion order and float32 weights."""
    import networkx as nx
    g = nx.DiGraph()
    pred, w = sk["pred"], sk["weight"]
    for n in sk["order"]:
        g.add_edge(int(pred[n]), int(n), weight=w[n])
    return g


def install_into(pipeline, engine: Engine) -> None:
    """Replace ``pipeline._construct_graph`` / ``_extract_skeleton`` (a ``SkeletonizationPipelineBase`` instance, AFTER
    ``set_params``) by the device stage.  The pipeline's own attributes are read and written like the originals do."""
    def _construct_graph():
        edges = construct_graph(engine, np.asarray(pipeline._skeletal_points, np.float32), pipeline._edge_search_kwargs,
                                None if pipeline._edges is None else np.asarray(pipeline._edges, np.int64))
        pipeline._edges = edges.tolist()
        pipeline._graph = None                               # the weighted nx.Graph is never materialised
        return None

    def _extract_skeleton():
        sk = extract_skeleton(engine, len(pipeline._skeletal_points), pipeline._edge_weight_function, pipeline._topology_extractor)
        pipeline._skeleton = skeleton_to_digraph(sk)
        return None

    fns = pipeline._pipeline
    for i, f in enumerate(fns):
        name = getattr(f, "__name__", "")
        if name == "_construct_graph":
            fns[i] = _construct_graph
        elif name == "_extract_skeleton":
            fns[i] = _extract_skeleton
