"""ctypes binding of ``libsqsm_b200.so`` (C ABI in ``include/sqsm_b200.h``).

There is no CPU fallback: importing this module without the built library raises, and
``Engine`` raises when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsqsm_b200.so")


class SqsmConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("batch_size", C.c_int32), ("max_iteration", C.c_int32),
        ("voxel_size", C.c_float), ("cube_size", C.c_float), ("buffer_size", C.c_float),
        ("monitored_by_chamfer_distance", C.c_int32), ("latch_early_stop", C.c_int32),
        ("down_kernel_flip", C.c_int32), ("conv_mode", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class SqsmIterStats(C.Structure):
    _fields_ = [
        ("n_in", C.c_int64), ("n_entries", C.c_int64), ("n_vox", C.c_int64),
        ("n_rows_l1", C.c_int64), ("n_rows_l2", C.c_int64), ("n_rows_l3", C.c_int64),
        ("n_out", C.c_int64), ("n_cubes", C.c_int64), ("n_batches", C.c_int64), ("far_face_rows", C.c_int64),
        ("pairs_l0", C.c_int64), ("pairs_l1", C.c_int64), ("pairs_l2", C.c_int64), ("pairs_l3", C.c_int64),
        ("chamfer", C.c_double), ("accepted", C.c_int32), ("skipped", C.c_int32), ("epoch", C.c_int32),
        ("gpu_launches", C.c_int32),
        ("ms_tiling", C.c_float), ("ms_voxel", C.c_float), ("ms_rulebook", C.c_float), ("ms_network", C.c_float),
        ("ms_chamfer", C.c_float),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class SqsmPrepConfig(C.Structure):
    _fields_ = [
        ("voxel_down_size", C.c_double), ("std_ratio", C.c_double), ("filter", C.c_int32), ("nb_neighbors", C.c_int32),
        ("density_adaptation", C.c_int32), ("keep_debug", C.c_int32), ("radius", C.c_double), ("nb_points", C.c_int32),
        ("reserved", C.c_int32 * 1),
    ]


class SqsmPrepStats(C.Structure):
    _fields_ = [
        ("n_in", C.c_int64), ("n_voxel", C.c_int64), ("n_kept", C.c_int64), ("gpu_launches", C.c_int64),
        ("n_far_filter", C.c_int64), ("n_far_scale", C.c_int64), ("global_shift", C.c_double * 3), ("scale", C.c_double), ("mean_nn", C.c_double),
        ("cloud_mean", C.c_double), ("std_dev", C.c_double), ("threshold", C.c_double),
        ("ms_total", C.c_float), ("ms_downsample", C.c_float), ("ms_filter", C.c_float), ("ms_scale", C.c_float),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["global_shift"] = np.array(list(self.global_shift), dtype=np.float64)
        return d


EXPORTS = [
    "sqsm_create", "sqsm_destroy", "sqsm_last_error", "sqsm_abi_version", "sqsm_set_weight",
    "sqsm_finalize_weights", "sqsm_set_points_f64", "sqsm_set_points_dev_f32", "sqsm_set_state_f32", "sqsm_contract_step",
    "sqsm_step_begin", "sqsm_step_segment", "sqsm_step_set_candidate", "sqsm_step_chamfer", "sqsm_step_finish",
    "sqsm_result_size", "sqsm_get_result", "sqsm_skeletal_sample", "sqsm_radius_graph",
    "sqsm_radius_graph_points", "sqsm_radius_graph_range", "sqsm_ball_lists", "sqsm_debug_capture", "sqsm_debug_level_size", "sqsm_debug_voxels",
    "sqsm_debug_batch_shapes", "sqsm_debug_level_tables", "sqsm_debug_canon_to_internal",
    "sqsm_debug_predictions", "sqsm_debug_features", "sqsm_profile_reset", "sqsm_profile_get",
    "sqsm_profile_enable", "sqsm_profile_families", "sqsm_timer_start", "sqsm_timer_stop", "sqsm_launch_count",
    "sqsm_get_stream", "sqsm_graph_construct", "sqsm_graph_edges", "sqsm_graph_info", "sqsm_graph_extract_skeleton",
    "sqsm_prepare_cloud_f64", "sqsm_prepared_size", "sqsm_get_prepared", "sqsm_set_points_prepared", "sqsm_debug_prepared", "sqsm_sizeof",
]

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load the CUDA library; raise loudly if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SQSM_LIB") or LIB_PATH              # SQSM_LIB: A/B runs of an alternative build (tuning only)
    if not os.path.isfile(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m smartqsm_b200.build` (nvcc, sm_100a). "
            "smartqsm_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp, cp, i64, i32, f64 = C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.c_double
    P = C.POINTER
    sig = {
        "sqsm_create": ([P(SqsmConfig), P(vp)], C.c_int),
        "sqsm_destroy": ([vp], None),
        "sqsm_last_error": ([vp], cp),
        "sqsm_abi_version": ([], C.c_int),
        "sqsm_set_weight": ([vp, cp, vp, i64], C.c_int),
        "sqsm_finalize_weights": ([vp], C.c_int),
        "sqsm_set_points_f64": ([vp, vp, i64], C.c_int),
        "sqsm_set_points_dev_f32": ([vp, vp, i64], C.c_int),
        "sqsm_set_state_f32": ([vp, vp, vp, i64], C.c_int),
        "sqsm_contract_step": ([vp, P(SqsmIterStats)], C.c_int),
        "sqsm_step_begin": ([vp, i32, i32, P(SqsmIterStats)], C.c_int),
        "sqsm_step_segment": ([vp, P(i64), vp, vp], C.c_int),
        "sqsm_step_set_candidate": ([vp, vp, vp, i64], C.c_int),
        "sqsm_step_chamfer": ([vp, i32, i32, P(f64), P(i64)], C.c_int),
        "sqsm_step_finish": ([vp, i32, f64, P(i32)], C.c_int),
        "sqsm_result_size": ([vp, P(i64)], C.c_int),
        "sqsm_get_result": ([vp, vp, vp], C.c_int),
        "sqsm_skeletal_sample": ([vp, f64, P(i64), vp, vp, vp], C.c_int),
        "sqsm_radius_graph": ([vp, f64, P(i64), vp], C.c_int),
        "sqsm_radius_graph_points": ([vp, vp, i64, f64, P(i64), vp, i64], C.c_int),
        "sqsm_radius_graph_range": ([vp, f64, i64, i64, P(i64), vp, i64], C.c_int),
        "sqsm_ball_lists": ([vp, vp, i64, vp, vp, P(i64), vp, i64], C.c_int),
        "sqsm_debug_capture": ([vp, i32], C.c_int),
        "sqsm_debug_level_size": ([vp, i32, P(i64)], C.c_int),
        "sqsm_debug_voxels": ([vp, vp, vp, vp], C.c_int),
        "sqsm_debug_batch_shapes": ([vp, vp, vp], C.c_int),
        "sqsm_debug_level_tables": ([vp, i32, vp, vp, vp, vp], C.c_int),
        "sqsm_debug_canon_to_internal": ([vp, vp], C.c_int),
        "sqsm_debug_predictions": ([vp, vp, vp, vp], C.c_int),
        "sqsm_debug_features": ([vp, cp, P(i32), P(i32), vp, i64], C.c_int),
        "sqsm_profile_reset": ([vp], C.c_int),
        "sqsm_profile_get": ([vp, cp, P(f64), P(i64), P(f64), P(f64)], C.c_int),
        "sqsm_profile_families": ([vp, vp, i64], C.c_int),
        "sqsm_timer_start": ([vp], C.c_int),
        "sqsm_timer_stop": ([vp, P(f64)], C.c_int),
        "sqsm_launch_count": ([vp, P(i64)], C.c_int),
        "sqsm_profile_enable": ([vp, i32], C.c_int),
        "sqsm_prepare_cloud_f64": ([vp, vp, i64, P(SqsmPrepConfig), P(SqsmPrepStats)], C.c_int),
        "sqsm_prepared_size": ([vp, P(i64), P(i64)], C.c_int),
        "sqsm_get_prepared": ([vp, vp], C.c_int),
        "sqsm_set_points_prepared": ([vp], C.c_int),
        "sqsm_debug_prepared": ([vp, vp, vp, vp], C.c_int),
        "sqsm_sizeof": ([i32], C.c_int),
        "sqsm_get_stream": ([vp, P(vp)], C.c_int),
        "sqsm_graph_construct": ([vp, vp, i64, vp, i64, f64, vp, i64, P(i64)], C.c_int),
        "sqsm_graph_edges": ([vp, vp], C.c_int),
        "sqsm_graph_info": ([vp, vp], C.c_int),
        "sqsm_graph_extract_skeleton": ([vp, i32, P(i64), P(i64), vp, vp, vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """Thin object wrapper over one ``sqsm_handle``."""

    def __init__(self, *, device: int = 0, batch_size: int = 16, max_iteration: int = 10, voxel_size: float = 0.01,
                 cube_size: float = 4.0, buffer_size: float = 0.4, monitored_by_chamfer_distance: bool = False,
                 latch_early_stop: bool = False, down_kernel_flip: bool = False, conv_mode: int = 0) -> None:
        self._lib = load_library()
        cfg = SqsmConfig(device=device, batch_size=batch_size, max_iteration=max_iteration, voxel_size=voxel_size,
                         cube_size=cube_size, buffer_size=buffer_size,
                         monitored_by_chamfer_distance=int(monitored_by_chamfer_distance),
                         latch_early_stop=int(latch_early_stop), down_kernel_flip=int(down_kernel_flip),
                         conv_mode=int(conv_mode))
        h = C.c_void_p()
        rc = self._lib.sqsm_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError("sqsm_create failed: " + self._lib.sqsm_last_error(None).decode())
        self._h = h
        self._keepalive = None

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sqsm_destroy(self._h)
            self._h = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed: " + self._lib.sqsm_last_error(self._h).decode())

    # ---- weights
    def load_state_dict(self, state: Dict[str, np.ndarray]) -> None:
        for name, v in state.items():
            if name.endswith("num_batches_tracked"):
                continue
            a = np.ascontiguousarray(np.asarray(v), dtype=np.float32)
            self._check(self._lib.sqsm_set_weight(self._h, name.encode(), _ptr(a), a.size), f"set_weight({name})")
        self._check(self._lib.sqsm_finalize_weights(self._h), "finalize_weights")

    # ---- plug-in path
    def set_points(self, points: np.ndarray) -> None:
        a = np.ascontiguousarray(points, dtype=np.float64)
        assert a.ndim == 2 and a.shape[1] == 3
        self._check(self._lib.sqsm_set_points_f64(self._h, _ptr(a), a.shape[0]), "set_points")

    def set_points_ptr_f64(self, host_ptr: int, n: int) -> None:
        self._check(self._lib.sqsm_set_points_f64(self._h, C.c_void_p(host_ptr), n), "set_points")

    def set_points_device_f32(self, dev_ptr: int, n: int) -> None:
        self._check(self._lib.sqsm_set_points_dev_f32(self._h, C.c_void_p(dev_ptr), n), "set_points_dev")

    # ---- cloud preparation (row f2)
    def prepare_cloud(self, points: np.ndarray, *, voxel_down_size: float = 0.01, filter: int = 1, nb_neighbors: int = 10,
                      std_ratio: float = 4.0, density_adaptation: bool = True, keep_debug: bool = False,
                      nb_points: int = 0, radius: float = 0.0) -> dict:
        a = np.ascontiguousarray(points, dtype=np.float64)
        assert a.ndim == 2 and a.shape[1] == 3
        cfg = SqsmPrepConfig(voxel_down_size=voxel_down_size, std_ratio=std_ratio, filter=int(filter),
                             nb_neighbors=int(nb_neighbors), density_adaptation=int(density_adaptation),
                             keep_debug=int(keep_debug), radius=float(radius), nb_points=int(nb_points))
        st = SqsmPrepStats()
        self._check(self._lib.sqsm_prepare_cloud_f64(self._h, _ptr(a), a.shape[0], C.byref(cfg), C.byref(st)), "prepare_cloud")
        return st.as_dict()

    def prepared_points(self) -> np.ndarray:
        n = C.c_int64()
        self._check(self._lib.sqsm_prepared_size(self._h, C.byref(n), None), "prepared_size")
        out = np.empty((n.value, 3), dtype=np.float64)
        self._check(self._lib.sqsm_get_prepared(self._h, _ptr(out)), "get_prepared")
        return out

    def set_points_prepared(self) -> None:
        self._check(self._lib.sqsm_set_points_prepared(self._h), "set_points_prepared")

    def debug_prepared(self, with_filter: bool = True):
        n, nv = C.c_int64(), C.c_int64()
        self._check(self._lib.sqsm_prepared_size(self._h, C.byref(n), C.byref(nv)), "prepared_size")
        vox = np.empty((nv.value, 3), dtype=np.float64)
        avg = np.empty(nv.value, dtype=np.float64) if with_filter else None
        keep = np.empty(nv.value, dtype=np.uint8) if with_filter else None
        self._check(self._lib.sqsm_debug_prepared(self._h, _ptr(vox), _ptr(avg), _ptr(keep)), "debug_prepared")
        return vox, avg, keep

    def set_state(self, points_f32: np.ndarray, disp_f32: np.ndarray) -> None:
        p = np.ascontiguousarray(points_f32, dtype=np.float32)
        d = np.ascontiguousarray(disp_f32, dtype=np.float32)
        assert p.shape == d.shape and p.ndim == 2 and p.shape[1] == 3
        self._check(self._lib.sqsm_set_state_f32(self._h, _ptr(p), _ptr(d), p.shape[0]), "set_state")

    def contract_step(self) -> dict:
        st = SqsmIterStats()
        self._check(self._lib.sqsm_contract_step(self._h, C.byref(st)), "contract_step")
        return st.as_dict()

    # ---- phased thunk (one tree sharded over several ranks; device pointers are raw ints)
    def step_begin(self, shard: int, n_shards: int) -> dict:
        st = SqsmIterStats()
        self._check(self._lib.sqsm_step_begin(self._h, shard, n_shards, C.byref(st)), "step_begin")
        return st.as_dict()

    def step_segment(self, d_xyz: int = 0, d_disp: int = 0) -> int:
        n = C.c_int64()
        self._check(self._lib.sqsm_step_segment(self._h, C.byref(n), C.c_void_p(d_xyz or None), C.c_void_p(d_disp or None)),
                    "step_segment")
        return n.value

    def step_set_candidate(self, d_xyz: int, d_disp: int, n: int) -> None:
        self._check(self._lib.sqsm_step_set_candidate(self._h, C.c_void_p(d_xyz), C.c_void_p(d_disp), n), "step_set_candidate")

    def step_chamfer(self, shard: int, n_shards: int):
        sums = (C.c_double * 2)()
        counts = (C.c_int64 * 2)()
        self._check(self._lib.sqsm_step_chamfer(self._h, shard, n_shards, sums, counts), "step_chamfer")
        return [sums[0], sums[1]], [counts[0], counts[1]]

    def step_finish(self, monitored: bool, chamfer: float) -> bool:
        acc = C.c_int32()
        self._check(self._lib.sqsm_step_finish(self._h, int(monitored), float(chamfer), C.byref(acc)), "step_finish")
        return bool(acc.value)

    def result(self) -> Tuple[np.ndarray, np.ndarray]:
        n = C.c_int64()
        self._check(self._lib.sqsm_result_size(self._h, C.byref(n)), "result_size")
        p = np.empty((n.value, 3), np.float32)
        d = np.empty((n.value, 3), np.float32)
        self._check(self._lib.sqsm_get_result(self._h, _ptr(p), _ptr(d)), "get_result")
        return p, d

    def result_into(self, p_ptr: int, d_ptr: int) -> int:
        n = C.c_int64()
        self._check(self._lib.sqsm_result_size(self._h, C.byref(n)), "result_size")
        self._check(self._lib.sqsm_get_result(self._h, C.c_void_p(p_ptr), C.c_void_p(d_ptr)), "get_result")
        return n.value

    def result_size(self) -> int:
        n = C.c_int64()
        self._check(self._lib.sqsm_result_size(self._h, C.byref(n)), "result_size")
        return n.value

    def skeletal_sample(self, voxel: float = 0.01, fetch: bool = True):
        m = C.c_int64()
        self._check(self._lib.sqsm_skeletal_sample(self._h, voxel, C.byref(m), None, None, None), "skeletal_sample")
        if not fetch:
            return m.value
        ids = np.empty(m.value, np.int64)
        pts = np.empty((m.value, 3), np.float32)
        rad = np.empty(m.value, np.float32)
        self._check(self._lib.sqsm_skeletal_sample(self._h, voxel, C.byref(m), _ptr(ids), _ptr(pts), _ptr(rad)),
                    "skeletal_sample")
        return ids, pts, rad

    def radius_graph(self, r: float = 0.055, fetch: bool = True):
        """Edges over the last skeletal sample (computed once per sample and radius on the device;
        ``fetch=False`` returns the count and leaves the edge list in HBM)."""
        e = C.c_int64()
        self._check(self._lib.sqsm_radius_graph(self._h, r, C.byref(e), None), "radius_graph")
        if not fetch:
            return e.value
        out = np.empty((e.value, 2), np.int64)
        if e.value:
            self._check(self._lib.sqsm_radius_graph(self._h, r, C.byref(e), _ptr(out)), "radius_graph")
        return out

    def skeletal_sample_into(self, voxel: float, ids_ptr: int, pts_ptr: int, rad_ptr: int) -> int:
        """Copy the skeletal sample into caller-provided host buffers (e.g. pinned memory); returns M."""
        m = C.c_int64()
        self._check(self._lib.sqsm_skeletal_sample(self._h, voxel, C.byref(m), C.c_void_p(ids_ptr), C.c_void_p(pts_ptr),
                                                   C.c_void_p(rad_ptr)), "skeletal_sample")
        return m.value

    def radius_graph_into(self, r: float, edges_ptr: int) -> int:
        """Copy the edge list into a caller-provided host buffer of at least 16 * E bytes; returns E."""
        e = C.c_int64()
        self._check(self._lib.sqsm_radius_graph(self._h, r, C.byref(e), C.c_void_p(edges_ptr)), "radius_graph")
        return e.value

    def radius_graph_range(self, r: float, node_lo: int, node_hi: int, edges_ptr: int = 0, capacity: int = -1):
        """The slice of the edge list whose first node lies in [node_lo, node_hi) (one rank's share of a sharded tree).
        ``edges_ptr`` = 0 returns the slice as an array; otherwise it is copied into that host buffer and E is returned."""
        e = C.c_int64()
        if edges_ptr:
            self._check(self._lib.sqsm_radius_graph_range(self._h, r, node_lo, node_hi, C.byref(e), C.c_void_p(edges_ptr),
                                                          capacity), "radius_graph_range")
            return e.value
        self._check(self._lib.sqsm_radius_graph_range(self._h, r, node_lo, node_hi, C.byref(e), None, -1), "radius_graph_range")
        out = np.empty((e.value, 2), np.int64)
        if e.value:
            self._check(self._lib.sqsm_radius_graph_range(self._h, r, node_lo, node_hi, C.byref(e), _ptr(out), e.value),
                        "radius_graph_range")
        return out

    def radius_graph_range_count(self, r: float, node_lo: int, node_hi: int) -> int:
        """Build the slice on the device and return its length only (nothing is copied out)."""
        e = C.c_int64()
        self._check(self._lib.sqsm_radius_graph_range(self._h, r, node_lo, node_hi, C.byref(e), None, -1), "radius_graph_range")
        return e.value

    def ball_lists(self, pts: np.ndarray, radii: np.ndarray):
        """Row f3: ``cKDTree(pts).query_ball_point(pts, r=radii)`` for every node at once -> (offsets int64[M + 1],
        indices int32[T]); the ball of node i is ``indices[offsets[i]:offsets[i + 1]]``, ascending."""
        a = np.ascontiguousarray(pts, dtype=np.float32)
        r = np.ascontiguousarray(radii, dtype=np.float64)
        assert a.ndim == 2 and a.shape[1] == 3 and r.shape == (a.shape[0],)
        off = np.empty(a.shape[0] + 1, np.int64)
        t = C.c_int64()
        self._check(self._lib.sqsm_ball_lists(self._h, _ptr(a), a.shape[0], _ptr(r), _ptr(off), C.byref(t), None, 0), "ball_lists")
        idx = np.empty(t.value, np.int32)
        if t.value:
            self._check(self._lib.sqsm_ball_lists(self._h, _ptr(a), a.shape[0], _ptr(r), _ptr(off), C.byref(t), _ptr(idx), t.value),
                        "ball_lists")
        return off, idx

    def radius_graph_points(self, pts: np.ndarray, r: float) -> np.ndarray:
        a = np.ascontiguousarray(pts, dtype=np.float32)
        e = C.c_int64()
        self._check(self._lib.sqsm_radius_graph_points(self._h, _ptr(a), a.shape[0], r, C.byref(e), None, 0),
                    "radius_graph_points")
        out = np.empty((e.value, 2), np.int64)
        if e.value:
            self._check(self._lib.sqsm_radius_graph_points(self._h, _ptr(a), a.shape[0], r, C.byref(e), _ptr(out),
                                                           e.value), "radius_graph_points")
        return out

    # ---- connectivity graph + skeleton extraction (row a16 / f1)
    def graph_construct(self, simplices: np.ndarray, r: float = 0.055, points: Optional[np.ndarray] = None,
                        preset_edges: Optional[np.ndarray] = None, fetch: bool = True):
        """``_construct_graph`` (core/skeletonization_pipeline_base.py:75-174).  ``simplices`` = Delaunay tetrahedra of the
        skeletal points (``scipy.spatial.Delaunay(points).simplices``); ``points=None`` uses the last skeletal sample."""
        sx = np.ascontiguousarray(simplices, dtype=np.int32)
        assert sx.ndim == 2 and sx.shape[1] == 4
        pts = None if points is None else np.ascontiguousarray(points, dtype=np.float32)
        pre = None if preset_edges is None else np.ascontiguousarray(preset_edges, dtype=np.int64).reshape(-1, 2)
        e = C.c_int64()
        self._check(self._lib.sqsm_graph_construct(self._h, _ptr(pts), 0 if pts is None else pts.shape[0], _ptr(sx), sx.shape[0],
                                                   float(r) if r else 0.0, _ptr(pre), 0 if pre is None else pre.shape[0],
                                                   C.byref(e)), "graph_construct")
        if not fetch:
            return e.value
        out = np.empty((e.value, 2), np.int64)
        if e.value:
            self._check(self._lib.sqsm_graph_edges(self._h, _ptr(out)), "graph_edges")
        return out

    def graph_info(self) -> dict:
        a = np.zeros(5, np.int64)
        self._check(self._lib.sqsm_graph_info(self._h, _ptr(a)), "graph_info")
        return dict(zip(("delaunay_root", "n_delaunay_edges", "n_spt_edges", "n_mst_edges", "n_edges"), (int(x) for x in a)))

    def graph_extract_skeleton(self, n_nodes: int, edge_weight_function: str = "l1") -> dict:
        """``_extract_skeleton`` with ``topology_extractor="spt"`` (:190-215) on the graph of ``graph_construct``."""
        wf = {"l1": 0, "l2": 1}.get(edge_weight_function)
        if wf is None:
            raise NotImplementedError(f"edge weight function {edge_weight_function!r}: the device graph stage implements l1 and l2")
        root, nr = C.c_int64(), C.c_int64()
        pred = np.empty(n_nodes, np.int64)
        order = np.empty(max(n_nodes - 1, 0), np.int64)
        weight = np.empty(n_nodes, np.float32)
        self._check(self._lib.sqsm_graph_extract_skeleton(self._h, wf, C.byref(root), C.byref(nr), _ptr(pred), _ptr(order),
                                                          _ptr(weight)), "graph_extract_skeleton")
        return {"root": root.value, "pred": pred, "order": order[:nr.value - 1], "weight": weight}

    # ---- inspection hooks
    def debug_capture(self, enable: bool = True) -> None:
        self._check(self._lib.sqsm_debug_capture(self._h, int(enable)), "debug_capture")

    def debug_level_size(self, level: int) -> int:
        n = C.c_int64()
        self._check(self._lib.sqsm_debug_level_size(self._h, level, C.byref(n)), "debug_level_size")
        return n.value

    def debug_voxels(self):
        n = self.debug_level_size(0)
        idx = np.empty((n, 4), np.int32)
        win = np.empty(n, np.int64)
        mask = np.empty(n, np.uint8)
        self._check(self._lib.sqsm_debug_voxels(self._h, _ptr(idx), _ptr(win), _ptr(mask)), "debug_voxels")
        return idx, win, mask.astype(bool)

    def debug_batch_shapes(self, n_batches: int):
        shp = np.empty((n_batches, 3), np.int32)
        desc = np.empty((3, n_batches), np.int32)
        self._check(self._lib.sqsm_debug_batch_shapes(self._h, _ptr(shp), _ptr(desc)), "debug_batch_shapes")
        return shp, desc

    def debug_level_tables(self, level: int, with_down: bool):
        n = self.debug_level_size(level)
        coords = np.empty((n, 4), np.int32)
        nbr = np.empty((27, n), np.int32)
        parent = np.empty(n, np.int32) if with_down else None
        koff = np.empty(n, np.int32) if with_down else None
        self._check(self._lib.sqsm_debug_level_tables(self._h, level, _ptr(coords), _ptr(nbr), _ptr(parent), _ptr(koff)),
                    "debug_level_tables")
        return coords, nbr, parent, koff

    def debug_canon_to_internal(self) -> np.ndarray:
        n = self.debug_level_size(0)
        perm = np.empty(n, np.int32)
        self._check(self._lib.sqsm_debug_canon_to_internal(self._h, _ptr(perm)), "debug_canon_to_internal")
        return perm

    def debug_predictions(self):
        n = self.debug_level_size(0)
        pred = np.empty((n, 3), np.float32)
        off = np.empty((n, 3), np.float32)
        reg = np.empty(n, np.float32)
        self._check(self._lib.sqsm_debug_predictions(self._h, _ptr(pred), _ptr(off), _ptr(reg)), "debug_predictions")
        return pred, off, reg

    def debug_features(self, stage: str) -> Tuple[np.ndarray, int]:
        lvl, ch = C.c_int32(), C.c_int32()
        self._check(self._lib.sqsm_debug_features(self._h, stage.encode(), C.byref(lvl), C.byref(ch), None, 0),
                    "debug_features")
        n = self.debug_level_size(lvl.value)
        out = np.empty((n, ch.value), np.float32)
        self._check(self._lib.sqsm_debug_features(self._h, stage.encode(), C.byref(lvl), C.byref(ch), _ptr(out), out.size),
                    "debug_features")
        return out, lvl.value

    # ---- measurement hooks
    def profile_enable(self, enable: bool = True) -> None:
        self._check(self._lib.sqsm_profile_enable(self._h, int(enable)), "profile_enable")

    def profile_reset(self) -> None:
        self._check(self._lib.sqsm_profile_reset(self._h), "profile_reset")

    def profile_get(self, family: str) -> dict:
        ms, n, by, fl = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        self._check(self._lib.sqsm_profile_get(self._h, family.encode(), C.byref(ms), C.byref(n), C.byref(by),
                                               C.byref(fl)), "profile_get")
        return {"ms": ms.value, "launches": n.value, "bytes": by.value, "flops": fl.value}

    def profile_all(self) -> Dict[str, dict]:
        buf = C.create_string_buffer(4096)
        self._check(self._lib.sqsm_profile_families(self._h, buf, 4096), "profile_families")
        names = [n for n in buf.value.decode().split(",") if n]
        return {n: self.profile_get(n) for n in names}

    def timer_start(self) -> None:
        self._check(self._lib.sqsm_timer_start(self._h), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self._lib.sqsm_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return ms.value

    def stream_ptr(self) -> int:
        """The engine's ``cudaStream_t`` as an integer (for ``torch.cuda.ExternalStream``)."""
        p = C.c_void_p()
        self._check(self._lib.sqsm_get_stream(self._h, C.byref(p)), "get_stream")
        return int(p.value or 0)

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.sqsm_launch_count(self._h, C.byref(n)), "launch_count")
        return n.value
