"""Host-side mirror of the reference plug-in interface for the ``spconv-contraction`` path.

``CoreAlgorithmBase`` and ``registry`` follow
``core/core_algorithms/core_algorithm_base.py:21-37`` and ``core/core_algorithms/_registry.py:25-33``;
``SpconvBasedContraction`` has the constructor, ``input`` / ``get_pipeline`` / ``output`` contract
and error behaviour of ``core/core_algorithms/spconv_based_contraction.py:34-151`` but runs every
stage on a B200 through ``libsqsm_b200.so``.  PyTorch is not needed on this path (only to read a
``.pth`` checkpoint).  No CPU fallback: ``device="cpu"`` raises.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from ._lib import Engine

__all__ = ["CoreAlgorithmBase", "SpconvBasedContraction", "registry", "register_into", "load_checkpoint"]


class CoreAlgorithmBase:
    """``core/core_algorithms/core_algorithm_base.py:21-37``."""
    _verbose: bool

    def __init__(self, *, verbose: bool = False, **kwargs) -> None:
        self._verbose = verbose

    def get_pipeline(self) -> List[Callable[[], Any]]:
        pass

    def input(self, **kwargs) -> None:
        for k, v in kwargs.items():
            setattr(self, f"_{k}", v)

    def output(self) -> Dict[str, Any]:
        pass


def _smarttreexx_dir() -> Optional[str]:
    """``os.path.dirname(external.SmartTreeXX.__file__)`` when a SmartQSM checkout is importable
    (spconv_based_contraction.py:60-61); ``SMARTQSM_ROOT`` may point at it explicitly."""
    root = os.environ.get("SMARTQSM_ROOT")
    if root and os.path.isdir(os.path.join(root, "external", "SmartTreeXX")):
        return os.path.join(root, "external", "SmartTreeXX")
    try:
        import external.SmartTreeXX as m  # type: ignore
        return os.path.dirname(m.__file__)
    except Exception:
        return None


def load_checkpoint(ckpt_path: str) -> Dict[str, np.ndarray]:
    """Read ``{"model_state_dict", "hparams"}`` written by ``external/SmartTreeXX/train.py:184-191``
    (``.pth``, needs torch) or a neutral ``.npz`` with the same tensor names."""
    if not os.path.isabs(ckpt_path):
        base = _smarttreexx_dir()
        cand = [os.path.abspath(os.path.join(base, ckpt_path))] if base else []
        cand.append(os.path.abspath(ckpt_path))
        for c in cand:
            if os.path.isfile(c):
                ckpt_path = c
                break
        else:
            raise FileNotFoundError(f"checkpoint {ckpt_path!r} not found (tried {cand})")
    if ckpt_path.endswith(".npz"):
        with np.load(ckpt_path) as z:
            return {k: z[k] for k in z.files}
    import torch
    ck = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    hp = ck["hparams"]
    if list(hp["u_channels"]) != [8, 16, 32, 64] or list(hp["mlp_hidden_layers"]) != [8, 4] \
            or hp.get("conv_block_reps", 1) != 1:
        raise ValueError(f"unsupported SmartTreeXX hparams {hp}: kernels are specialised for "
                         "u_channels=[8,16,32,64], mlp_hidden_layers=[8,4], conv_block_reps=1")
    return {k: v.detach().cpu().numpy() for k, v in ck["model_state_dict"].items()}


def _device_ordinal(device: str) -> int:
    d = str(device)
    if d == "cuda":
        return int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else 0
    if d.startswith("cuda:"):
        return int(d.split(":", 1)[1])
    raise RuntimeError(f"device={device!r}: the B200 implementation has no CPU path; use device='cuda[:N]'")


class SpconvBasedContraction(CoreAlgorithmBase):
    """Drop-in for ``core/core_algorithms/spconv_based_contraction.py:34-151`` on one B200."""

    def __init__(self, *, ckpt_path: Optional[str] = None, batch_size: int, max_iteration: int, device: str = "cuda",
                 voxel_size: float = 0.01, cube_size: float = 4.0, buffer_size: float = 0.4, verbose: bool = False,
                 monitored_by_chamfer_distance: bool = False, state_dict: Optional[Dict[str, np.ndarray]] = None,
                 latch_early_stop: Optional[bool] = None, conv_mode: int = 0) -> None:
        super().__init__(verbose=verbose)
        self._verbose = verbose
        self._max_iteration = max_iteration
        self._epoch = 0
        self._cube_size, self._buffer_size, self._voxel_size = cube_size, buffer_size, voxel_size
        self._device, self._batch_size = device, batch_size
        # the reference latches the early stop only in verbose mode (F10, :123-127); latching never
        # changes results, so it can be forced on to save the recomputation
        latch = verbose if latch_early_stop is None else latch_early_stop
        self._engine = Engine(device=_device_ordinal(device), batch_size=batch_size, max_iteration=max_iteration,
                              voxel_size=voxel_size, cube_size=cube_size, buffer_size=buffer_size,
                              monitored_by_chamfer_distance=monitored_by_chamfer_distance, latch_early_stop=latch,
                              conv_mode=conv_mode)
        if state_dict is None:
            if ckpt_path is None:
                raise ValueError("ckpt_path (or state_dict) is required")
            state_dict = load_checkpoint(ckpt_path)
        self._engine.load_state_dict(state_dict)
        self._chamfer_distance: Optional[float] = float("inf") if monitored_by_chamfer_distance else None
        self._early_stopped = False
        self._points: Optional[np.ndarray] = None
        self._loaded = False
        self._prepared = False
        self._host: Optional[Tuple[np.ndarray, np.ndarray]] = None
        self.history: List[dict] = []

    def input(self, **kwargs) -> None:
        pts = kwargs.get("points")
        if pts is not None and hasattr(pts, "global_shift") and hasattr(pts, "scale"):
            # a PreparedCloud: `skeletonization.run(points=prepared)` keeps the device-resident hand-over of
            # prepare_input(); one made on another engine is scaled on the host like the reference does (:922)
            if self._prepared and pts.engine is self._engine:
                return
            kwargs = dict(kwargs, points=pts.points * pts.scale)
        super().input(**kwargs)
        if "points" in kwargs:
            self._loaded = False
            self._prepared = False
            self._epoch = 0
            self._host = None
            self.history = []

    def prepare_input(self, raw_points: np.ndarray, *, voxel_down_size: float = 0.01, filter: Optional[str] = None,
                      kwargs_for_filter: Optional[Dict[str, Any]] = None,
                      using_density_adaptation_in_skeletonization: bool = False):
        """``entrypoints/smartqsm.py:882-922`` on this algorithm's engine: shift, voxel
        down-sampling, outlier filter, density scale, then ``input(points = prepared * scale)`` without the host
        round trip.  Returns the :class:`~smartqsm_b200.preprocess.PreparedCloud` (``points``, ``global_shift``,
        ``scale``) the entry point needs afterwards (``skeletal_points / scale``, ``:940-941``)."""
        from .preprocess import prepare_cloud
        prepared = prepare_cloud(raw_points, voxel_down_size=voxel_down_size, filter=filter,
                                 kwargs_for_filter=kwargs_for_filter,
                                 using_density_adaptation_in_skeletonization=using_density_adaptation_in_skeletonization,
                                 engine=self._engine)
        self._points = None
        self._loaded = False
        self._prepared = True
        self._epoch = 0
        self._host = None
        self.history = []
        return prepared

    def _contract(self):
        self._epoch += 1
        if not self._loaded:                       # epoch 1: float64 -> float32 copy, zero displacements (:81-83)
            if self._prepared:
                self._engine.set_points_prepared()
            elif self._points is None:
                raise RuntimeError("input(points=...) was not called")
            else:
                self._engine.set_points(self._points)
            self._loaded = True
        st = self._engine.contract_step()
        self.history.append(st)
        self._host = None
        if st["skipped"]:
            if not self._verbose:
                return
            return f"Skipped the {self._epoch}th contraction.", None
        if not np.isnan(st["chamfer"]) and st["accepted"]:
            self._chamfer_distance = st["chamfer"]
        if not st["accepted"]:
            if not self._verbose:
                return
            self._early_stopped = True
            return f"Early stopped at the {self._epoch}th contraction.", None
        if not self._verbose:
            return
        cloud = None
        try:                                      # the reference hands an Open3D cloud to the GUI (:135-140)
            import open3d as o3d  # type: ignore
            cloud = o3d.geometry.PointCloud()
            cloud.points = o3d.utility.Vector3dVector(self.output()["contracted_points"])
        except Exception:
            cloud = None
        return f"Completed {self._epoch} contraction(s). {st['n_out']} point(s) remained.", cloud

    def get_pipeline(self):
        return [self._contract] * self._max_iteration

    def output(self) -> Dict[str, np.ndarray]:
        if self._host is None:
            self._host = self._engine.result()
        return {"contracted_points": self._host[0], "displacements": self._host[1]}

    # ---- the two downstream operators of the path that also run on the device ----------------
    def produce_skeletal_points_and_radii(self, voxel_size_for_downsampling: float = 0.01):
        """``ThinningBasedSkeletonizationPipeline._produce_skeletal_points_and_radii``
        (core/thinning_based_skeletonization_pipeline.py:68-80): ``(sample_ids, skeletal_points, radii)``."""
        return self._engine.skeletal_sample(voxel_size_for_downsampling)

    def radius_edges(self, r: float = 0.055) -> np.ndarray:
        """Radius-neighbour part of ``_construct_graph`` (core/skeletonization_pipeline_base.py:132-144,158)."""
        return self._engine.radius_graph(r)

    @property
    def engine(self) -> Engine:
        return self._engine


registry: Dict[str, Dict[str, type]] = {
    "thinning": {
        "spconv_based_contraction": SpconvBasedContraction,
        "b200_spconv_contraction": SpconvBasedContraction,
    },
}


def register_into(reference_registry: Dict[str, Dict[str, type]], shadow: bool = True) -> None:
    """Install the B200 implementation into the reference's ``registry`` (``_registry.py:25-33``)."""
    reference_registry.setdefault("thinning", {})["b200_spconv_contraction"] = SpconvBasedContraction
    if shadow:
        reference_registry["thinning"]["spconv_based_contraction"] = SpconvBasedContraction
