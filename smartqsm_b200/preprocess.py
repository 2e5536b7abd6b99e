"""Host-side mirror of the cloud preparation the reference entry point performs before it hands
``points * scale`` to the skeletonisation (``entrypoints/smartqsm.py:882-913,922``) - SURVEY.md §8 row f2.

``prepare_cloud`` takes the raw float64 cloud and the keys of the reference's YAML config
(``voxel_down_size``, ``filter``, ``kwargs_for_filter``, ``using_density_adaptation_in_skeletonization``)
and runs bounding box -> global shift -> translate -> voxel down-sampling (float64 voxel means) ->
statistical / radius outlier removal -> density-adaptation scale on the GPU through ``libsqsm_b200.so``.
No CPU fallback.  ``SpconvBasedContraction.prepare_input(raw_points, ...)`` runs the same preparation on the
algorithm's own engine and feeds ``points * scale`` to the contraction without a host round trip.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from ._lib import Engine

__all__ = ["prepare_cloud", "PreparedCloud"]

_FILTERS = {None: 0, "statistical_outlier_removal": 1, "radius_outlier_removal": 2}


@dataclass
class PreparedCloud:
    """What ``smartqsm.py`` holds after line 913: ``points`` (shifted, down-sampled, filtered, float64,
    not scaled), ``global_shift`` (:883-887), ``scale`` (:908-913) and the filter statistics."""
    global_shift: np.ndarray
    scale: float
    stats: Dict[str, Any]
    _engine: Engine = field(repr=False)
    _points: Optional[np.ndarray] = field(default=None, repr=False)

    @property
    def points(self) -> np.ndarray:
        if self._points is None:
            self._points = self._engine.prepared_points()
        return self._points

    @property
    def engine(self) -> Engine:
        return self._engine


def prepare_cloud(points: np.ndarray, *, voxel_down_size: float = 0.01, filter: Optional[str] = None,
                  kwargs_for_filter: Optional[Dict[str, Any]] = None,
                  using_density_adaptation_in_skeletonization: bool = False, device: int = 0,
                  engine: Optional[Engine] = None, keep_debug: bool = False, **engine_kwargs) -> PreparedCloud:
    """``entrypoints/smartqsm.py:882-913`` on the GPU.  Keyword names are the reference's config keys."""
    if filter not in _FILTERS:
        raise NotImplementedError(f"Filter {filter} is not implemented.")                    # :903-904
    kw = dict(kwargs_for_filter or {})
    eng = engine if engine is not None else Engine(device=device, **engine_kwargs)
    st = eng.prepare_cloud(points, voxel_down_size=float(voxel_down_size), filter=_FILTERS[filter],
                           nb_neighbors=int(kw.get("nb_neighbors", 10)), std_ratio=float(kw.get("std_ratio", 4.0)),
                           nb_points=int(kw.get("nb_points", 0)), radius=float(kw.get("radius", 0.0)),
                           density_adaptation=bool(using_density_adaptation_in_skeletonization), keep_debug=keep_debug)
    return PreparedCloud(global_shift=st["global_shift"], scale=float(st["scale"]), stats=st, _engine=eng)
