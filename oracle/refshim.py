"""Run the reference's OWN Python on top of stub ``spconv`` / ``open3d`` modules (test infrastructure).

The reference's hot path imports three un-vendored wheels that cannot be installed here (spconv,
cumm, open3d).  Everything ELSE on the path is the reference's own Python: cube tiling, mask logic and
collation (``external/SmartTreeXX/datasets/synthetic_tree.py``), the U-Net wiring, heads and decode
(``external/SmartTreeXX/models/*.py``) and the contraction / Chamfer driver
(``core/core_algorithms/spconv_based_contraction.py``).  ``install()`` registers minimal stand-ins for the
third-party entry points those files call, implemented with the oracle's restated primitives
(``oracle/tiling.py``, ``oracle/network.py``, ``oracle/contraction.py``), so that the reference files can
be imported from ``/root/reference`` and executed UNMODIFIED.  ``tests/golden/make_golden.py`` uses this
to produce golden vectors; agreement pins every piece of the restatement that is the reference's own
code.  It does not pin the third-party semantics themselves (see ``oracle/__init__.py``).

Only usable where ``/root/reference`` exists (the authoring container).
"""
from __future__ import annotations

import os
import sys
import types
from collections import OrderedDict
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import contraction as OC
from . import network as ON
from . import tiling as OT

def _reference_root() -> str:
    """The reference checkout (authoring container) or its unmodified copy under ``baseline/_ref/SmartQSM``
    (``oracle.weights.install_reference_python``; git-ignored, travels to the GPU box)."""
    if os.path.isdir("/root/reference/external/SmartTreeXX"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "SmartQSM")


REFERENCE_ROOT = _reference_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "external", "SmartTreeXX", "models")) and \
        os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


# ------------------------------------------------------------------------------------ spconv.pytorch

class SparseConvTensor:
    """spconv.pytorch.SparseConvTensor as used at synthetic_tree.py:198-203 and sparse_module.py:33,103."""

    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict: Optional[dict] = None):
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    def _np_indices(self) -> np.ndarray:
        return self.indices.detach().cpu().numpy().astype(np.int32)

    def _shape(self) -> List[int]:
        return [int(s) for s in self.spatial_shape]


class SparseModule(nn.Module):
    """spconv.pytorch.modules.SparseModule: marker base class."""


def _is_sparse(m: nn.Module) -> bool:
    return isinstance(m, (SparseModule, SparseSequential, _ConvBase))


class SparseSequential(SparseModule):
    """spconv.pytorch.SparseSequential: sparse modules get the tensor, dense ones its features."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)
        for k, m in kwargs.items():
            self.add_module(k, m)

    def add(self, module, name=None):
        self.add_module(str(len(self._modules)) if name is None else name, module)

    def __iter__(self):
        return iter(self._modules.values())

    def __len__(self):
        return len(self._modules)

    def forward(self, x):
        for m in self._modules.values():
            if _is_sparse(m):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


class _ConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True,
                 indice_key=None):
        super().__init__()
        assert not bias, "the reference builds every sparse convolution with bias=False"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = [kernel_size] * 3
        self.stride = [stride] * 3
        self.padding = [padding] * 3
        self.dilation = [dilation] * 3
        self.indice_key = indice_key
        # spconv 2.x weight layout [C_out, kz, ky, kx, C_in] (checkpoint shapes)
        self.weight = nn.Parameter(torch.zeros(out_channels, kernel_size, kernel_size, kernel_size, in_channels))


class SubMConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        if self.kernel_size[0] == 1:
            return x.replace_feature(ON.subm_conv(x.features, self.weight, None))
        key = ("subm", self.indice_key)
        if self.indice_key is None or key not in x.indice_dict:
            nbr = ON.subm_rulebook(x._np_indices(), x._shape())
            if self.indice_key is not None:
                x.indice_dict[key] = nbr
        else:
            nbr = x.indice_dict[key]
        return x.replace_feature(ON.subm_conv(x.features, self.weight, nbr))


class SparseConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        assert self.kernel_size[0] == 2 and self.stride[0] == 2 and self.padding[0] == 0
        rb = ON.down_rulebook(x._np_indices(), x._shape())
        x.indice_dict[("down", self.indice_key)] = (rb, x.indices, x.spatial_shape)
        feats = ON.down_conv(x.features, self.weight, rb)
        return SparseConvTensor(feats, torch.from_numpy(rb.coarse_indices), rb.coarse_shape, x.batch_size, x.indice_dict)


class SparseInverseConv3d(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, indice_key=indice_key)

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        rb, fine_indices, fine_shape = x.indice_dict[("down", self.indice_key)]
        feats = ON.inverse_conv(x.features, self.weight, rb)
        return SparseConvTensor(feats, fine_indices, fine_shape, x.batch_size, x.indice_dict)


class PointToVoxel:
    """spconv.pytorch.utils.PointToVoxel, CPU semantics, max_num_points_per_voxel=1."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel,
                 device=None):
        assert max_num_points_per_voxel == 1
        self.vsize = [float(v) for v in vsize_xyz]
        self.range = np.array([float(v) for v in coors_range_xyz], np.float32)
        self.max_num_voxels = max_num_voxels

    def generate_voxel_with_id(self, pc: torch.Tensor):
        p = pc.detach().cpu().numpy().astype(np.float32)
        s = OT.CubeSample(np.zeros(3, np.int64), np.arange(len(p)), self.range[:3], self.range[3:], self.range[:3],
                          self.range[3:], np.ones(len(p), bool))
        v = OT.voxelise_sample(p[:, :3], s, self.vsize[0])
        voxels = torch.from_numpy(p[v.winner_local]).unsqueeze(1)             # [n_v, 1, F]
        coords = torch.from_numpy(v.coords_zyx.copy())                         # int32 [n_v, 3] (z, y, x)
        num = torch.ones(len(v.winner_local), dtype=torch.int32)
        pid = torch.full((len(p),), -1, dtype=torch.int64)
        return voxels, coords, num, pid


# ------------------------------------------------------------------------------------ open3d

class _Vec(np.ndarray):
    pass


class _PointCloud:
    def __init__(self, points=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)

    def get_min_bound(self):
        return np.asarray(self.points, np.float64).min(axis=0)

    def get_max_bound(self):
        return np.asarray(self.points, np.float64).max(axis=0)

    def voxel_down_sample_and_trace(self, voxel_size, min_bound, max_bound, approximate_class=False):
        pts = np.asarray(self.points, np.float64)
        mean, first = OC.voxel_trace(pts, voxel_size, np.asarray(min_bound, np.float64))
        # per-voxel original indices in ascending order (Open3D returns them like that)
        ref = (pts - np.asarray(min_bound, np.float64)[None, :]) / float(voxel_size)
        ijk = np.floor(ref).astype(np.int64)
        ijk -= ijk.min(axis=0)
        dims = ijk.max(axis=0) + 1
        lin = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
        order = np.argsort(lin, kind="stable")
        _, starts = np.unique(lin[order], return_index=True)
        groups = np.split(order, starts[1:])
        groups.sort(key=lambda g: g[0])
        return _PointCloud(mean), None, [list(g) for g in groups]

    def compute_point_cloud_distance(self, other):
        from scipy.spatial import cKDTree
        return cKDTree(np.asarray(other.points, np.float64)).query(np.asarray(self.points, np.float64), k=1)[0]

    def paint_uniform_color(self, c):
        return self


def _make_open3d() -> types.ModuleType:
    o3d = types.ModuleType("open3d")
    o3d.geometry = types.SimpleNamespace(PointCloud=_PointCloud, TriangleMesh=object, LineSet=object)
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, dtype=np.float64),
                                        Vector2iVector=lambda a: np.asarray(a), IntVector=list)
    return o3d


# ------------------------------------------------------------------------------------ install

_installed = False


def install() -> None:
    """Register the stand-ins and make the reference checkout importable."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("neither /root/reference nor baseline/_ref/SmartQSM is present: the reference shim needs the "
                           "reference's own Python (run __graft_entry__.build() in the authoring container)")
    sp = types.ModuleType("spconv")
    spp = types.ModuleType("spconv.pytorch")
    spu = types.ModuleType("spconv.pytorch.utils")
    spm = types.ModuleType("spconv.pytorch.modules")
    for name, obj in (("SparseConvTensor", SparseConvTensor), ("SparseSequential", SparseSequential),
                      ("SubMConv3d", SubMConv3d), ("SparseConv3d", SparseConv3d),
                      ("SparseInverseConv3d", SparseInverseConv3d), ("SparseModule", SparseModule)):
        setattr(spp, name, obj)
    spu.PointToVoxel = PointToVoxel
    spm.SparseModule = SparseModule
    sp.pytorch = spp
    spp.utils = spu
    spp.modules = spm
    sys.modules.update({"spconv": sp, "spconv.pytorch": spp, "spconv.pytorch.utils": spu, "spconv.pytorch.modules": spm,
                        "open3d": _make_open3d()})
    # imported at module level by files on the way to the skeletonisation pipelines but never touched on this path:
    # cvxpy (utils/scipy_extra.py:6, used by an unrelated spline fit), distinctipy (utils/get_distinct_colors.py:1, GUI colours)
    for name in ("cvxpy", "distinctipy", "distinctipy._colorsets_data"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            m.colors = []
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def _load_reference_file(rel_path: str, name: str):
    """Import ONE reference file unmodified under a private package, so that its relative import of
    ``.core_algorithm_base`` resolves without executing ``core/core_algorithms/__init__.py`` (which drags in the
    two unrelated skeletonisers and their extra dependencies)."""
    import importlib.util
    pkg_name = "_refshim_core_algorithms"
    pkg_dir = os.path.join(REFERENCE_ROOT, os.path.dirname(rel_path))
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [pkg_dir]
        sys.modules[pkg_name] = pkg
    full = f"{pkg_name}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, rel_path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_model(state: Dict[str, np.ndarray]):
    """The reference's ``SmartTreeXX`` (models/smart_tree_xx.py) with ``state`` loaded, eval mode."""
    install()
    from external.SmartTreeXX.models.smart_tree_xx import SmartTreeXX
    m = SmartTreeXX(u_channels=[8, 16, 32, 64], mlp_hidden_layers=[8, 4], conv_block_reps=1)
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state.items()}
    for k in list(m.state_dict().keys()):
        if k.endswith("num_batches_tracked") and k not in sd:
            sd[k] = torch.zeros((), dtype=torch.int64)
    m.load_state_dict(sd)
    m.eval()
    return m


def reference_contraction(state: Dict[str, np.ndarray], **kwargs):
    """The reference's ``SpconvBasedContraction`` (core/core_algorithms/spconv_based_contraction.py) with the
    checkpoint loading replaced by ``state`` (``torch.load`` is patched to return it)."""
    install()
    mod = _load_reference_file("core/core_algorithms/spconv_based_contraction.py", "spconv_based_contraction")
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state.items()}
    ck = {"model_state_dict": sd, "hparams": {"u_channels": [8, 16, 32, 64], "mlp_hidden_layers": [8, 4], "conv_block_reps": 1}}
    real_load = torch.load
    torch.load = lambda *a, **k: ck          # spconv_based_contraction.py:62
    try:
        obj = mod.SpconvBasedContraction(ckpt_path="/nonexistent.pth", device="cpu", **kwargs)
    finally:
        torch.load = real_load
    return obj


# ------------------------------------------------------------------------------------ skeletonisation pipelines

def reference_graph_stage(skeletal_points_f32: np.ndarray, radii: Optional[np.ndarray] = None, r: float = 0.055,
                          edge_weight_function: str = "l1", topology_extractor: str = "spt"):
    """The reference's OWN ``_construct_graph`` + ``_extract_skeleton``
    (core/skeletonization_pipeline_base.py:75-238) on given skeletal points.  SciPy and networkx are the real
    libraries; only ``open3d`` (touched in verbose branches) is a stub.  Returns the pipeline object: ``_edges``,
    ``_graph``, ``_skeleton`` hold the results."""
    install()
    from core.skeletonization_pipeline_base import SkeletonizationPipelineBase
    p = SkeletonizationPipelineBase(verbose=False)
    p.set_params(edge_search_kwargs={"r": r}, edge_weight_function=edge_weight_function,
                 topology_extractor=topology_extractor)
    p._skeletal_points = np.ascontiguousarray(skeletal_points_f32, dtype=np.float32)
    p._radii = radii
    p._construct_graph()
    edges = [list(e) for e in p._edges]
    p._extract_skeleton()
    p._merged_edges = np.array(edges, dtype=np.int64).reshape(-1, 2)
    return p


def reference_thinning_pipeline(registry_patch=None, **set_params_kwargs):
    """The reference's ``ThinningBasedSkeletonizationPipeline`` (core/thinning_based_skeletonization_pipeline.py),
    imported unmodified.  ``registry_patch(registry)`` may register another plug-in class before ``set_params``
    looks the algorithm up (``:55-58``) - this is how the B200 plug-in is driven through the reference's own code."""
    install()
    import core.thinning_based_skeletonization_pipeline as T
    if registry_patch is not None:
        registry_patch(T.registry)
    p = T.ThinningBasedSkeletonizationPipeline(verbose=False)
    p.set_params(**set_params_kwargs)
    return p
