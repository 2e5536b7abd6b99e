"""Oracle (test infrastructure): rulebooks and the SmartTreeXX sparse U-Net forward.

Restates ``external/SmartTreeXX/models/smart_tree_xx.py:9-78`` and
``external/SmartTreeXX/models/sparse_module.py:10-153`` on top of a CPU restatement of
spconv's SubMConv3d / SparseConv3d(k2,s2) / SparseInverseConv3d (un-vendored; semantics in
SURVEY.md App. A/C, parity contract App. D).  See ``oracle/__init__.py``.

Conventions (App. A): weights ``[C_out, kz, ky, kx, C_in]``; voxel index ``(z, y, x)``;
SubM ``out[p] = sum_k W[k] . in[p + k - 1]`` (cross-correlation, centre offset first as a
plain matmul); down ``out[o] = sum_{k in {0,1}^3} W[k] . in[2o + k]``; inverse
``out[i] = W[k] . in_c[o]`` for the unique ``(o, k)`` with ``i = 2o + k``.
"Clean bounds" rule (App. D.3): a non-centre SubM pair is accepted only if its OUTPUT
location satisfies ``0 <= c < spatial_shape`` on every axis; there is no hash-key aliasing.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-4                       # smart_tree_xx.py:19
DOWN_KERNEL_FLIP = False            # App. A: k2s2 / inverse kernel-index convention, switchable in ONE place


# --------------------------------------------------------------------------- rulebooks

_PAD = 2                            # coordinate padding so that -1 / +1 neighbours never alias
_SPAN = 1024                        # > 480 + 2*_PAD


def _lin(b: np.ndarray, zyx: np.ndarray) -> np.ndarray:
    z = zyx[:, 0].astype(np.int64) + _PAD
    y = zyx[:, 1].astype(np.int64) + _PAD
    x = zyx[:, 2].astype(np.int64) + _PAD
    return ((b.astype(np.int64) * _SPAN + z) * _SPAN + y) * _SPAN + x


class _Lookup:
    """Exact coordinate -> row map (the role of spconv's indice hash table, without aliasing)."""

    def __init__(self, indices: np.ndarray) -> None:
        key = _lin(indices[:, 0], indices[:, 1:4])
        self.order = np.argsort(key, kind="stable")
        self.sorted = key[self.order]

    def find(self, b: np.ndarray, zyx: np.ndarray) -> np.ndarray:
        if len(self.sorted) == 0:
            return np.full(len(b), -1, np.int64)
        key = _lin(b, zyx)
        pos = np.searchsorted(self.sorted, key)
        pos = np.minimum(pos, len(self.sorted) - 1)
        hit = self.sorted[pos] == key
        return np.where(hit, self.order[pos], -1)


def subm_rulebook(indices: np.ndarray, spatial_shape) -> np.ndarray:
    """SubM 3^3 neighbour table ``nbr[27, N]`` (App. C "Rulebook, SubM"; K5 in SURVEY §2.2).

    ``nbr[k, o]`` = input row at ``coords[o] + (kz-1, ky-1, kx-1)`` (k = kz*9 + ky*3 + kx), or -1.
    Non-centre entries exist only for output rows inside ``[0, spatial_shape)``.
    """
    n = len(indices)
    nbr = np.full((27, n), -1, np.int64)
    look = _Lookup(indices)
    b = indices[:, 0]
    zyx = indices[:, 1:4].astype(np.int64)
    shape = np.asarray(spatial_shape, dtype=np.int64)
    out_ok = np.all((zyx >= 0) & (zyx < shape[None, :]), axis=1)
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                k = kz * 9 + ky * 3 + kx
                if k == 13:
                    nbr[k] = np.arange(n)
                    continue
                q = zyx + np.array([kz - 1, ky - 1, kx - 1])[None, :]
                r = look.find(b, q)
                nbr[k] = np.where(out_ok, r, -1)
    return nbr


@dataclass
class DownRulebook:
    """k=2, s=2 rulebook (K6): ``parent[i]`` coarse row (or -1), ``koff[i]`` = kz*4+ky*2+kx."""
    parent: np.ndarray
    koff: np.ndarray
    coarse_indices: np.ndarray      # int32[N1,4]
    coarse_shape: List[int]


def out_shape_k2s2(shape) -> List[int]:
    """``(s + 2*0 - 1*(2-1) - 1)//2 + 1`` with Python floor division (sparse_module.py:91-95)."""
    return [int((int(s) - 2) // 2 + 1) for s in shape]


def down_rulebook(indices: np.ndarray, spatial_shape) -> DownRulebook:
    """``SparseConv3d(k=2, s=2)`` indices on CPU: ``out = in // 2`` accepted iff
    ``out < (shape-2)//2 + 1``; coarse voxels ordered by FIRST APPEARANCE (App. C, App. D.2)."""
    oshape = out_shape_k2s2(spatial_shape)
    zyx = indices[:, 1:4].astype(np.int64)
    o = zyx // 2
    ok = np.all(o < np.asarray(oshape)[None, :], axis=1) & np.all(o >= 0, axis=1)
    koff = (zyx[:, 0] & 1) * 4 + (zyx[:, 1] & 1) * 2 + (zyx[:, 2] & 1)
    key = _lin(indices[:, 0], o)
    key_ok = key[ok]
    rows_ok = np.nonzero(ok)[0]
    uniq, first, inv = np.unique(key_ok, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                 # first-appearance order of coarse voxels
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    parent = np.full(len(indices), -1, np.int64)
    parent[rows_ok] = rank[inv]
    src = rows_ok[first[order]]
    coarse = np.concatenate([indices[src, 0:1], (zyx[src] // 2).astype(np.int32)], axis=1).astype(np.int32)
    return DownRulebook(parent, koff, coarse, oshape)


def safe_to_downsample(indices: np.ndarray, spatial_shape) -> bool:
    """``UBlock.forward`` descent guard (sparse_module.py:105-109 with :74-99)."""
    shape = [int(s) for s in spatial_shape]
    if not all(d > 2 for d in shape):                        # :107
        return False
    if not all(s >= 2 for s in shape):                       # :86-88 (min_required = 2)
        return False
    oshape = out_shape_k2s2(shape)
    o = indices[:, 1:4].astype(np.int64) // 2
    ok = np.all((o >= 0) & (o < np.asarray(oshape)[None, :]), axis=1)
    return bool(ok.any())                                    # :96-97


# --------------------------------------------------------------------------- weights

class Weights:
    """State dict of ``SmartTreeXX`` (App. A) as torch CPU tensors of one dtype."""

    def __init__(self, state: Dict[str, np.ndarray], dtype=torch.float32) -> None:
        self.t = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in state.items()
                  if not k.endswith("num_batches_tracked")}
        self.dtype = dtype

    def __getitem__(self, k: str) -> torch.Tensor:
        return self.t[k]


BN_RECORDER: Optional[Dict[str, list]] = None   # set to {} to collect (sum, sum of squares, count) of every BatchNorm INPUT


def _bn_relu(x: torch.Tensor, w: Weights, prefix: str, relu: bool = True) -> torch.Tensor:
    """``nn.BatchNorm1d(eps=1e-4)`` in eval mode + ``nn.ReLU`` (App. A)."""
    if BN_RECORDER is not None:
        xd = x.detach().double()
        acc = BN_RECORDER.setdefault(prefix, [torch.zeros(x.shape[1], dtype=torch.float64), torch.zeros(x.shape[1], dtype=torch.float64), 0])
        acc[0] += xd.sum(0)
        acc[1] += (xd * xd).sum(0)
        acc[2] += x.shape[0]
    y = F.batch_norm(x, w[prefix + ".running_mean"], w[prefix + ".running_var"],
                     w[prefix + ".weight"], w[prefix + ".bias"], False, 0.1, BN_EPS)
    return F.relu(y) if relu else y


def subm_conv(x: torch.Tensor, weight: torch.Tensor, nbr: np.ndarray) -> torch.Tensor:
    """``SubMConv3d`` forward, CPU native algorithm: centre matmul, then per offset
    gather -> mm -> scatter-add (K7)."""
    co = weight.shape[0]
    wk = weight.reshape(co, -1, weight.shape[-1])            # [C_out, K, C_in]
    if wk.shape[1] == 1:                                     # kernel_size=1 (i_branch)
        return x @ wk[:, 0, :].T
    out = x @ wk[:, 13, :].T
    for k in range(27):
        if k == 13:
            continue
        src = nbr[k]
        rows = np.nonzero(src >= 0)[0]
        if len(rows) == 0:
            continue
        out.index_add_(0, torch.from_numpy(rows), x[torch.from_numpy(src[rows])] @ wk[:, k, :].T)
    return out


def _kidx(koff: int) -> int:
    return 7 - koff if DOWN_KERNEL_FLIP else koff


def down_conv(x: torch.Tensor, weight: torch.Tensor, rb: DownRulebook) -> torch.Tensor:
    """``SparseConv3d(k=2, s=2)`` forward: ``out[o] = sum_k W[k] . in[2o+k]``."""
    co = weight.shape[0]
    wk = weight.reshape(co, 8, weight.shape[-1])
    out = torch.zeros((len(rb.coarse_indices), co), dtype=x.dtype)
    for k in range(8):
        rows = np.nonzero((rb.koff == k) & (rb.parent >= 0))[0]
        if len(rows) == 0:
            continue
        out.index_add_(0, torch.from_numpy(rb.parent[rows]), x[torch.from_numpy(rows)] @ wk[:, _kidx(k), :].T)
    return out


def inverse_conv(xc: torch.Tensor, weight: torch.Tensor, rb: DownRulebook) -> torch.Tensor:
    """``SparseInverseConv3d(k=2)`` forward on the saved pairs, roles swapped, same ``k``;
    fine voxels that did not survive the down-sample receive zeros (App. A last paragraph)."""
    co = weight.shape[0]
    wk = weight.reshape(co, 8, weight.shape[-1])
    out = torch.zeros((len(rb.parent), co), dtype=xc.dtype)
    for k in range(8):
        rows = np.nonzero((rb.koff == k) & (rb.parent >= 0))[0]
        if len(rows) == 0:
            continue
        out[torch.from_numpy(rows)] = xc[torch.from_numpy(rb.parent[rows])] @ wk[:, _kidx(k), :].T
    return out


# --------------------------------------------------------------------------- network

@dataclass
class LevelTrace:
    indices: np.ndarray
    spatial_shape: List[int]
    nbr: np.ndarray
    down: Optional[DownRulebook] = None
    descended: bool = False


@dataclass
class ForwardTrace:
    """Integer side products of one forward, kept for rulebook parity tests."""
    levels: List[LevelTrace] = field(default_factory=list)
    feats: Dict[str, torch.Tensor] = field(default_factory=dict)
    # output of every sparse convolution keyed by the reference module name ("unet.u.conv.2", ...): what a forward hook on
    # the genuine spconv layers records (scripts/reference_dump.py); filled only when keep=True
    taps: Optional[Dict[str, torch.Tensor]] = None
    tap_indices: Optional[Dict[str, np.ndarray]] = None


def _tap(trace: Optional["ForwardTrace"], name: str, feats: torch.Tensor, indices: np.ndarray) -> None:
    if trace is not None and trace.taps is not None:
        trace.taps[name] = feats.clone()
        trace.tap_indices[name] = indices


def _residual_block(x, w: Weights, prefix: str, nbr, c_in: int, c_out: int, trace=None, indices=None) -> torch.Tensor:
    """``ResidualBlock.forward`` (sparse_module.py:32-39; ctor :11-30)."""
    if c_in == c_out:
        ident = x
    else:
        ident = subm_conv(x, w[prefix + ".i_branch.0.weight"], nbr)
        _tap(trace, prefix + ".i_branch.0", ident, indices)
    y = _bn_relu(x, w, prefix + ".conv_branch.0")
    y = subm_conv(y, w[prefix + ".conv_branch.2.weight"], nbr)
    _tap(trace, prefix + ".conv_branch.2", y, indices)
    y = _bn_relu(y, w, prefix + ".conv_branch.3")
    y = subm_conv(y, w[prefix + ".conv_branch.5.weight"], nbr)
    _tap(trace, prefix + ".conv_branch.5", y, indices)
    return y + ident


def _ublock(x, indices, spatial_shape, w: Weights, prefix: str, planes: List[int], trace: ForwardTrace,
            keep: bool, nbr: Optional[np.ndarray] = None) -> torch.Tensor:
    """``UBlock.forward`` (sparse_module.py:101-119), recursive."""
    if nbr is None:
        nbr = subm_rulebook(indices, spatial_shape)
    lt = LevelTrace(indices, [int(s) for s in spatial_shape], nbr)
    trace.levels.append(lt)
    lvl = len(trace.levels) - 1
    x = _residual_block(x, w, prefix + ".blocks.block0", nbr, planes[0], planes[0], trace, indices)
    if keep:
        trace.feats[f"enc{lvl}"] = x.clone()
    if len(planes) > 1 and safe_to_downsample(indices, spatial_shape):
        rb = down_rulebook(indices, spatial_shape)
        lt.down, lt.descended = rb, True
        y = _bn_relu(x, w, prefix + ".conv.0")
        y = down_conv(y, w[prefix + ".conv.2.weight"], rb)
        _tap(trace, prefix + ".conv.2", y, rb.coarse_indices)
        y = _ublock(y, rb.coarse_indices, rb.coarse_shape, w, prefix + ".u", planes[1:], trace, keep)
        y = _bn_relu(y, w, prefix + ".deconv.0")
        y = inverse_conv(y, w[prefix + ".deconv.2.weight"], rb)
        _tap(trace, prefix + ".deconv.2", y, indices)
        x = torch.cat((x, y), dim=1)                          # :114
        x = _residual_block(x, w, prefix + ".blocks_tail.block0", nbr, 2 * planes[0], planes[0], trace, indices)
        if keep:
            trace.feats[f"dec{lvl}"] = x.clone()
    return x


def _mlp(x, w: Weights, prefix: str, n_hidden: int) -> torch.Tensor:
    """``MLP`` (sparse_module.py:121-153): (Linear+BN+ReLU) x n_hidden, Linear."""
    for i in range(n_hidden):
        x = F.linear(x, w[f"{prefix}.sequence.{3 * i}.weight"], w[f"{prefix}.sequence.{3 * i}.bias"])
        x = _bn_relu(x, w, f"{prefix}.sequence.{3 * i + 1}")
    j = 3 * n_hidden
    return F.linear(x, w[f"{prefix}.sequence.{j}.weight"], w[f"{prefix}.sequence.{j}.bias"])


def forward(feats: np.ndarray, indices: np.ndarray, spatial_shape, w: Weights,
            planes=(8, 16, 32, 64), n_hidden: int = 2, keep: bool = False):
    """``SmartTreeXX.forward`` (smart_tree_xx.py:64-78) for ONE collated batch.

    Returns ``(direction [N,3], distance [N,1], ForwardTrace)``.
    """
    trace = ForwardTrace()
    if keep:
        trace.taps, trace.tap_indices = {}, {}
    x = torch.as_tensor(feats).to(w.dtype)
    nbr0 = subm_rulebook(indices, spatial_shape)
    x = subm_conv(x, w["input_conv.0.weight"], nbr0)                               # :65
    _tap(trace, "input_conv.0", x, indices)
    if keep:
        trace.feats["input_conv"] = x.clone()
    x = _ublock(x, indices, spatial_shape, w, "unet", list(planes), trace, keep, nbr0)  # :66 (same indice_key "subm1")
    x = _bn_relu(x, w, "output_layer.0")                                           # :67
    offset = _mlp(x, w, "offset_linear", n_hidden)                                 # :69
    regression = _mlp(x, w, "regression_linear", n_hidden)                         # :74
    direction = F.normalize(offset)                                                # :75
    distance = F.softplus(regression, beta=1.0, threshold=20.0)                    # :76
    if keep:
        trace.feats["offset"] = offset
        trace.feats["regression"] = regression
    return direction, distance, trace


# --------------------------------------------------------------------------- weights I/O

def layer_shapes(planes=(8, 16, 32, 64), hidden=(8, 4)) -> Dict[str, tuple]:
    """Every tensor of the SmartTreeXX state dict with its shape (App. A; 161 tensors with
    ``num_batches_tracked``)."""
    s: Dict[str, tuple] = {"input_conv.0.weight": (planes[0], 3, 3, 3, 3)}

    def bn(p, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[f"{p}.{n}"] = (c,)

    def block(p, ci, co):
        if ci != co:
            s[p + ".i_branch.0.weight"] = (co, 1, 1, 1, ci)
        bn(p + ".conv_branch.0", ci)
        s[p + ".conv_branch.2.weight"] = (co, 3, 3, 3, ci)
        bn(p + ".conv_branch.3", co)
        s[p + ".conv_branch.5.weight"] = (co, 3, 3, 3, co)

    def ublock(p, pl):
        block(p + ".blocks.block0", pl[0], pl[0])
        if len(pl) > 1:
            bn(p + ".conv.0", pl[0])
            s[p + ".conv.2.weight"] = (pl[1], 2, 2, 2, pl[0])
            ublock(p + ".u", pl[1:])
            bn(p + ".deconv.0", pl[1])
            s[p + ".deconv.2.weight"] = (pl[0], 2, 2, 2, pl[1])
            block(p + ".blocks_tail.block0", 2 * pl[0], pl[0])

    ublock("unet", list(planes))
    bn("output_layer.0", planes[0])
    for head, last in (("offset_linear", 3), ("regression_linear", 1)):
        dims = [planes[0]] + list(hidden) + [last]
        for i in range(len(dims) - 1):
            s[f"{head}.sequence.{3 * i}.weight"] = (dims[i + 1], dims[i])
            s[f"{head}.sequence.{3 * i}.bias"] = (dims[i + 1],)
            if i < len(dims) - 2:
                bn(f"{head}.sequence.{3 * i + 1}", dims[i + 1])
    return s


def random_state(seed: int = 0, planes=(8, 16, 32, 64), hidden=(8, 4)) -> Dict[str, np.ndarray]:
    """Seeded random-init weights of the SmartTreeXX architecture (used when the shipped
    checkpoints are not available, e.g. on the GPU box).  Scales are chosen so activations
    stay O(1) and the predicted step is a few centimetres."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shp in layer_shapes(planes, hidden).items():
        if name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shp)
        elif name.endswith("running_mean"):
            v = rng.normal(0.0, 0.2, shp)
        elif name.endswith(".weight") and len(shp) == 1:
            v = rng.uniform(0.8, 1.2, shp)
        elif name.endswith(".bias") and len(shp) == 1 and ".sequence." not in name:
            v = rng.normal(0.0, 0.1, shp)
        elif name.endswith(".bias"):
            v = rng.normal(0.0, 0.1, shp)
        else:
            fan_in = int(np.prod(shp[1:]))
            active = fan_in / 3.0 if len(shp) == 5 and shp[1] == 3 else fan_in
            v = rng.normal(0.0, 1.0 / np.sqrt(max(active, 1.0)), shp)
        out[name] = v.astype(np.float32)
    # the first conv sees absolute coordinates of up to tens of metres: keep it tame
    out["input_conv.0.weight"] *= 0.1
    # last regression bias: softplus(-3) ~ 0.05 m steps
    out["regression_linear.sequence.6.bias"] = np.array([-3.0], np.float32)
    return out
