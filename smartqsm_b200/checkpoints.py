"""Weights for the measurement and demo scripts: the reference's checkpoints when their neutral ``.npz`` export is present
under ``baseline/_ref/`` (written by ``__graft_entry__.build()`` where a SmartQSM checkout exists; AGPL artefacts, never
committed), else seeded random-init weights of the same architecture.  Product-side helper: nothing here touches
``oracle/``.  For a real run the plug-in loads the ``.pth`` named by the config (``core_algorithms.load_checkpoint``).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

__all__ = ["layer_shapes", "random_state", "blob_path", "load_state", "weights_kind"]

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOB_DIR = os.path.join(_ROOT, "baseline", "_ref")
NAMES = {"leafoff": "for_leafoff_seasons", "leafon": "for_leafon_seasons"}


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


def blob_path(which: str) -> Optional[str]:
    p = os.path.join(BLOB_DIR, NAMES[which] + ".npz")
    return p if os.path.isfile(p) else None


def load_state(which: str = "leafoff", allow_random: bool = True) -> Dict[str, np.ndarray]:
    """Real checkpoint when its blob is present, else seeded random-init weights of the same architecture."""
    p = blob_path(which)
    if p is not None:
        with np.load(p) as z:
            return {k: z[k] for k in z.files}
    if not allow_random:
        raise FileNotFoundError(f"no weight blob for {which!r} under {BLOB_DIR}")
    return random_state(seed=0 if which == "leafoff" else 1)


def weights_kind(which: str = "leafoff") -> str:
    return "reference-checkpoint" if blob_path(which) else "random-init"
