"""Oracle-side weight access (test infrastructure).

The reference's only fixtures are its two checkpoints
(``external/SmartTreeXX/ckpts/for_leafoff_seasons.pth`` / ``for_leafon_seasons.pth``, loaded at
``core/core_algorithms/spconv_based_contraction.py:62-69``).  They are AGPL artefacts and are
NOT committed here; ``export_reference_checkpoints`` converts them (when ``/root/reference``
is present, i.e. in the authoring container) into neutral ``.npz`` files under
``baseline/_ref/`` which is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

from .network import random_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CKPT_DIR = "/root/reference/external/SmartTreeXX/ckpts"
BLOB_DIR = os.path.join(ROOT, "baseline", "_ref")
NAMES = {"leafoff": "for_leafoff_seasons", "leafon": "for_leafon_seasons"}


def export_reference_checkpoints(verbose: bool = False) -> None:
    """``.pth`` -> ``baseline/_ref/<name>.npz`` (only where the reference checkout exists)."""
    if not os.path.isdir(REF_CKPT_DIR):
        return
    import torch
    os.makedirs(BLOB_DIR, exist_ok=True)
    for short, stem in NAMES.items():
        src = os.path.join(REF_CKPT_DIR, stem + ".pth")
        dst = os.path.join(BLOB_DIR, stem + ".npz")
        if not os.path.isfile(src) or (os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src)):
            continue
        ck = torch.load(src, map_location="cpu", weights_only=False)
        hp = ck["hparams"]
        assert list(hp["u_channels"]) == [8, 16, 32, 64] and list(hp["mlp_hidden_layers"]) == [8, 4] \
            and hp.get("conv_block_reps", 1) == 1, hp
        np.savez(dst, **{k: v.numpy() for k, v in ck["model_state_dict"].items()})
        if verbose:
            print(f"exported {src} -> {dst}")


REF_ROOT = "/root/reference"
REF_INSTALL = os.path.join(BLOB_DIR, "SmartQSM")


def install_reference_python(verbose: bool = False) -> None:
    """The reference "install" of this tier: its pure-Python packages on the path (``core/``, ``utils/``,
    ``external/SmartTreeXX/{datasets,models}``) copied UNMODIFIED under ``baseline/_ref/SmartQSM`` - git-ignored like the
    checkpoints, so it never enters the history, but it travels to the GPU box, where ``/root/reference`` does not exist.
    Used by ``oracle/refshim`` only (the boundary test drives the reference's own ``ThinningBasedSkeletonizationPipeline``
    with the B200 plug-in registered, on a GPU).  Only where the reference checkout exists (this container)."""
    import shutil
    if not os.path.isdir(os.path.join(REF_ROOT, "core")):
        return
    for rel in ("core", "utils", os.path.join("external", "SmartTreeXX", "datasets"), os.path.join("external", "SmartTreeXX", "models")):
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(REF_INSTALL, rel)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth"))
    for rel in (os.path.join("external", "__init__.py"), os.path.join("external", "SmartTreeXX", "__init__.py")):
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(REF_INSTALL, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
        elif not os.path.isfile(dst):
            open(dst, "w").close()
    if verbose:
        print(f"installed the reference's Python packages under {REF_INSTALL}")


def blob_path(which: str) -> Optional[str]:
    p = os.path.join(BLOB_DIR, NAMES[which] + ".npz")
    return p if os.path.isfile(p) else None


def load_state(which: str = "leafoff", allow_random: bool = True) -> Dict[str, np.ndarray]:
    """Real checkpoint when its blob is present, else seeded random-init weights of the same
    architecture (``data`` fields of bench/test output say which)."""
    p = blob_path(which)
    if p is not None:
        with np.load(p) as z:
            return {k: z[k] for k in z.files}
    if not allow_random:
        raise FileNotFoundError(f"no weight blob for {which!r} under {BLOB_DIR}")
    return random_state(seed=0 if which == "leafoff" else 1)


def weights_kind(which: str = "leafoff") -> str:
    return "reference-checkpoint" if blob_path(which) else "random-init"
