"""Shared helpers of the parity tests: turn the oracle's per-batch traces into global tables."""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def oracle_level_tables(tr, batch_size: int) -> List[dict]:
    """Concatenate the per-batch traces of ``oracle.contraction.contract_once(keep=True)``.

    Returns one dict per U-Net level with ``coords`` int64[n,4] = (global sample, z, y, x), ``nbr``
    int64[27,n] (global row numbers of that level), ``parent`` / ``koff`` (or None) and the per-batch
    ``descended`` flags; rows are in the oracle's canonical order (batch, cube, first appearance).
    """
    levels: List[dict] = []
    for lvl in range(4):
        coords, nbr, parent, koff, desc = [], [], [], [], []
        base = 0
        next_base = 0
        any_rows = False
        for bi, ft in enumerate(tr.fwd):
            if lvl >= len(ft.levels):
                desc.append(False)
                continue
            lt = ft.levels[lvl]
            any_rows = True
            n = len(lt.indices)
            c = lt.indices.astype(np.int64).copy()
            c[:, 0] += bi * batch_size
            coords.append(c)
            nb = lt.nbr.copy()
            nb[nb >= 0] += base
            nbr.append(nb)
            desc.append(bool(lt.descended))
            if lt.descended:
                p = lt.down.parent.copy()
                p[p >= 0] += next_base
                parent.append(p)
                koff.append(lt.down.koff)
                next_base += len(lt.down.coarse_indices)
            else:
                parent.append(np.full(n, -1, np.int64))
                koff.append(((c[:, 1] & 1) * 4 + (c[:, 2] & 1) * 2 + (c[:, 3] & 1)))
            base += n
        if not any_rows:
            break
        levels.append({
            "coords": np.concatenate(coords), "nbr": np.concatenate(nbr, axis=1),
            "parent": np.concatenate(parent), "koff": np.concatenate(koff), "descended": desc,
        })
    return levels


def oracle_features(tr, name: str) -> np.ndarray:
    out = [ft.feats[name].numpy() for ft in tr.fwd if name in ft.feats]
    return np.concatenate(out, axis=0) if out else np.zeros((0, 0), np.float32)


def coord_key(c: np.ndarray) -> np.ndarray:
    c = c.astype(np.int64)
    return ((c[:, 0] * 1024 + c[:, 1]) * 1024 + c[:, 2]) * 1024 + c[:, 3]


def match_rows(gpu_coords: np.ndarray, oracle_coords: np.ndarray) -> np.ndarray:
    """g2o[r_gpu] = oracle row with the same (sample, z, y, x); asserts both voxel sets are equal."""
    gk, ok = coord_key(gpu_coords), coord_key(oracle_coords)
    assert len(gk) == len(ok), f"row count differs: gpu {len(gk)} oracle {len(ok)}"
    go, oo = np.argsort(gk, kind="stable"), np.argsort(ok, kind="stable")
    assert np.array_equal(gk[go], ok[oo]), "voxel sets differ"
    g2o = np.empty(len(gk), np.int64)
    g2o[go] = oo
    return g2o


def remap(table: np.ndarray, g2o: np.ndarray) -> np.ndarray:
    """Translate a table of GPU row numbers (-1 = none) into oracle row numbers."""
    out = np.full(table.shape, -1, np.int64)
    m = table >= 0
    out[m] = g2o[table[m]]
    return out


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| relative to the tensor scale (max |b|, at least 1e-30)."""
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(float(np.max(np.abs(b))), 1e-30))


def make_raw_scan(n_surface: int, seed: int, leaf_on: bool = False, outliers: int = 200) -> np.ndarray:
    """A raw scan as the entry point would read it from disk (input of ``smartqsm.py:882``): several noisy
    returns per surface point, a sprinkle of far outliers, projected (UTM-like) coordinates, shuffled."""
    from smartqsm_b200.synthetic import make_tree
    pts = make_tree(n_surface, seed, leaf_on)
    rng = np.random.default_rng(seed)
    extra = pts[rng.integers(0, len(pts), len(pts) // 2)]
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    raw = np.concatenate([pts + rng.normal(0, 0.002, pts.shape), extra + rng.normal(0, 0.003, extra.shape),
                          rng.uniform(lo - 1.0, hi + 1.0, (outliers, 3))])
    raw += np.array([351234.5, 5412345.25, 102.75])
    rng.shuffle(raw)
    return np.ascontiguousarray(raw)
