#!/usr/bin/env python
"""Row f3 timing: sqsm_ball_lists on skeleton-like clouds against SciPy's KDTree.query_ball_point (the reference's call,
core/refinement.py:448-451).  Usage (GPU): python scripts/ball_lists_bench.py  -> markdown on stdout."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.spatial import KDTree                            # noqa: E402
from smartqsm_b200._lib import Engine                      # noqa: E402
from smartqsm_b200.synthetic import make_tree              # noqa: E402


def radii_for(pts, rng):
    """Thin twigs everywhere, a thick trunk near the ground: 2 mm ... 0.3 m, times the default occupancy factor 1.5."""
    h = pts[:, 2] - pts[:, 2].min()
    base = 0.3 * np.exp(-h / 2.0) + 0.002
    return base * rng.uniform(0.8, 1.2, len(pts)) * 1.5


def main():
    rng = np.random.default_rng(0)
    eng = Engine()
    print("| nodes | list entries | first call ms (pool growth) | device ms (incl. host copies) | SciPy s | ratio | equal |")
    print("|---|---|---|---|---|---|---|")
    for n, with_scipy in ((100_000, True), (400_000, True), (1_700_000, False)):
        pts = make_tree(n, 4).astype(np.float32)
        radii = radii_for(pts, rng)
        t0 = time.perf_counter()
        eng.ball_lists(pts, radii)                          # first call at this size: the stream-ordered pool grows
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        off, idx = eng.ball_lists(pts, radii)
        t_dev = time.perf_counter() - t0
        t_sp, eq = float("nan"), "-"
        if with_scipy:
            t0 = time.perf_counter()
            lists = KDTree(pts).query_ball_point(pts, r=radii)
            t_sp = time.perf_counter() - t0
            eq = all(np.array_equal(np.sort(np.asarray(l)), idx[off[i]:off[i + 1]]) for i, l in enumerate(lists))
        print(f"| {n} | {off[-1]} | {t_first * 1e3:.1f} | {t_dev * 1e3:.1f} | {t_sp:.2f} | {t_sp / t_dev:.0f} | {eq} |", flush=True)


if __name__ == "__main__":
    main()
