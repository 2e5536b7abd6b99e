#!/usr/bin/env python
"""How much of the SubM rulebook could a tile-level skip remove?  ("first measure ...")

Runs one contraction thunk with the debug capture on, fetches the dense [27][n] rulebook of every level in the engine's
own row order (bricks in sorted key order, (z, y, x) inside a brick - the order the tensor-core kernel tiles) and reports,
per level:

* occupancy: present (row, offset) pairs / 27 n
* the share of (128-row tile x offset) slots with no present row        -> what skipping a whole MMA K-slice could save
* the share of (128-row tile x 128-byte K chunk) slots that are empty     -> what skipping a ring stage could save
* the share of (4-row x chunk-row) LDGSTS instructions whose lanes are all zero-fill -> what the producer could skip
* the same tile x offset share after sorting the rows of every 1024-row window by their 27-bit mask

Usage (GPU): python scripts/rulebook_occupancy.py [points] [leafon|leafoff]  -> markdown table on stdout
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartqsm_b200._lib import Engine                      # noqa: E402
from smartqsm_b200.checkpoints import load_state           # noqa: E402
from smartqsm_b200.synthetic import make_tree              # noqa: E402


def stats(nbr: np.ndarray, ci: int):
    kv, n = nbr.shape
    pres = nbr >= 0                                         # [27][n]
    nt = n // 128
    p = pres[:, :nt * 128].reshape(kv, nt, 128)
    occ = pres.mean()
    tile_off_empty = 1.0 - p.any(axis=2).mean()
    opc = max(1, 64 // ci)                                  # kernel offsets per 128-byte chunk of FP16 operands
    nch = (kv + opc - 1) // opc
    pad = np.zeros((nch * opc, nt, 128), bool)
    pad[:kv] = p
    ch = pad.reshape(nch, opc, nt, 128)
    tile_chunk_empty = 1.0 - ch.any(axis=(1, 3)).mean()
    instr = ch.reshape(nch, opc, nt, 32, 4)                 # one LDGSTS warp instruction = 4 tile rows x one chunk row
    instr_empty = 1.0 - instr.any(axis=(1, 4)).mean()
    # rows sorted by mask inside windows of 1024 rows
    mask = np.zeros(n, np.uint32)
    for k in range(kv):
        mask |= pres[k].astype(np.uint32) << k
    nw = n // 1024
    order = np.argsort(mask[:nw * 1024].reshape(nw, 1024), axis=1, kind="stable") + (np.arange(nw) * 1024)[:, None]
    ps = pres[:, order.reshape(-1)].reshape(kv, nw * 8, 128)
    sorted_empty = 1.0 - ps.any(axis=2).mean()
    return occ, tile_off_empty, tile_chunk_empty, instr_empty, sorted_empty


def main():
    n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    kind = sys.argv[2] if len(sys.argv) > 2 else "leafon"
    eng = Engine()
    eng.load_state_dict(load_state(kind))
    eng.debug_capture(True)
    eng.set_points(make_tree(n_pts, 3, kind == "leafon"))
    st = eng.contract_step()
    print(f"`scripts/rulebook_occupancy.py {n_pts} {kind}`: {st['n_vox']} level-0 rows, first thunk\n")
    print("| level | C_in | rows | occupancy | tile x offset empty | tile x chunk empty | LDGSTS instr. all zero-fill | tile x offset empty after mask sort (1024-row windows) |")
    print("|---|---|---|---|---|---|---|---|")
    for lvl, ci in ((0, 8), (1, 16), (2, 32), (3, 64)):
        n = eng.debug_level_size(lvl)
        if n < 1024:
            continue
        _, nbr, _, _ = eng.debug_level_tables(lvl, False)
        o, a, b, c, d = stats(nbr, ci)
        print(f"| {lvl} | {ci} | {n} | {o:.3f} | {a:.3f} | {b:.3f} | {c:.3f} | {d:.3f} |")


if __name__ == "__main__":
    main()
