#!/usr/bin/env python
"""Benchmark of the B200 ``spconv-contraction`` path (BASELINE.json metric: points/sec).

A *step* = the whole stage for the workload's input: ``max_iteration`` (10) ``_contract`` thunks in
the reference configuration (Chamfer monitor on, no early-stop latch: a rejected iteration is
recomputed exactly like ``core/core_algorithms/spconv_based_contraction.py:123-127`` does), then
skeletal sampling + radii and the radius-neighbour edge list.  points/sec = input points / step time.

Workloads (BASELINE.json ``configs``):

* ``C3`` (default, the metric's configuration) 20 M-point leaf-on tree; ``C2`` 5 M leaf-off; ``C1`` 1 M leaf-off.
  One tree per GPU (``torchrun``: seed = base + rank), no data-path collective, ``scaling: weak``.
* ``C4`` a stand plot of 64 trees of about 2 M points (seeds 100..163), partitioned over the ranks by
  longest-processing-time bin packing (``smartqsm_b200.sharding.partition_trees``), ``scaling: strong``.
* ``C5`` ONE tree (``--points``, default 200 M) sharded over the ranks by cube batches with one all-gather per thunk
  (``ShardedTreeContraction``), ``scaling: strong``.

Keys of the JSON line:

* ``value``      inputs already resident in HBM (float32 device array) when the timed region starts;
                 device time from CUDA events on the engine's stream, max over ranks.
* ``e2e``        the same step through the plug-in / C ABI with HOST buffers: float64 input copied
                 from pinned host memory, results (contracted points, displacements, skeletal points,
                 radii, edges) copied back, all inside the timed region.
* ``roofline``   the longest single kernel family (``conv_*``, ``cd_*``, ``rb_*``, ``vx_*``): algorithmic bytes
                  over its device time; ``stage_frac`` = all algorithmic bytes of the step over
                 the step time over the same peak.
* ``mode_b``     SURVEY §8(d) iteration accounting (b): monitor off, exactly 10 accepted iterations.
* ``cpu_baseline`` / ``--impl reference``  the restated reference CPU path (``oracle/``) on the box's
                 host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (points, leaf_on, checkpoint)
    "C3": (20_000_000, True, "leafon"),    # spconv-contraction-LEAFON, 20M-pt leaf-on TLS tree (the metric's config)
    "C2": (5_000_000, False, "leafoff"),   # spconv-contraction-LEAFOFF, 5M-pt leaf-off tree
    "C1": (1_000_000, False, "leafoff"),   # the reference's CPU-runnable case
    "C4": (2_000_000, False, "leafoff"),   # 64-tree stand plot, ~2M points per tree, tree-sharded
    "C5": (200_000_000, True, "leafon"),   # one giant scan, sharded by cube batches
}
C4_TREES, C4_SEED0 = 64, 100
CFG = dict(batch_size=16, max_iteration=10, voxel_size=0.01, cube_size=4.0, buffer_size=0.4)
SKELETAL_VOXEL = 0.01      # voxel_size_for_downsampling (configs/spconv-contraction-*.yaml)
EDGE_R = 0.055             # edge_search_kwargs.r
STAGE_FAMILIES = ("tiling", "voxelise", "rulebook", "subm_l0", "subm_l1", "subm_l2", "subm_l3", "down_up", "heads", "chamfer",
                  "skeletal", "radius_graph")       # these partition the step; conv_* / cd_* / rb_* / vx_* are kernels inside them
KERNEL_PREFIXES = ("conv_", "cd_", "rb_", "vx_")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SQSM_BENCH_WORKLOAD", "C3"), choices=list(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override the point count (C5: size of the sharded tree)")
    ap.add_argument("--trees", type=int, default=0, help="C4: number of trees (default 64)")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-points", type=int, default=0)
    ap.add_argument("--device-only", action="store_true", help="tuning runs: skip the e2e, latched and mode (b) legs")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers

def load_state(which: str):
    """Checkpoint blob exported by build() from the reference checkout if present, else seeded
    random-init weights of the same architecture (smartqsm_b200/checkpoints.py; no oracle code on the GPU arm)."""
    from smartqsm_b200 import checkpoints
    return checkpoints.load_state(which), checkpoints.weights_kind(which)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/sqsm_clocks_{os.getpid()}.csv"

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def run_oracle_stage(state, pts, latch=False):
    """The restated reference CPU path for one tree: 10 thunks (monitor on) + sampling + edges."""
    from oracle import contraction as OC
    o = OC.SpconvBasedContractionOracle(state=state, monitored_by_chamfer_distance=True, latch=latch, **CFG)
    o.input(points=pts)
    for f in o.get_pipeline():
        f()
    out = o.output()
    ids, sk, rad = OC.skeletal_points_and_radii(out["contracted_points"], out["displacements"], SKELETAL_VOXEL)
    OC.radius_edges(sk, EDGE_R)
    return o.history


def cached_tree(n: int, seed: int, leaf_on: bool) -> np.ndarray:
    """The seeded synthetic tree; kept under /tmp so that back-to-back bench runs on one box do not regenerate it
    (28 s for 20 M points).  Same generator, same seed, same array."""
    from smartqsm_b200.synthetic import make_tree
    path = f"/tmp/sqsm_tree_{n}_{seed}_{int(leaf_on)}.npy"
    try:
        a = np.load(path)
        if a.shape == (n, 3) and a.dtype == np.float64:
            return a
    except Exception:
        pass
    a = make_tree(n, seed, leaf_on)
    try:
        np.save(path + f".{os.getpid()}.tmp.npy", a)
        os.replace(path + f".{os.getpid()}.tmp.npy", path)
    except Exception:
        pass
    return a


def cpu_sample_points(budget_s: float) -> int:
    # the oracle runs the whole stage at roughly 2.5-3.5 k points/s on 8-16 cores
    return int(np.clip(budget_s * 2500, 4000, 120_000))


# ------------------------------------------------------------------------------------------ reference arm

def workload_text(name: str, n_pts: int, n_trees: int, leaf_on: bool) -> str:
    kind = "leaf-on" if leaf_on else "leaf-off"
    if name == "C4":
        return f"C4: spconv-contraction stage on a synthetic stand plot of {n_trees} {kind} trees (~{n_pts} points each, seeds {C4_SEED0}..), tree-sharded (LPT)"
    if name == "C5":
        return f"C5: spconv-contraction stage on ONE synthetic {n_pts}-point {kind} tree sharded over the GPUs by cube batches"
    return f"{name}: spconv-contraction stage on a synthetic {n_pts}-point {kind} tree per GPU (tree-sharded)"


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from smartqsm_b200.synthetic import make_tree
    n_full, leaf_on, which = WORKLOADS[args.workload]
    if args.points:
        n_full = args.points
    state, wkind = load_state(which)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    total = max(1, args.steps + args.warmup)
    n = args.cpu_sample_points or cpu_sample_points(150.0 / total)
    pts = make_tree(n, args.seed, leaf_on)
    for _ in range(args.warmup):
        run_oracle_stage(state, pts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hist = run_oracle_stage(state, pts)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = n / dt
    sample = (f"whole stage (10 _contract thunks, Chamfer monitor on, skeletal sampling, radius graph) on a "
              f"{n}-point synthetic {'leaf-on' if leaf_on else 'leaf-off'} tree from the same generator; the reference's own "
              f"CPU path cannot run (spconv / Open3D not installable), this is its restatement (oracle/)")
    line = {
        "impl": "reference", "metric": "points/sec", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.workload in ("C1", "C2", "C3") else "strong", "vs_baseline": None, "dtype": "f32",
        "data": f"synthetic; weights: {wkind}",
        "config": {"workload": workload_text(args.workload, n_full, args.trees or C4_TREES, leaf_on) +
                               f" (timed on a bounded {n}-point sample)", **CFG, "monitored_by_chamfer_distance": True},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "iterations": {"accepted": sum(1 for h in hist if h.get("accepted")), "executed": len(hist)},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ B200 arm

def tree_stage(eng, fetch_host: bool, host_out=None):
    """After the input was loaded: the 10 thunks, sampling and the edge list of ONE tree."""
    stats = [eng.contract_step() for _ in range(CFG["max_iteration"])]
    return stats, *tree_tail(eng, fetch_host, host_out)


def tree_tail(eng, fetch_host: bool, host_out=None):
    if fetch_host:
        # results land in pinned host buffers (host_out: dict of pinned torch tensors sized for the worst case)
        n = eng.result_into(host_out["P"].data_ptr(), host_out["D"].data_ptr())
        m = eng.skeletal_sample(SKELETAL_VOXEL, fetch=False)
        eng.skeletal_sample_into(SKELETAL_VOXEL, host_out["ids"].data_ptr(), host_out["sk"].data_ptr(), host_out["rad"].data_ptr())
        e = eng.radius_graph(EDGE_R, fetch=False)
        if e * 16 > host_out["edges"].numel() * host_out["edges"].element_size():
            raise RuntimeError("edge buffer too small")
        eng.radius_graph_into(EDGE_R, host_out["edges"].data_ptr())
        return n * 24 + m * (8 + 12 + 4) + e * 16, (m, e)
    m = eng.skeletal_sample(SKELETAL_VOXEL, fetch=False)
    e = eng.radius_graph(EDGE_R, fetch=False)
    return 0, (m, e)


def tree_tail_sharded(eng, rank: int, world: int, fetch_host: bool, host_out=None):
    """C5: every rank holds the final state; the skeletal sample is repeated everywhere (2 ms) and the radius-neighbour
    edge list is produced in node ranges, one per rank (``sqsm_radius_graph_range``: the whole list is the concatenation in
    rank order).  Rank 0 copies the state and the sample out, every rank its slice of the edges."""
    m = eng.skeletal_sample(SKELETAL_VOXEL, fetch=False)
    lo, hi = (m * rank) // world, (m * (rank + 1)) // world
    if fetch_host:
        b = 0
        if rank == 0:
            n = eng.result_into(host_out["P"].data_ptr(), host_out["D"].data_ptr())
            eng.skeletal_sample_into(SKELETAL_VOXEL, host_out["ids"].data_ptr(), host_out["sk"].data_ptr(), host_out["rad"].data_ptr())
            b = n * 24 + m * (8 + 12 + 4)
        cap = host_out["edges"].numel() * host_out["edges"].element_size() // 16
        e = eng.radius_graph_range(EDGE_R, lo, hi, host_out["edges"].data_ptr(), cap)
        return b + e * 16, (m if rank == 0 else 0, e)
    e = eng.radius_graph_range_count(EDGE_R, lo, hi)              # device-resident leg: the slice stays in HBM
    return 0, (m if rank == 0 else 0, e)


class TreeJob:
    """C1-C4: this rank's trees, one after the other on one engine (the reference's file loop, smartqsm.py:837)."""

    def __init__(self, trees64):
        import torch
        self.host = [torch.from_numpy(t).pin_memory() for t in trees64]                    # float64, pinned: e2e input
        self.dev = [torch.from_numpy(t.astype(np.float32)).cuda() for t in trees64]        # float32, resident: `value` input
        self.n = [len(t) for t in trees64]
        self.points = sum(self.n)
        self.max_points = max(self.n) if self.n else 0
        self.h2d = self.points * 24

    def step(self, eng, e2e: bool, host_out=None):
        d2h, sizes, stats = 0, [], []
        for i in range(len(self.n)):
            if e2e:
                eng.set_points_ptr_f64(self.host[i].data_ptr(), self.n[i])
            else:
                eng.set_points_device_f32(self.dev[i].data_ptr(), self.n[i])
            stats, b, sz = tree_stage(eng, e2e, host_out)
            d2h += b
            sizes.append(sz)
        return stats, d2h, sizes


class ShardedJob:
    """C5: one tree on all ranks; every rank holds the cloud, works on its batch range, one all-gather per thunk."""

    def __init__(self, tree64, rank, world, device):
        import torch
        self.host = torch.from_numpy(tree64).pin_memory()
        self.dev = torch.from_numpy(tree64.astype(np.float32)).cuda()
        self.rank, self.world, self.device = rank, world, device
        self.points = len(tree64) if rank == 0 else 0          # the job's points are counted once
        self.max_points = len(tree64)
        self.n = len(tree64)
        self.h2d = len(tree64) * 24                            # every rank loads the whole cloud
        self.exchange_bytes = 0

    def step(self, eng, e2e: bool, host_out=None):
        from smartqsm_b200.sharding import ShardedTreeContraction
        if e2e:
            eng.set_points_ptr_f64(self.host.data_ptr(), self.n)
        else:
            eng.set_points_device_f32(self.dev.data_ptr(), self.n)
        drv = ShardedTreeContraction(eng, self.rank, self.world, self.device, monitored=eng_monitored(eng))
        stats = [drv.contract() for _ in range(CFG["max_iteration"])]
        self.exchange_bytes = drv.exchange_bytes
        d2h, sz = tree_tail_sharded(eng, self.rank, self.world, e2e, host_out)
        return stats, d2h, [sz]


def result_digest(eng) -> dict:
    """SHA-256 of the final state (contracted points, displacements), the skeletal sample ids and radii, and the edge
    count: the sharded run (C5) must reproduce the single-GPU digest."""
    import hashlib
    P, D = eng.result()
    ids, sk, rad = eng.skeletal_sample(SKELETAL_VOXEL)
    e = eng.radius_graph(EDGE_R, fetch=False)
    h = hashlib.sha256()
    for a in (P, D, ids, rad):
        h.update(np.ascontiguousarray(a).tobytes())
    return {"sha256": h.hexdigest(), "n_final": int(len(P)), "skeletal_points": int(len(ids)), "edges": int(e)}


_MONITORED = {}


def eng_monitored(eng) -> bool:
    return _MONITORED.get(id(eng), True)


def c4_sizes(n_trees: int, per_tree: int):
    rng = np.random.default_rng(C4_SEED0)
    return [int(per_tree * (0.6 + 0.8 * rng.random())) for _ in range(n_trees)]           # ~2 M +- 40 %: LPT has work to do


def main_b200(args):
    import torch
    import torch.distributed as dist
    from smartqsm_b200._lib import Engine
    from smartqsm_b200.sharding import partition_trees

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_pts, leaf_on, which = WORKLOADS[args.workload]
    if args.points:
        n_pts = args.points
    n_trees = args.trees or C4_TREES
    state, wkind = load_state(which)
    if args.workload == "C4":
        sizes_all = c4_sizes(n_trees, n_pts)
        mine = partition_trees(sizes_all, world)[rank]
        job = TreeJob([cached_tree(sizes_all[i], C4_SEED0 + i, leaf_on) for i in mine])
        scaling = "strong"
    elif args.workload == "C5":
        if rank == 0:
            cached_tree(n_pts, args.seed, leaf_on)               # generated once, the other ranks read the /tmp copy
        barrier()
        job = ShardedJob(cached_tree(n_pts, args.seed, leaf_on), rank, world, device)
        scaling = "strong"
    else:
        job = TreeJob([cached_tree(n_pts, args.seed + rank, leaf_on)])       # one tree per rank (tree sharding)
        scaling = "weak"

    def make_engine(monitored: bool, latch: bool):
        e = Engine(device=local, monitored_by_chamfer_distance=monitored, latch_early_stop=latch, **CFG)
        e.load_state_dict(state)
        _MONITORED[id(e)] = monitored
        return e

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(eng, e2e: bool, host_out=None, warm: int = 1, profile: bool = False):
        for _ in range(warm):
            job.step(eng, e2e, host_out)
        if profile:
            eng.profile_enable(True)
            eng.profile_reset()
        barrier()
        eng.timer_start()
        for _ in range(args.steps):
            out = job.step(eng, e2e, host_out)
        ms = eng.timer_stop()
        barrier()
        return ms, out

    eng = make_engine(True, False)
    torch.cuda.synchronize()
    # ---- device-resident timing (the headline `value`)
    for _ in range(args.warmup):
        job.step(eng, False)
    eng.profile_enable(True)
    eng.profile_reset()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    l0 = eng.launch_count()
    eng.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stats, _, sizes = job.step(eng, False)
    dev_ms = eng.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop() if rank == 0 else {}
    launches = eng.launch_count() - l0
    prof = eng.profile_all()
    eng.profile_enable(False)
    total_pts = allsum(float(job.points))
    edges_total = int(allsum(float(sum(s[1] for s in sizes))))     # C5: every rank holds one slice of the edge list

    if args.device_only:
        dev_ms_max = allmax(dev_ms)
        if rank == 0:
            fams = {k: v for k, v in prof.items() if v["launches"] > 0 and v["ms"] > 0}
            print(json.dumps({"ms_per_step": dev_ms_max / args.steps, "value": total_pts / (dev_ms_max / args.steps * 1e-3),
                              "launches": launches,
                              "kernel_families": {k: {"ms_per_step": v["ms"] / args.steps, "GB/s": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                                      "launches_per_step": v["launches"] / args.steps} for k, v in sorted(fams.items())},
                              "env": {k: v for k, v in os.environ.items() if k.startswith("SQSM_")}}))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- end to end through the plug-in boundary (host buffers in, host results out)
    cap_pts = int(job.max_points * 1.1) + 1024
    cap_edges = int(max(s[1] for s in sizes) * 1.05) + 1024       # edge counts of the device-resident run (deterministic)
    host_out = {"P": torch.empty((cap_pts, 3), dtype=torch.float32).pin_memory(),
                "D": torch.empty((cap_pts, 3), dtype=torch.float32).pin_memory(),
                "ids": torch.empty(cap_pts, dtype=torch.int64).pin_memory(),
                "sk": torch.empty((cap_pts, 3), dtype=torch.float32).pin_memory(),
                "rad": torch.empty(cap_pts, dtype=torch.float32).pin_memory(),
                "edges": torch.empty((cap_edges, 2), dtype=torch.int64).pin_memory()}
    e2e_ms, (_, d2h, _) = timed(eng, True, host_out)
    eng.close()

    # ---- the same stage with the early-stop latch (results identical: a rejected thunk leaves the state unchanged, so
    # every later thunk would recompute the same rejection; reported beside the headline, never instead of it)
    eng_l = make_engine(True, True)
    latched_ms, (stats_l, _, sizes_l) = timed(eng_l, False)
    latched_e2e_ms, _ = timed(eng_l, True, host_out)              # ... and end to end through host buffers
    digest = result_digest(eng_l) if (rank == 0 and args.workload == "C5") else None      # rank 0 holds the final state
    eng_l.close()
    if sizes_l != sizes:
        raise RuntimeError(f"latched run produced {sizes_l}, unlatched {sizes}: the latch must not change results")

    # ---- SURVEY 8(d) iteration accounting (b): monitor off, exactly 10 accepted iterations
    eng_b = make_engine(False, False)
    modeb_ms, (stats_b, _, sizes_b) = timed(eng_b, False, profile=True)
    prof_b = eng_b.profile_all()
    eng_b.close()

    dev_ms_max, e2e_ms_max = allmax(dev_ms), allmax(e2e_ms)
    latched_ms_max, modeb_ms_max = allmax(latched_ms), allmax(modeb_ms)
    latched_e2e_ms_max = allmax(latched_e2e_ms)
    h2d_all, d2h_all = int(allsum(float(job.h2d))), int(allsum(float(d2h)))
    launches_all = int(allsum(float(launches)))
    ms_per_step = dev_ms_max / args.steps
    value = total_pts / (ms_per_step * 1e-3)
    e2e_value = total_pts / (e2e_ms_max / args.steps * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") \
            else (6650.0, "fallback (B200_PROFILING.md)")

        def stage_frac(p, ms):
            by = sum(v["bytes"] for k, v in p.items() if k in STAGE_FAMILIES)
            return by / (ms * 1e-3) / 1e9 / hbm_peak, by

        fams = {k: v for k, v in prof.items() if v["launches"] > 0 and v["ms"] > 0}
        kernels = {k: v for k, v in fams.items() if k.startswith(KERNEL_PREFIXES)}        # single-kernel families
        dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
        roofline = None
        if dom:
            f = fams[dom]
            ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
            traffic, traffic_note = None, None
            try:    # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture of this command
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tr.get(args.workload, {}).get(dom)
                traffic_note = tr.get("_note")
            except Exception:
                pass
            sf, sbytes = stage_frac(prof, dev_ms)
            ranked = sorted(kernels, key=lambda k: -kernels[k]["ms"])[:6]
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                        "avg_launch_ms": f["ms"] / f["launches"], "launches": f["launches"],
                        "algorithmic_bytes_per_launch": f["bytes"] / f["launches"],
                        "share_of_step": f["ms"] / dev_ms,
                        "stage_frac": sf, "stage_algorithmic_bytes_per_step": sbytes / args.steps,
                        "kernel_selection": "longest single kernel family among conv_*, cd_*, rb_*, vx_*",
                        "top_kernels": {k: {"ms_per_step": kernels[k]["ms"] / args.steps,
                                            "frac": kernels[k]["bytes"] / (kernels[k]["ms"] * 1e-3) / 1e9 / hbm_peak} for k in ranked}}
        fam_table = {k: {"ms_per_step": v["ms"] / args.steps, "GB/s": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                         "launches_per_step": v["launches"] / args.steps} for k, v in sorted(fams.items())}
        acc = sum(1 for s in stats if s["accepted"])
        n_vox = sum(s["n_vox"] for s in stats)
        far = sum(s.get("far_face_rows", 0) for s in stats)
        sfb, _ = stage_frac(prof_b, modeb_ms)
        n_here = job.max_points
        line = {
            "metric": "points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": f"synthetic; weights: {wkind}",
            "config": {"workload": workload_text(args.workload, n_pts, n_trees, leaf_on),
                       **CFG, "monitored_by_chamfer_distance": True, "latch_early_stop": False,
                       "skeletal_voxel": SKELETAL_VOXEL, "edge_r": EDGE_R, "total_points": int(total_pts),
                       "l2": "inputs larger than L2 (%.0f MB point array per tree, >1 GB of per-iteration tables)" % (n_here * 12 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": launches_all,
            "clocks": clk,
            "roofline": roofline,
            "kernel_families": fam_table,
            "iterations": {"mode": "a: reference configuration, Chamfer monitor on, a rejected iteration is recomputed",
                           "executed": len(stats), "accepted": acc, "rejected": len(stats) - acc,
                           "n_in": [s["n_in"] for s in stats], "n_vox": [s["n_vox"] for s in stats],
                           "rows_l1_l2_l3": [stats[0]["n_rows_l1"], stats[0]["n_rows_l2"], stats[0]["n_rows_l3"]],
                           "pairs_l0_l3": [stats[0]["pairs_l0"], stats[0]["pairs_l1"], stats[0]["pairs_l2"], stats[0]["pairs_l3"]],
                           "n_entries": stats[0]["n_entries"], "n_cubes": stats[0]["n_cubes"],
                           "far_face_rows": far, "far_face_fraction": far / max(n_vox, 1),
                           "stage_ms_iter1": {k: stats[0].get(k) for k in ("ms_tiling", "ms_voxel", "ms_rulebook", "ms_network", "ms_chamfer")},
                           "skeletal_points": [s[0] for s in sizes], "edges": [s[1] for s in sizes]},
            "wall_ms_per_step": wall_ms / args.steps,
            "latched": {"value": total_pts / (latched_ms_max / args.steps * 1e-3), "unit": "points/s",
                        "ms_per_step": latched_ms_max / args.steps,
                        "e2e_value": total_pts / (latched_e2e_ms_max / args.steps * 1e-3),
                        "thunks_computed": sum(1 for s_ in stats_l if not s_["skipped"]),
                        "note": "latch_early_stop=True: thunks after the first rejected iteration are no-ops; "
                                "output identical to the headline run"},
            "mode_b": {"value": total_pts / (modeb_ms_max / args.steps * 1e-3), "unit": "points/s",
                       "ms_per_step": modeb_ms_max / args.steps, "stage_frac": sfb,
                       "accepted": sum(1 for s_ in stats_b if s_["accepted"]), "n_in": [s_["n_in"] for s_ in stats_b],
                       "note": "SURVEY 8(d) accounting (b): monitored_by_chamfer_distance=False, exactly 10 accepted iterations"},
        }
        if args.workload == "C4":
            line["config"]["trees"] = n_trees
            line["config"]["tree_sizes_min_max"] = [min(sizes_all), max(sizes_all)]
        if args.workload == "C5":
            line["result_digest"] = digest
            line["exchange"] = {"collective": "all_gather_into_tensor (NCCL), once per thunk, on the engine's stream",
                                "bytes_per_thunk_per_rank": job.exchange_bytes,
                                "edges_total": edges_total,
                                "replicated_per_rank": "tiling, Chamfer voxel means, skeletal sampling; the radius graph is "
                                                       "produced in node ranges, one per rank (sqsm_radius_graph_range)"}
        if not args.no_cpu_baseline:
            import torch as _t
            from smartqsm_b200.synthetic import make_tree
            cores = os.cpu_count() or 1
            _t.set_num_threads(cores)
            ns = args.cpu_sample_points or cpu_sample_points(20.0)
            sp = make_tree(ns, args.seed, leaf_on)
            t1 = time.perf_counter()
            run_oracle_stage(state, sp)
            dt = time.perf_counter() - t1
            line["cpu_baseline"] = {
                "value": ns / dt, "unit": "points/s", "cores": cores, "kind": "port",
                "sample": f"whole stage (10 thunks, monitor on, sampling, edges) on a {ns}-point synthetic tree "
                          f"from the same generator; restated reference CPU path (oracle/), {dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
