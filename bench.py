#!/usr/bin/env python
"""Benchmark of the B200 ``spconv-contraction`` path (BASELINE.json metric: points/sec).

A *step* = the whole stage for one synthetic tree: ``max_iteration`` (10) ``_contract`` thunks in
the reference configuration (Chamfer monitor on, no early-stop latch: a rejected iteration is
recomputed exactly like ``core/core_algorithms/spconv_based_contraction.py:123-127`` does), then
skeletal sampling + radii and the radius-neighbour edge list.  points/sec = input points / step time.

* ``value``      inputs already resident in HBM (float32 device array) when the timed region starts;
                 device time from CUDA events on the engine's stream, max over ranks.
* ``e2e``        the same step through the plug-in / C ABI with HOST buffers: float64 input copied
                 from pinned host memory, results (contracted points, displacements, skeletal points,
                 radii, edges) copied back, all inside the timed region.
* ``roofline``   the dominant kernel family, algorithmic bytes over its device time.
* ``cpu_baseline`` / ``--impl reference``  the restated reference CPU path (``oracle/``) on the box's
                 host cores on a bounded sample of the same workload.

Multi-GPU (``torchrun``): trees are independent (north_star: "partitioned by tree"), every rank runs
its own tree (seed = base + rank), no data-path collective; weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (points, leaf_on, checkpoint)
    "C3": (20_000_000, True, "leafon"),    # spconv-contraction-LEAFON, 20M-pt leaf-on TLS tree (the metric's config)
    "C2": (5_000_000, False, "leafoff"),   # spconv-contraction-LEAFOFF, 5M-pt leaf-off tree
    "C1": (1_000_000, False, "leafoff"),   # the reference's CPU-runnable case
}
CFG = dict(batch_size=16, max_iteration=10, voxel_size=0.01, cube_size=4.0, buffer_size=0.4)
SKELETAL_VOXEL = 0.01      # voxel_size_for_downsampling (configs/spconv-contraction-*.yaml)
EDGE_R = 0.055             # edge_search_kwargs.r


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SQSM_BENCH_WORKLOAD", "C3"), choices=list(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override the point count (debugging only)")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-points", type=int, default=0)
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


def cpu_sample_points(budget_s: float) -> int:
    # the oracle runs the whole stage at roughly 2.5-3.5 k points/s on 8-16 cores
    return int(np.clip(budget_s * 2500, 4000, 120_000))


# ------------------------------------------------------------------------------------------ reference arm

def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from smartqsm_b200.synthetic import make_tree
    n_full, leaf_on, which = WORKLOADS[args.workload]
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
              f"{n}-point synthetic {'leaf-on' if leaf_on else 'leaf-off'} tree from the same generator")
    line = {
        "impl": "reference", "metric": "points/sec", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic; weights: {wkind}",
        "config": {"workload": f"{args.workload}: spconv-contraction stage, {n_full}-point synthetic tree "
                               f"(timed on a bounded {n}-point sample)", **CFG,
                   "monitored_by_chamfer_distance": True},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "iterations": {"accepted": sum(1 for h in hist if h.get("accepted")), "executed": len(hist)},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ B200 arm

def one_step(eng, set_input, fetch_host: bool, host_out=None):
    """set_input() loads the tree; then the 10 thunks, sampling and the edge list."""
    set_input()
    stats = [eng.contract_step() for _ in range(CFG["max_iteration"])]
    if fetch_host:
        # results land in pinned host buffers (host_out: dict of pinned torch tensors sized for the worst case)
        n = eng.result_into(host_out["P"].data_ptr(), host_out["D"].data_ptr())
        m = eng.skeletal_sample(SKELETAL_VOXEL, fetch=False)
        eng.skeletal_sample_into(SKELETAL_VOXEL, host_out["ids"].data_ptr(), host_out["sk"].data_ptr(), host_out["rad"].data_ptr())
        e = eng.radius_graph(EDGE_R, fetch=False)
        if e * 16 > host_out["edges"].numel() * host_out["edges"].element_size():
            raise RuntimeError("edge buffer too small")
        eng.radius_graph_into(EDGE_R, host_out["edges"].data_ptr())
        d2h = n * 24 + m * (8 + 12 + 4) + e * 16
        return stats, d2h, (m, e)
    m = eng.skeletal_sample(SKELETAL_VOXEL, fetch=False)
    e = eng.radius_graph(EDGE_R, fetch=False)
    return stats, 0, (m, e)


def main_b200(args):
    import torch
    import torch.distributed as dist
    from smartqsm_b200._lib import Engine
    from smartqsm_b200.synthetic import make_tree

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_pts, leaf_on, which = WORKLOADS[args.workload]
    if args.points:
        n_pts = args.points
    state, wkind = load_state(which)
    pts64 = make_tree(n_pts, args.seed + rank, leaf_on)              # one tree per rank (tree sharding)
    n = len(pts64)
    eng = Engine(device=local, monitored_by_chamfer_distance=True, latch_early_stop=False, **CFG)
    eng.load_state_dict(state)
    host_in = torch.from_numpy(pts64).pin_memory()                    # float64, pinned: e2e input
    dev_in = torch.from_numpy(pts64.astype(np.float32)).cuda()        # float32, resident: `value` input
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    set_dev = lambda: eng.set_points_device_f32(dev_in.data_ptr(), n)
    set_host = lambda: eng.set_points_ptr_f64(host_in.data_ptr(), n)

    # ---- device-resident timing
    for _ in range(args.warmup):
        one_step(eng, set_dev, False)
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
        stats, _, sizes = one_step(eng, set_dev, False)
    dev_ms = eng.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clk = clocks.stop() if rank == 0 else {}
    launches = eng.launch_count() - l0
    prof = eng.profile_all()
    eng.profile_enable(False)

    # ---- end to end through the plug-in boundary (host buffers in, host results out)
    cap_pts = int(n * 1.1) + 1024
    cap_edges = int(sizes[1] * 1.05) + 1024           # edge count of the device-resident run (same input, deterministic)
    host_out = {"P": torch.empty((cap_pts, 3), dtype=torch.float32).pin_memory(),
                "D": torch.empty((cap_pts, 3), dtype=torch.float32).pin_memory(),
                "ids": torch.empty(cap_pts, dtype=torch.int64).pin_memory(),
                "sk": torch.empty((cap_pts, 3), dtype=torch.float32).pin_memory(),
                "rad": torch.empty(cap_pts, dtype=torch.float32).pin_memory(),
                "edges": torch.empty((cap_edges, 2), dtype=torch.int64).pin_memory()}
    one_step(eng, set_host, True, host_out)
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        _, d2h, _ = one_step(eng, set_host, True, host_out)
    e2e_ms = eng.timer_stop()
    barrier()

    # ---- the same stage with the early-stop latch (results identical: a rejected thunk leaves the state unchanged, so
    # every later thunk would recompute the same rejection; reported beside the headline, never instead of it)
    eng.close()
    eng_l = Engine(device=local, monitored_by_chamfer_distance=True, latch_early_stop=True, **CFG)
    eng_l.load_state_dict(state)
    set_dev_l = lambda: eng_l.set_points_device_f32(dev_in.data_ptr(), n)
    one_step(eng_l, set_dev_l, False)
    barrier()
    eng_l.timer_start()
    for _ in range(args.steps):
        stats_l, _, sizes_l = one_step(eng_l, set_dev_l, False)
    latched_ms = eng_l.timer_stop()
    barrier()
    if sizes_l != sizes:
        raise RuntimeError(f"latched run produced {sizes_l}, unlatched {sizes}: the latch must not change results")

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

    dev_ms_max = allmax(dev_ms)
    e2e_ms_max = allmax(e2e_ms)
    latched_ms_max = allmax(latched_ms)
    h2d_all, d2h_all = int(allsum(float(n * 24))), int(allsum(float(d2h)))
    total_pts = allsum(float(n))
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
        fams = {k: v for k, v in prof.items() if v["launches"] > 0 and v["ms"] > 0}
        kernels = {k: v for k, v in fams.items() if k.startswith("conv_")}     # one entry per convolution kernel shape
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
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                        "avg_launch_ms": f["ms"] / f["launches"], "launches": f["launches"],
                        "algorithmic_bytes_per_launch": f["bytes"] / f["launches"],
                        "gflops_fp32": f["flops"] / (f["ms"] * 1e-3) / 1e9 if f["flops"] else None,
                        "share_of_step": f["ms"] / dev_ms}
        fam_table = {k: {"ms_per_step": v["ms"] / args.steps, "GB/s": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                         "launches_per_step": v["launches"] / args.steps} for k, v in sorted(fams.items())}
        # families nest: stage families (tiling, voxelise, rulebook, subm_l*, down_up, heads, chamfer, skeletal,
        # radius_graph) partition the step; conv_* / cd_* / rb_* / vx_* are per-kernel views inside them
        acc = sum(1 for s in stats if s["accepted"])
        line = {
            "metric": "points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": f"synthetic; weights: {wkind}",
            "config": {"workload": f"{args.workload}: spconv-contraction stage on a synthetic {n}-point "
                                   f"{'leaf-on' if leaf_on else 'leaf-off'} tree per GPU (tree-sharded)",
                       **CFG, "monitored_by_chamfer_distance": True, "latch_early_stop": False,
                       "skeletal_voxel": SKELETAL_VOXEL, "edge_r": EDGE_R,
                       "l2": "inputs larger than L2 (%.0f MB point array, >1 GB of per-iteration tables)" % (n * 12 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": launches_all,
            "clocks": clk,
            "roofline": roofline,
            "kernel_families": fam_table,
            "iterations": {"executed": len(stats), "accepted": acc, "rejected": len(stats) - acc,
                           "n_in": [s["n_in"] for s in stats], "n_vox": [s["n_vox"] for s in stats],
                           "rows_l1_l2_l3": [stats[0]["n_rows_l1"], stats[0]["n_rows_l2"], stats[0]["n_rows_l3"]],
                           "pairs_l0_l3": [stats[0]["pairs_l0"], stats[0]["pairs_l1"], stats[0]["pairs_l2"], stats[0]["pairs_l3"]],
                           "n_entries": stats[0]["n_entries"], "n_cubes": stats[0]["n_cubes"],
                           "stage_ms_iter1": {k: stats[0][k] for k in ("ms_tiling", "ms_voxel", "ms_rulebook", "ms_network", "ms_chamfer")},
                           "skeletal_points": sizes[0], "edges": sizes[1]},
            "wall_ms_per_step": wall_ms / args.steps,
            "latched": {"value": total_pts / (latched_ms_max / args.steps * 1e-3), "unit": "points/s",
                        "ms_per_step": latched_ms_max / args.steps,
                        "thunks_computed": sum(1 for s_ in stats_l if not s_["skipped"]),
                        "note": "latch_early_stop=True: thunks after the first rejected iteration are no-ops; "
                                "output identical to the headline run"},
        }
        if not args.no_cpu_baseline:
            import torch as _t
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
