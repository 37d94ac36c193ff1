#!/usr/bin/env python
"""Diagnostic: does what else lives in the process change the single-query plan time? (bench.py measured C1 at 3.2 ms
where tools/plan_latency_diag.py measures 2.65 ms)   python tools/plan_env_diag.py none|torch|checker|cloud"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
what = sys.argv[1] if len(sys.argv) > 1 else "none"
if what == "torch":
    import torch
    torch.zeros(1 << 20, device="cuda").sum().item()
import bench  # noqa: E402
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

if what == "checker":
    orc, kind = bench.cpu_checker()
    print("checker:", kind)
if what == "cloud":
    big = capi.Context(0)
    for p in capi.ackermann_ptgs():
        big.add_ptg(p)
    xs, ys = wl.uniform_cloud(1 << 20, 200.0)
    big.set_obstacles(xs, ys)
    big.eval_nodes(wl.node_poses(4096, 200.0), 5.0)
xs, ys, start, goal, lo, hi = bench.demo_inputs()
ctx = capi.Context(0)
pl = capi.Planner(ctx, capi.holonomic_ptgs(), wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
pl.add_obstacles(xs, ys)
pl.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
ts = [pl.plan(start, goal, lo, hi)["time_s"] for _ in range(6)]
st = pl.gpu_stats()
print("%-8s C1 plan min %.3f ms; last plan phases (ms): %s" % (
    what, 1e3 * min(ts[1:]), {k: round(1e3 * v, 3) for k, v in st["phase_seconds"].items()}))
