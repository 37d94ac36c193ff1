#!/usr/bin/env python
"""Diagnostic: one C1 or C2 plan (the reference's demo query) repeated a few times, for an ncu launch list of the
kernels a single query launches (which kernel, how long) next to the planner's own phase timers.

  python tools/plan_latency_diag.py c1|c2
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
xs, ys, start, goal, lo, hi = bench.demo_inputs()
ctx = capi.Context(0)
ptgs = capi.holonomic_ptgs() if which == "c1" else capi.ackermann_ptgs()
pl = capi.Planner(ctx, ptgs, wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
pl.add_obstacles(xs, ys)
pl.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
if which == "c2":
    pl.add_waypoints(bench.WAYPOINTS01[0], bench.WAYPOINTS01[1], 1.5, 0.5, True)
for i in range(4):
    r = pl.plan(start, goal, lo, hi)
    st = pl.gpu_stats()
    print("plan %d: %.3f ms, %d expansions, phases (ms): %s" % (
        i, 1e3 * r["time_s"], r["expanded"], {k: round(1e3 * v, 3) for k, v in st["phase_seconds"].items()}))
pl.close()
