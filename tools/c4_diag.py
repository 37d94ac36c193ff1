#!/usr/bin/env python
"""Diagnostic (GPU box): config C4 on the pillars map at several expansion caps — success rate, expansions, wall time and
the planner's phase timers. Usage: python tools/c4_diag.py [queries] [cap cap ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
caps = [int(a) for a in sys.argv[2:]] or [600, 1500, 4000]
threads = int(os.environ.get("C4_THREADS", "0")) or (os.cpu_count() or 1)
mx, my, pillars = wl.pillars_map(1 << 20, 200.0)
qs = wl.planning_queries_free(nq, pillars, 200.0)
ctx = capi.Context(0)
ptgs = capi.ackermann_ptgs()
for cap in caps:
    params = list(wl.DEMO_ASTAR_PARAMS)
    params[8] = cap
    pl = capi.Planner(ctx, ptgs, params, wl.DEMO_ASTAR_TIMESTAMPS)
    pl.add_obstacles(mx, my)
    if not os.environ.get("C4_NO_COSTMAP"):
        pl.add_costmap(mx, my, bench.C4_COSTMAP[0], bench.C4_COSTMAP[1], bench.C4_COSTMAP[2], False, 0.0, None)
    pl.plan_many(qs[:16], n_threads=threads)
    t0 = time.perf_counter()
    res = pl.plan_many(qs, n_threads=threads)
    dt = time.perf_counter() - t0
    st = pl.gpu_stats()
    exp = sum(r["expanded"] for r in res)
    ok = [r for r in res if r["success"]]
    print(json.dumps({"cap": cap, "queries": nq, "threads": threads, "batch_s": round(dt, 3), "successes": len(ok),
                      "expansions": int(exp), "exp_per_s": round(exp / dt), "plans_per_s": round(nq / dt, 1),
                      "mean_exp_of_successes": round(sum(r["expanded"] for r in ok) / max(1, len(ok)), 1),
                      "gpu_s": round(st["gpu_seconds"], 3),
                      "phases": {k: round(v, 3) for k, v in st["phase_seconds"].items() if v > 0}}))
    pl.close()
ctx.close()
