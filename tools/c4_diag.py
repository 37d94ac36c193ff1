#!/usr/bin/env python
"""Diagnostic: the C4-shaped batch of bench.py alone (independent queries in lock-step on the wall map), for A/B runs of
host-side changes:   python tools/c4_diag.py [queries] [threads]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nt = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
mx, my = wl.walls_cloud(1 << 20, bench.EXTENT, wall_fraction=1.0, spacing=0.008, seed=wl.C3_WALLS_SEED)
qs = wl.planning_queries(nq)
params = list(wl.DEMO_ASTAR_PARAMS)
params[8] = bench.C4_MAX_EXPANSIONS
ctx = capi.Context(0)
pl = capi.Planner(ctx, capi.ackermann_ptgs(), params, wl.DEMO_ASTAR_TIMESTAMPS)
pl.add_obstacles(mx, my)
pl.plan_many(qs[:8], n_threads=nt)
for rep in range(3):
    t0 = time.perf_counter()
    res = pl.plan_many(qs, n_threads=nt)
    dt = time.perf_counter() - t0
    st = pl.gpu_stats()
    print("C4 %d queries, %d threads: %.3f s -> %.0f plans/s, %d expansions; phases (s): %s" % (
        nq, nt, dt, nq / dt, sum(r["expanded"] for r in res), {k: round(v, 3) for k, v in st["phase_seconds"].items()}))
pl.close()
ctx.close()
