#!/usr/bin/env python
"""Diagnostic (no GPU): a C4-shaped batch through the product planner's HOST logic with the device entry points answered
by the oracle (tests/hostsim, LD_PRELOAD) — the planner's host phase timers (F+pop+A, C/D) are then measurable on any
machine; the "GPU" phases are the stand-in's and mean nothing.
   LD_PRELOAD=tests/_build/libmppb_sim.so python tools/hostsim_c4.py [queries] [threads] [max_expansions]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import run_plans as rp  # noqa: E402
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402
from oracle import orc  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nt = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 300
extent = 60.0
mx, my = wl.walls_cloud(8000, extent, wall_fraction=0.9, seed=33)
qs = rp.random_queries(np.random.default_rng(5), nq, extent, 8.0, 25.0)
params = list(wl.DEMO_ASTAR_PARAMS)
params[8] = cap
o_ptgs = orc.ackermann_ptgs()
rp.register_ptgs(o_ptgs)
ctx = rp.SimContext()
pl = capi.Planner(ctx, capi.ackermann_ptgs(), params, wl.DEMO_ASTAR_TIMESTAMPS)
pl.add_obstacles(mx, my)
for rep in range(2):
    t0 = time.perf_counter()
    res = pl.plan_many(qs, n_threads=nt)
    dt = time.perf_counter() - t0
    st = pl.gpu_stats()
    ph = st["phase_seconds"]
    ex = sum(r["expanded"] for r in res)
    print("%d queries, %d threads, %d expansions, %.2f s; host phases: F+pop+A %.3f s (%.1f us/exp/thread), C/D %.3f s (%.1f us/exp/thread)" % (
        nq, nt, ex, dt, ph["F_relax+pop+A_enumerate"], 1e6 * ph["F_relax+pop+A_enumerate"] * nt / ex,
        ph["CD_select_build"], 1e6 * ph["CD_select_build"] * nt / ex))
