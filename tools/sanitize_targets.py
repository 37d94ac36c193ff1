#!/usr/bin/env python
"""Small invocation of every kernel (dense and sparse phase 2, boxed variant, HolonomicBlend with and without filter,
edge costs, costmap distances, collision-grid build, one plan) for compute-sanitizer runs:
  compute-sanitizer --tool memcheck  python tools/sanitize_targets.py
  compute-sanitizer --tool racecheck python tools/sanitize_targets.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrpt_path_planning_b200 import capi, workloads as wl

ctx = capi.Context(0)
ptgs = capi.ackermann_ptgs(num_paths=31, ctx=ctx)  # colgrid_* kernels
for p in ptgs:
    ctx.add_ptg(p)
for dens in (0.3, 3.0, 30.0):  # sparse, boundary, dense phase 2
    xs, ys = wl.uniform_cloud(int(dens * 900), 30.0, seed=int(dens * 10))
    ctx.set_obstacles(xs, ys)
    poses = wl.node_poses(24, 30.0, margin=0.5, seed=3)
    out = ctx.eval_nodes(poses, 5.0)
    boxes = np.tile(np.float32([5, 5, 25, 25]), (len(poses), 1))
    outb = ctx.eval_nodes_boxed(poses, boxes, 5.0)
    print("grid density %.1f: constrained %.2f / boxed %.2f" % (dens, np.isfinite(out).mean(), np.isfinite(outb).mean()))
ctx.set_costmap(0, np.random.default_rng(1).uniform(0, 1, (200, 200)), -5.0, -5.0, 0.2, False)
ctx.set_waypoints(1, [3.0, 5.0, 9.0], [2.0, 5.0, 8.0], 1.5, 0.5, True)
rng = np.random.default_rng(2)
E = 50
frm = np.stack([rng.uniform(0, 30, E), rng.uniform(0, 30, E), rng.uniform(-3, 3, E)], 1)
print("edge costs", ctx.eval_edge_costs(frm, (np.arange(E + 1) * 7).astype(np.uint32), rng.uniform(-2, 2, (7 * E, 2))).sum())
d2 = ctx.costmap_nearest_d2(xs, ys, 0.0, 0.0, 0.25, 100, 100, 0.5) if hasattr(ctx, "costmap_nearest_d2") else None
ctx.close()

ctx = capi.Context(0)
h = capi.holonomic_ptgs()[0]
ctx.add_ptg(h)
xs, ys = wl.uniform_cloud(3000, 20.0, seed=5)
ctx.set_obstacles(xs, ys)
rows = []
for b in range(6):
    for sp in (1 / 3, 1.0):
        h.set_trim_speed(sp)
        h.update_dyn_state([0.3 * (b % 2), 0.1, 0.0], [4.0, 1.0, 0.0], 0.0, force=True)
        for k in range(0, 191, 12):
            p = h.holo_params(k)
            rows.append((b, 0, p.T_ramp, p.vxi, p.vyi, p.vxf, p.vyf, p.vf))
cands = np.array(rows, dtype=capi.HOLO_CAND_DTYPE)
poses = wl.node_poses(6, 20.0, margin=2.0, seed=6)
a = ctx.eval_holo_candidates(poses, cands, 5.0)
ctx.set_holo_filter(False)
b = ctx.eval_holo_candidates(poses, cands, 5.0)
print("holo filtered == unfiltered:", np.array_equal(a, b), np.isfinite(a).mean())
ctx.close()

ctx = capi.Context(0)
obs = np.loadtxt(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "obstacles_01.txt")).astype(np.float32)
pl = capi.Planner(ctx, capi.holonomic_ptgs(), wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
pl.add_obstacles(obs[:, 0].copy(), obs[:, 1].copy())
pl.add_costmap(obs[:, 0].copy(), obs[:, 1].copy(), 0.04, 0.5, 4.0, False, 15.0, [0.5, 0, 0])
r = pl.plan([0.5, 0, 0, 0, 0, 0], [4, 2.5, np.deg2rad(45.0), 0, 0, 0], [-1.5, -2.0, -np.pi], [6.0, 4.5, np.pi])
print("plan", r["success"], r["expanded"])
pl.close()
ctx.close()
