#!/usr/bin/env python
"""Latency of the host-buffer entry points on planner-sized batches (one A* expansion), pinned buffers.
Run twice (MPPB_NO_ZEROCOPY=1 for the staged-copy path) to compare."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402


def bench(fn, n=300):
    for _ in range(20):
        fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    ts.sort()
    return 1e6 * ts[len(ts) // 2]


L = capi.lib()
ctx = capi.Context(0)
for p in capi.ackermann_ptgs(ctx=ctx):
    ctx.add_ptg(p)
xs, ys = wl.uniform_cloud(20000, 50.0, seed=7)
ctx.set_obstacles(xs, ys)
out = {}
for n in (1, 16, 512):
    poses = capi.pinned_empty((n, 3), np.float64)
    poses[:] = wl.node_poses(n, 50.0, margin=1.0, seed=8)
    rows = capi.pinned_empty((n, 242), np.float32)
    out["eval_nodes_%d_us" % n] = bench(lambda: ctx.eval_nodes(poses, 5.0, out=rows))
cm = np.random.default_rng(1).uniform(0, 1, (775, 775))
ctx.set_costmap(0, cm, -15.0, -15.0, 0.04, False)
ctx.set_waypoints(1, [3.0, 5.0, 9.0], [2.0, 5.0, 8.0], 1.5, 0.5, True)
rng = np.random.default_rng(3)
for E in (20, 200):
    frm = capi.pinned_empty((E, 3), np.float64)
    frm[:] = np.stack([rng.uniform(0, 10, E), rng.uniform(0, 10, E), rng.uniform(-3, 3, E)], 1)
    off = capi.pinned_empty((E + 1,), np.uint32)
    off[:] = np.arange(E + 1) * 7
    rel = capi.pinned_empty((7 * E, 2), np.float64)
    rel[:] = rng.uniform(-2, 2, (7 * E, 2))
    res = capi.pinned_empty((E, 2), np.float64)
    fn = lambda: L.mppb_eval_edge_costs(ctx.h, E, capi._ptr(frm), capi._ptr(off), capi._ptr(rel), capi._ptr(res))
    out["edge_costs_%d_us" % E] = bench(fn)
ctx.close()
ctx = capi.Context(0)
h = capi.holonomic_ptgs()[0]
ctx.add_ptg(h)
obs = np.loadtxt(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "obstacles_01.txt")).astype(np.float32)
ctx.set_obstacles(obs[:, 0].copy(), obs[:, 1].copy())
h.update_dyn_state((0.3, 0.1, 0.0), (3.0, 2.0, 0.5), 0.0)
for M in (42, 129):
    cands = capi.pinned_empty((M,), capi.HOLO_CAND_DTYPE)
    for i in range(M):
        h.set_trim_speed([1 / 3, 2 / 3, 1.0][i % 3])
        c = h.holo_params((i * 4) % 191)
        cands[i] = (0, 0, c.T_ramp, c.vxi, c.vyi, c.vxf, c.vyf, c.vf)
    poses = capi.pinned_empty((1, 3), np.float64)
    poses[:] = [[1.0, 0.5, 0.2]]
    res = capi.pinned_empty((M,), np.float64)
    fn = lambda: L.mppb_eval_holo_candidates(ctx.h, 1, capi._ptr(poses), M, capi._ptr(cands), 5.0, capi._ptr(res))
    out["holo_%d_us" % M] = bench(fn)
out["zero_copy"] = os.environ.get("MPPB_NO_ZEROCOPY") is None
print(json.dumps(out))
