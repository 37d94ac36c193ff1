#!/usr/bin/env python
"""Diagnostic (GPU box): latency of ONE node's expansion call (the single-query planner's per-expansion GPU round trip)
on the reference's demo map: wall time of mppb_expand_nodes through the binding, device time between CUDA events, and
the collision-only call for comparison."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

xs, ys, start, goal, lo, hi = bench.demo_inputs()
ctx = capi.Context(0)
ptgs = capi.ackermann_ptgs()
for p in ptgs:
    ctx.add_ptg(p)
ctx.set_obstacles_clipped(xs, ys, [lo[0], lo[1], hi[0], hi[1]])
pl = capi.Planner(ctx, ptgs, wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)  # only to build the evaluators the same way
ctx.set_expand_params(wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
nodes = capi.pinned_empty((1,), capi.EXPAND_NODE_DTYPE)
nodes["pose"][0] = start[:3]
nodes["bbox_min"][0] = lo
nodes["bbox_max"][0] = hi
nodes["goal_k"][0][:] = -1
nodes["goal_step"][0][:] = 0xFFFFFFFF
stride = ctx.expand_max_neighbors
out = capi.pinned_empty((1, stride), capi.NEIGHBOR_DTYPE)
counts = capi.pinned_empty((1, 2), np.uint32)
L = capi.lib()


def call():
    rc = L.mppb_expand_nodes(ctx.h, 1, nodes.ctypes.data, 0, 5.0, out.ctypes.data, counts.ctypes.data)
    assert rc == 0


for _ in range(200):
    call()
N = 2000
t0 = time.perf_counter()
for _ in range(N):
    call()
wall = (time.perf_counter() - t0) / N
dev = []
for _ in range(200):
    ctx.timer_begin()
    L.mppb_expand_nodes_begin(ctx.h, 1, nodes.ctypes.data, 0, 5.0, out.ctypes.data, counts.ctypes.data)
    dev.append(ctx.timer_end())
poses = capi.pinned_empty((1, 3), np.float64)
poses[0] = start[:3]
row = capi.pinned_empty((1, ctx.total_paths), np.float32)
for _ in range(200):
    ctx.eval_nodes(poses, 5.0, out=row)
t0 = time.perf_counter()
for _ in range(N):
    ctx.eval_nodes(poses, 5.0, out=row)
wall_rows = (time.perf_counter() - t0) / N
print("expand 1 node: wall %.1f us per call, device %.1f us between events (median); neighbours %d; collision-only call %.1f us" % (
    1e6 * wall, 1e3 * float(np.median(dev)), int(counts[0, 0]), 1e6 * wall_rows))
