#!/usr/bin/env python
"""Diagnostic: latency of single-node mppb_eval_nodes calls (one A* expansion) at several cloud densities."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrpt_path_planning_b200 import capi, workloads as wl

ctx = capi.Context(0)
for p in capi.ackermann_ptgs(ctx=ctx):
    ctx.add_ptg(p)
ext = 50.0
for dens in [float(a) for a in sys.argv[1:]] or [8.0]:
    xs, ys = wl.uniform_cloud(max(1, int(dens * ext * ext)), ext, seed=7)
    ctx.set_obstacles(xs, ys)
    P = wl.node_poses(48, ext, margin=1.0, seed=8)
    poses = capi.pinned_empty((1, 3), np.float64)
    rows = capi.pinned_empty((1, 242), np.float32)
    med = []
    for i in range(len(P)):
        poses[:] = P[i]
        for _ in range(3):
            ctx.eval_nodes(poses, 5.0, out=rows)
        ts = []
        for _ in range(25):
            t = time.perf_counter(); ctx.eval_nodes(poses, 5.0, out=rows); ts.append(time.perf_counter() - t)
        ts.sort()
        med.append(1e6 * ts[len(ts) // 2])
    med.sort()
    print("density %5.2f pts/m2 (~%4d pts in window): single-node latency min %.1f  median %.1f  max %.1f us" % (
        dens, int(dens * 100), med[0], med[len(med) // 2], med[-1]))
