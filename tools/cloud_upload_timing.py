#!/usr/bin/env python
"""Steady-state time of mppb_set_obstacles (upload + bin index) for a 2^20-point cloud."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrpt_path_planning_b200 import capi, workloads as wl
ctx = capi.Context(0)
xs, ys = wl.uniform_cloud(1 << 20, 200.0)
for i in range(6):
    t = time.perf_counter(); ctx.set_obstacles(xs, ys); ctx.sync(); print("set_obstacles #%d: %.3f ms" % (i, 1e3 * (time.perf_counter() - t)))
px = capi.pinned_empty((1 << 20,), np.float32); py = capi.pinned_empty((1 << 20,), np.float32); px[:] = xs; py[:] = ys
for i in range(3):
    t = time.perf_counter(); ctx.set_obstacles(px, py); ctx.sync(); print("pinned input #%d: %.3f ms" % (i, 1e3 * (time.perf_counter() - t)))
