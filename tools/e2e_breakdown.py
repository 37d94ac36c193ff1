#!/usr/bin/env python
"""Diagnostic: where the end-to-end time of mppb_eval_nodes goes (host cos/sin, H2D, kernel, D2H)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrpt_path_planning_b200 import capi, workloads as wl

B, N = 4096, 1 << 20
ctx = capi.Context(0)
for p in capi.ackermann_ptgs():
    ctx.add_ptg(p)
xs, ys = wl.uniform_cloud(N, 200.0)
poses = wl.node_poses(B, 200.0)
ctx.set_obstacles(xs, ys)
h_poses = capi.pinned_empty((B, 3), np.float64); h_poses[:] = poses
h_out = capi.pinned_empty((B, 242), np.float32)
for _ in range(5):
    ctx.eval_nodes(h_poses, 5.0, out=h_out)
t = time.perf_counter()
for _ in range(20):
    capi.nodes_xycs_from_poses(h_poses)
print("host cos/sin (4096 poses): %.1f us" % ((time.perf_counter() - t) / 20 * 1e6))
d_out = torch.empty((B, 242), dtype=torch.float32, device="cuda")
h_t = torch.empty((B, 242), dtype=torch.float32).pin_memory()
torch.cuda.synchronize()
for name, fn in (("D2H 3.96MB", lambda: h_t.copy_(d_out, non_blocking=True)),):
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(name, "%.1f us" % (min(ts) * 1e6))
ts = []
for _ in range(20):
    t = time.perf_counter(); ctx.eval_nodes(h_poses, 5.0, out=h_out); ts.append(time.perf_counter() - t)
print("e2e eval_nodes: min %.1f us, median %.1f us" % (min(ts) * 1e6, sorted(ts)[10] * 1e6))
