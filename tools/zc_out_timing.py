#!/usr/bin/env python
"""Diagnostic: the 4096-node tpobs_grid launch with its output rows in HBM versus stored straight into pinned host
memory over PCIe (what mppb_eval_nodes does for pinned caller buffers), device-timed."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mrpt_path_planning_b200 import capi, workloads as wl

B, N = 4096, 1 << 20
torch.cuda.init()
stream = torch.cuda.Stream()
ctx = capi.Context(0, stream=stream.cuda_stream)
for p in capi.ackermann_ptgs(ctx=ctx):
    ctx.add_ptg(p)
xs, ys = wl.uniform_cloud(N, 200.0)
ctx.set_obstacles(xs, ys)
nodes = torch.from_numpy(capi.nodes_xycs_from_poses(wl.node_poses(B, 200.0))).cuda()
d_out = torch.empty((B, 242), dtype=torch.float32, device="cuda")
h_out = capi.pinned_empty((B, 242), np.float32)


def run(out, nb):
    ts = []
    for _ in range(30):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream)
            ctx.eval_nodes_device(nb, nodes, 5.0, out)
            b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


for nb in (1024, 2048, 4096):
    print("%d nodes: rows to HBM %.1f us, rows to pinned host %.1f us (%.2f MB)" % (
        nb, run(d_out, nb), run(h_out, nb), nb * 968e-6))
