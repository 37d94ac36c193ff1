#!/usr/bin/env python
"""Kernel time of the C3 batch as a function of the cloud index's bin size (device-resident inputs)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrpt_path_planning_b200 import capi, workloads as wl

stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream)
for p in capi.ackermann_ptgs(ctx=ctx):
    ctx.add_ptg(p)
xs, ys = wl.uniform_cloud(1 << 20, 200.0)
poses = wl.node_poses(4096, 200.0)
with torch.cuda.stream(stream):
    d_nodes = torch.from_numpy(capi.nodes_xycs_from_poses(poses)).cuda()
    d_out = torch.empty((4096, 242), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for bs in [float(a) for a in sys.argv[1:]] or [0.5, 0.75, 1.0, 1.5]:
    ctx.set_obstacles(xs, ys, bin_size=bs)
    ts = []
    with torch.cuda.stream(stream):
        for i in range(25):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ctx.eval_nodes_device(4096, d_nodes, 5.0, d_out); b.record(stream)
            ts.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ts[5:])
    print("bin %.2f m: kernel median %.4f ms  (min %.4f)" % (bs, t[len(t) // 2], t[0]))
