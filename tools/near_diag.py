#!/usr/bin/env python
"""Near-first evaluation of tpobs_grid_kernel on the C3 workload: device-resident timing per mode and how many nodes
needed the second launch.  python tools/near_diag.py [walls]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

walls = len(sys.argv) > 1 and sys.argv[1] == "walls"
xs, ys = (wl.walls_cloud(1 << 20, 200.0) if walls else wl.uniform_cloud(1 << 20, 200.0))
poses = wl.node_poses(4096, 200.0, seed=wl.C3_NODES_SEED + int(os.environ.get("NODE_SEED_OFFSET", "0")))
ctx = capi.Context(0)
for p in capi.ackermann_ptgs():
    ctx.add_ptg(p)
ctx.set_obstacles(xs, ys)
xycs = torch.from_numpy(capi.nodes_xycs_from_poses(poses)).cuda()
out = torch.empty((4096, ctx.total_paths), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for mode in [int(m) for m in os.environ.get("NEAR_MODES", "0,2,0,2").split(",")]:
    ctx.set_grid_near(mode)
    s0 = ctx.grid_near_stats()
    ts = []
    for it in range(25):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.timer_begin()
        ctx.eval_nodes_device(4096, xycs.data_ptr(), 5.0, out.data_ptr())
        ms = ctx.timer_end()
        if it >= 5:
            ts.append(ms)
    s1 = ctx.grid_near_stats()
    o = out.cpu().numpy()
    if ref is None:
        ref = o
    print("mode %d: %.4f ms (min %.4f) per 4096-node batch; two-launch calls %d, nodes %d, needed the second launch %d; identical %s" % (
        mode, float(np.median(ts)), min(ts), s1[0] - s0[0], s1[1] - s0[1], s1[2] - s0[2], np.array_equal(o.view(np.uint32), ref.view(np.uint32))))
ctx.close()
