#!/usr/bin/env python
"""Small driver that launches each product kernel a few times on bench-shaped inputs (for ncu captures).

  python tools/profile_targets.py [grid|holo|cost|expand|all]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
xs, ys = wl.uniform_cloud(1 << 20, 200.0)
poses = wl.node_poses(4096, 200.0)

if what in ("grid", "all"):
    ctx = capi.Context(0)
    for p in capi.ackermann_ptgs():
        ctx.add_ptg(p)
    ctx.set_obstacles(xs, ys)
    for _ in range(4):
        t = time.perf_counter()
        ctx.eval_nodes(poses, 5.0)
        print("grid: eval_nodes 4096 nodes %.3f ms" % (1e3 * (time.perf_counter() - t)))
    ctx.close()

if what in ("holo", "all"):
    ptg = capi.holonomic_ptgs()[0]
    ctx = capi.Context(0)
    ctx.add_ptg(ptg)
    ctx.set_obstacles(xs, ys)
    nb = int(os.environ.get("HOLO_NODES", "512"))
    ks = [0, 15, 31, 47, 63, 79, 95, 110, 126, 142, 158, 174, 190, 101]
    rows = []
    rng = np.random.default_rng(1)
    for b in range(nb):
        v = rng.uniform(-0.7, 0.7, 2) * (b % 3 != 0)
        for sp in (1 / 3, 2 / 3, 1.0):
            ptg.set_trim_speed(sp)
            ptg.update_dyn_state([v[0], v[1], 0.0], [5.0, 1.0, 0.0], 0.0, force=True)
            for k in ks:
                p = ptg.holo_params(k)
                rows.append((b, 0, p.T_ramp, p.vxi, p.vyi, p.vxf, p.vyf, p.vf))
    cands = capi.pinned_empty((len(rows),), capi.HOLO_CAND_DTYPE)
    cands[:] = np.array(rows, dtype=capi.HOLO_CAND_DTYPE)
    hp = capi.pinned_empty((nb, 3), np.float64)
    hp[:] = poses[:nb]
    out = capi.pinned_empty((len(rows),), np.float64)
    L = capi.lib()
    ctx.set_holo_filter(int(os.environ.get("HOLO_FILTER", "1")))
    for _ in range(5):
        ctx.timer_begin()
        t = time.perf_counter()
        rc = L.mppb_eval_holo_candidates(ctx.h, nb, hp.ctypes.data, len(cands), cands.ctypes.data, 5.0, out.ctypes.data)
        dt = time.perf_counter() - t
        dev_ms = ctx.timer_end()
        assert rc == 0
        print("holo: %d candidates (%d nodes) wall %.3f ms, device (copies + kernel, CUDA events) %.3f ms -> %.3g candidates/s, finite %.2f" % (
            len(cands), nb, 1e3 * dt, dev_ms, len(cands) / dt, np.isfinite(out).mean()))
    ctx.close()

if what in ("cost", "all"):
    ctx = capi.Context(0)
    rng = np.random.default_rng(2)
    cells = rng.uniform(0, 4, (775, 775)) * (rng.uniform(size=(775, 775)) < 0.1)
    ctx.set_costmap(0, cells, 0.0, 0.0, 0.04, False)
    ctx.set_waypoints(1, rng.uniform(0, 31, 5), rng.uniform(0, 31, 5), 1.5, 0.5, True)
    n = 200000
    frm = np.stack([rng.uniform(0, 31, n), rng.uniform(0, 31, n), rng.uniform(-3, 3, n)], 1)
    off = (np.arange(n + 1) * 7).astype(np.uint32)
    rel = rng.uniform(-2, 2, (7 * n, 2))
    for _ in range(3):
        t = time.perf_counter()
        ctx.eval_edge_costs(frm, off, rel)
        dt = time.perf_counter() - t
        print("cost: %d edges x 2 evaluators %.3f ms -> %.3g edges/s" % (n, 1e3 * dt, n / dt))
    ctx.close()

if what in ("expand", "all"):
    # one C4-shaped round: 4096 nodes on the pillars map, both Ackermann PTGs, obstacle costmap, demo parameters
    import bench

    mx, my, pillars = wl.pillars_map(1 << 20, 200.0)
    ctx = capi.Context(0)
    ptgs = capi.ackermann_ptgs()
    for p in ptgs:
        ctx.add_ptg(p)
    ctx.set_obstacles(mx, my)
    pl = capi.Planner(ctx, ptgs, wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
    pl.add_obstacles(mx, my)
    pl.add_costmap(mx, my, bench.C4_COSTMAP[0], bench.C4_COSTMAP[1], bench.C4_COSTMAP[2], False, 0.0, None)
    info, cells = pl.costmap_cells(0)
    ctx.set_costmap(0, cells, info["xmin"], info["ymin"], info["res"], False)
    ctx.set_expand_params(wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
    qs = wl.planning_queries_free(4096, pillars, 200.0)
    nodes = capi.pinned_empty((len(qs),), capi.EXPAND_NODE_DTYPE)
    for i, (s0, g0, lo, hi, _) in enumerate(qs):
        nodes["pose"][i] = s0[:3]
        nodes["bbox_min"][i] = lo
        nodes["bbox_max"][i] = hi
        c, sn = np.cos(s0[2]), np.sin(s0[2])
        rx, ry = (g0[0] - s0[0]) * c + (g0[1] - s0[1]) * sn, -(g0[0] - s0[0]) * sn + (g0[1] - s0[1]) * c
        nodes["goal_k"][i][:] = -1
        nodes["goal_step"][i][:] = 0xFFFFFFFF
        for p, ptg in enumerate(ptgs):
            exact, k, dn = ptg.inverse_map(rx, ry, 0.4)
            if exact:
                nodes["goal_k"][i][p] = k
                found, step = ptg.step_for_dist(k, dn)
                if found:
                    nodes["goal_step"][i][p] = step
    stride = ctx.expand_max_neighbors
    out = capi.pinned_empty((len(qs), stride), capi.NEIGHBOR_DTYPE)
    counts = capi.pinned_empty((len(qs), 2), np.uint32)
    L = capi.lib()
    for _ in range(4):
        ctx.timer_begin()
        t = time.perf_counter()
        rc = L.mppb_expand_nodes(ctx.h, len(qs), nodes.ctypes.data, 1, 5.0, out.ctypes.data, counts.ctypes.data)
        dt = time.perf_counter() - t
        dev_ms = ctx.timer_end()
        assert rc == 0
        print("expand: %d nodes wall %.3f ms, device (copies + 2 kernels, CUDA events) %.3f ms, %.1f neighbours per node, %d TPS points" % (
            len(qs), 1e3 * dt, dev_ms, counts[:, 0].mean(), int(counts[:, 1].sum())))
    pl.close()
    ctx.close()
