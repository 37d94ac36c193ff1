"""Record-level parity of the device-side A* expansion (mppb_expand_nodes = tpobs_grid_kernel + expand_kernel) against the
reference's sequential form of the same work: candidate TPS points in the reference's order, end poses, world-box test,
lattice cells, blocked test, goal-cell rule, best candidate per cell, neighbour pose/twist, interpolated samples and
cost_path_segment (TPS_Astar.cpp:538-826, :269-345; edge_interpolated_path.cpp:51-72; Planner.cpp:15-29).

The checker is tests/hostsim/mppb_sim.cpp's mppb_expand_nodes — std::set / std::map / sequential loops as in the
reference, collision rows and evaluator values from the oracle — loaded here as a plain library (not pre-loaded) and
called on a stand-in context of its own. Bar: every field of every neighbour record bit-identical, same order."""
import ctypes as C

import numpy as np
import pytest

from mrpt_path_planning_b200 import workloads as wl
from test_planner_hostsim_cpu import build_sim

pytestmark = pytest.mark.gpu

WAYPOINTS = ([12.0, 25.0, 31.0, 44.0], [9.0, 30.0, 12.0, 41.0])


def _sim():
    sim = C.CDLL(build_sim())
    sim.mppb_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    sim.mppb_set_obstacles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int]
    sim.mppb_set_ptg_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    sim.mppb_set_ptg_paths.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    sim.mppb_set_expand_params.argtypes = [C.c_void_p, C.c_void_p]
    sim.mppb_set_costmap.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    sim.mppb_set_waypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int]
    sim.mppb_clear_evaluators.argtypes = [C.c_void_p]
    sim.mppb_clear_ptgs.argtypes = [C.c_void_p]
    sim.mppb_expand_max_neighbors.argtypes = [C.c_void_p]
    sim.mppb_expand_max_neighbors.restype = C.c_uint32
    sim.mppb_expand_nodes.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    sim.mppb_expand_last_rows.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    sim.mppbsim_set_oracle_ptgs.argtypes = [C.c_void_p, C.c_int]
    sim.mppbsim_set_oracle_evaluator.argtypes = [C.c_int, C.c_void_p, C.c_int]
    sim.mppb_last_error.argtypes = [C.c_void_p]
    sim.mppb_last_error.restype = C.c_char_p
    return sim


def _nodes(capi, ptgs, rng, n, extent, res_xy, goal_range):
    nodes = np.zeros(n, dtype=capi.EXPAND_NODE_DTYPE)
    for i in range(n):
        x, y, phi = rng.uniform(6, extent - 6), rng.uniform(6, extent - 6), rng.uniform(-np.pi, np.pi)
        ang, d = rng.uniform(-np.pi, np.pi), rng.uniform(*goal_range)
        gx, gy = x + d * np.cos(ang), y + d * np.sin(ang)
        lo = np.minimum([x, y], [gx, gy]) - 4.0
        hi = np.maximum([x, y], [gx, gy]) + 4.0
        nodes["pose"][i] = (x, y, phi)
        nodes["bbox_min"][i] = (lo[0], lo[1], -np.pi)
        nodes["bbox_max"][i] = (hi[0], hi[1], np.pi)
        c, s = np.cos(phi), np.sin(phi)
        rx, ry = (gx - x) * c + (gy - y) * s, -(gx - x) * s + (gy - y) * c
        nodes["goal_k"][i][:] = -1
        nodes["goal_step"][i][:] = 0xFFFFFFFF
        for p, ptg in enumerate(ptgs):
            exact, k, dn = ptg.inverse_map(rx, ry, res_xy)
            if exact:
                nodes["goal_k"][i][p] = k
                found, step = ptg.step_for_dist(k, dn)  # the reference passes the normalised distance (TPS_Astar.cpp:658)
                if found:
                    nodes["goal_step"][i][p] = step
    return nodes


@pytest.mark.parametrize("case", ["demo_params_boxes", "other_params_no_boxes", "no_evaluators_dense"])
def test_expand_nodes_equals_reference_flow(capi, product_ackermann, case):
    from oracle import orc  # the stand-in links liboracle.so: its PTG / evaluator handles must come from that library

    sim = _sim()
    extent = 50.0
    rng = np.random.default_rng({"demo_params_boxes": 1, "other_params_no_boxes": 2, "no_evaluators_dense": 3}[case])
    if case == "no_evaluators_dense":
        xs, ys = wl.uniform_cloud(40000, extent, seed=11)  # dense: most paths blocked, many points inside the robot
    else:
        xs, ys = wl.walls_cloud(30000, extent, wall_fraction=0.9, spacing=0.05, seed=12)
    params = list(wl.DEMO_ASTAR_PARAMS)
    stamps = list(wl.DEMO_ASTAR_TIMESTAMPS)
    if case == "other_params_no_boxes":
        params = [0.30, np.deg2rad(15.0), 0.1, 0.5, 31, 1, 8, 0, 0]
        stamps = [0.25, 0.75, 1.5, 3.0]
    use_boxes = case != "other_params_no_boxes"
    o_ptgs = orc.ackermann_ptgs()
    nodes = _nodes(capi, product_ackermann, rng, 192, extent, params[0], (0.3, 9.0))

    # --- product: real context on the GPU ---
    ctx = capi.Context(0)
    for p in product_ackermann:
        ctx.add_ptg(p)
    ctx.set_obstacles(xs, ys)
    evs = []
    if case != "no_evaluators_dense":
        cm = orc.CostMap(xs[::7], ys[::7], 0.05, 0.5, 4.0, case == "other_params_no_boxes", 0.0, None)
        info = cm.info()
        ctx.set_costmap(0, cm.cells(), info["xmin"], info["ymin"], info["res"], case == "other_params_no_boxes")
        wp = orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 6.0, 0.5, case != "other_params_no_boxes")
        ctx.set_waypoints(1, WAYPOINTS[0], WAYPOINTS[1], 6.0, 0.5, case != "other_params_no_boxes")
        evs = [(cm, 0), (wp, 1)]
    ctx.set_expand_params(params, stamps)
    got, got_counts = ctx.expand_nodes(nodes, 5.0, use_node_boxes=use_boxes)
    got_rows = ctx.expand_last_rows(len(nodes))

    # --- checker: the stand-in's sequential implementation on a context of its own ---
    sctx = C.c_void_p()
    assert sim.mppb_ctx_create(0, None, C.byref(sctx)) == 0
    fx, fy = np.ascontiguousarray(xs, np.float32), np.ascontiguousarray(ys, np.float32)
    assert sim.mppb_set_obstacles(sctx, fx.ctypes.data, fy.ctypes.data, fx.size, 0.0, 0) == 0
    keep = []
    for i, p in enumerate(product_ackermann):
        g = capi.GridDesc()
        assert capi.lib().mppb_ptg_grid_export(p.h, C.byref(g)) == 0
        assert sim.mppb_set_ptg_grid(sctx, i, C.byref(g)) == 0
        d = p.paths_desc()
        keep.append(d)
        assert sim.mppb_set_ptg_paths(sctx, i, C.byref(d)) == 0
    arr = (C.c_void_p * len(o_ptgs))(*[p.h for p in o_ptgs])
    sim.mppbsim_set_oracle_ptgs(arr, len(o_ptgs))
    sim.mppb_clear_evaluators(sctx)
    for slot in range(4):
        sim.mppbsim_set_oracle_evaluator(slot, None, 0)
    for ev, kind in evs:
        if kind == 0:
            sim.mppb_set_costmap(sctx, kind, None)
        else:
            sim.mppb_set_waypoints(sctx, kind, None, None, 0, 1.0, 1.0, 0)
        sim.mppbsim_set_oracle_evaluator(kind, ev.h, kind)
    ts = np.ascontiguousarray(stamps, dtype=np.float64)
    a = capi.AstarParams(params[0], params[1], params[2], params[3], int(params[4]), int(params[5]), int(params[6]),
                         0.0, 0, ts.ctypes.data_as(capi.c_double_p), len(ts))
    assert sim.mppb_set_expand_params(sctx, C.byref(a)) == 0
    stride = sim.mppb_expand_max_neighbors(sctx)
    assert stride == ctx.expand_max_neighbors and stride >= got.shape[1]
    want = np.zeros((len(nodes), stride), dtype=capi.NEIGHBOR_DTYPE)
    want_counts = np.zeros((len(nodes), 2), dtype=np.uint32)
    rc = sim.mppb_expand_nodes(sctx, len(nodes), nodes.ctypes.data, int(use_boxes), 5.0, want.ctypes.data,
                               want_counts.ctypes.data)
    assert rc == 0, sim.mppb_last_error(sctx)
    want_rows = np.empty_like(got_rows)
    sim.mppb_expand_last_rows(sctx, len(nodes), want_rows.ctypes.data)

    # collision rows: the checker's are min(init, row) (values 0 / table float / refDistance); same after the min
    init = np.concatenate([p.init_dists() for p in product_ackermann])
    assert np.array_equal(np.minimum(init[None, :], got_rows.astype(np.float64)), want_rows.astype(np.float64))
    assert np.array_equal(got_counts, want_counts)
    assert got_counts[:, 0].max() > 5 and got_counts[:, 0].sum() > 10 * len(nodes) or case == "no_evaluators_dense"
    n_targets = 0
    for i in range(len(nodes)):
        c = int(got_counts[i, 0])
        for f in capi.NEIGHBOR_DTYPE.names:
            assert np.array_equal(got[i, :c][f], want[i, :c][f]), (case, i, f, got[i, :c][f], want[i, :c][f])
        # (a record whose step is none of the sample steps is a direct-to-goal point)
        n_targets += int(np.sum(~np.isin(got[i, :c]["step"], [int(round(t / 1e-3)) for t in stamps])))
    assert n_targets > 0 or case == "no_evaluators_dense"
    ctx.close()
    sim.mppb_ctx_destroy.argtypes = [C.c_void_p]
    sim.mppb_ctx_destroy(sctx)
