#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Runs the product planner's HOST logic without a GPU: started by tests/test_planner_hostsim_cpu.py
with LD_PRELOAD=tests/_build/libmppb_sim.so, so that the device entry points libmpp_b200's host code calls are answered
by the oracle (tests/hostsim/mppb_sim.cpp). Every plan of the product planner must equal the oracle planner's: same
counts, same motion tree, same path. Prints one JSON line; any assertion fails the process."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mrpt_path_planning_b200 import capi, workloads as wl  # noqa: E402
from oracle import orc  # noqa: E402
from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS, c1_inputs  # noqa: E402

WAYPOINTS = ([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0])
sim = C.CDLL(os.path.join(ROOT, "tests", "_build", "libmppb_sim.so"))
assert "libmppb_sim" in os.environ.get("LD_PRELOAD", ""), "start me with LD_PRELOAD=libmppb_sim.so"


class SimContext:
    """what capi.Planner needs of a context: the handle (made by the stand-in, not by libmpp_b200)"""

    def __init__(self):
        h = C.c_void_p()
        sim.mppb_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        assert sim.mppb_ctx_create(0, None, C.byref(h)) == 0
        self.h = h


def register_ptgs(o_ptgs):
    arr = (C.c_void_p * len(o_ptgs))(*[p.h for p in o_ptgs])
    sim.mppbsim_set_oracle_ptgs(arr, len(o_ptgs))


def register_evaluator(slot, handle, kind):
    sim.mppbsim_set_oracle_evaluator.argtypes = [C.c_int, C.c_void_p, C.c_int]
    sim.mppbsim_set_oracle_evaluator(slot, handle.h if handle is not None else None, kind)


def register_holo(product_ptg, oracle_ptg, n_speeds=3):
    """table (vxf, vyf, vf) -> (k, trimmed speed) from the product's host PTG (its |V|, |omega| formulas read dir and
    the trimmed speed only), and the oracle HolonomicBlend object the stand-in evaluates with"""
    inc = 1.0 / n_speeds
    speeds, sp = [], inc
    while sp < 1.001:  # TPS_Astar.cpp:636-640
        speeds.append(sp)
        sp += inc
    rows = []
    product_ptg.update_dyn_state((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), 0.0, force=True)
    for sp in speeds:
        product_ptg.set_trim_speed(sp)
        for k in range(product_ptg.num_paths):
            c = product_ptg.holo_params(k)
            rows.append((c.vxf, c.vyf, c.vf, k, sp))
    a = np.array(rows)
    vxf, vyf, vf = (np.ascontiguousarray(a[:, i]) for i in range(3))
    ks = np.ascontiguousarray(a[:, 3].astype(np.int32))
    sps = np.ascontiguousarray(a[:, 4])
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sim.mppbsim_set_holo.argtypes = [C.c_void_p, C.c_size_t, dp, dp, dp, ip, dp]
    sim.mppbsim_set_holo(oracle_ptg.h, len(rows), vxf.ctypes.data_as(dp), vyf.ctypes.data_as(dp), vf.ctypes.data_as(dp),
                         ks.ctypes.data_as(ip), sps.ctypes.data_as(dp))
    return (vxf, vyf, vf, ks, sps)  # keep alive until copied (the stand-in copies at once)


def call_counts():
    v = (C.c_uint64 * 4)()
    sim.mppbsim_call_counts(v)
    return dict(eval_nodes=v[0], eval_nodes_begin=v[1], eval_edge_costs=v[2], sync=v[3])


# the product planner uses the device-side expansion for collision-grid PTG sets unless MPPB_NO_EXPAND is set; the
# test runs this script once each way, so that both the slim host path and the full host phases are compared
EXPAND = os.environ.get("MPPB_NO_EXPAND") is None


def expand_counts(ctx):
    v = (C.c_uint64 * 2)()
    sim.mppbsim_expand_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    sim.mppbsim_expand_counts(ctx.h, v)
    return (v[0], v[1])


def same_plan(prod, ora, r_p, r_o, query=0):
    for key in ("success", "tree_nodes", "tree_parents", "expanded", "tps_points", "edges_costed", "best_node",
                "goal_node"):
        assert r_p[key] == r_o[key], (key, r_p[key], r_o[key])
    assert r_p["path_cost"] == r_o["path_cost"]
    assert r_p["best_cost_to_goal"] == r_o["best_cost_to_goal"]
    assert np.array_equal(prod.tree(query), ora.tree())


def random_queries(rng, n, extent, dmin, dmax, margin=4.0):
    qs = []
    while len(qs) < n:
        s = rng.uniform(8, extent - 8, 2)
        g = rng.uniform(8, extent - 8, 2)
        d = np.hypot(*(g - s))
        if not (dmin <= d <= dmax):
            continue
        lo = np.minimum(s, g) - margin
        hi = np.maximum(s, g) + margin
        qs.append(([s[0], s[1], rng.uniform(-3, 3), 0, 0, 0], [g[0], g[1], rng.uniform(-3, 3), 0, 0, 0],
                   [lo[0], lo[1], -np.pi], [hi[0], hi[1], np.pi], False))
    return qs


def main():
    out = {}
    # batches of two or more queries alternate between two waves (normally from 64 queries on)
    os.environ["MPPB_WAVES_MIN_QUERIES"] = "2"
    # must fail loudly without the stand-in's context: the real library has no CPU path
    try:
        capi.Context(0)
        real_ctx = True
    except capi.MppbError:
        real_ctx = False
    out["real_context_available"] = real_ctx

    xs, ys, start, goal, lo, hi = c1_inputs()
    p_ptgs = capi.ackermann_ptgs()  # host-built tables
    o_ptgs = orc.ackermann_ptgs()
    register_ptgs(o_ptgs)
    ctx = SimContext()

    # --- C2: Ackermann PTGs, costmap + preferred waypoints (single query: split calls, deferred tree insertions) ---
    cm = orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    wp = orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, True)
    register_evaluator(0, cm, 0)
    register_evaluator(1, wp, 1)
    prod = capi.Planner(ctx, p_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    prod.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    prod.add_waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, True)
    ora = orc.Planner(o_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    ora.add_obstacles(xs, ys)
    ora.add_costmap(cm)
    ora.add_waypoints(wp)
    r_p, r_o = prod.plan(start, goal, lo, hi), ora.plan(start, goal, lo, hi)
    same_plan(prod, ora, r_p, r_o)
    assert r_p["expanded"] > 5
    n_p, e_p = prod.path()
    n_o, e_o = ora.path()
    assert np.array_equal(n_p, n_o) and np.array_equal(e_p, e_o)
    out["c2"] = dict(expanded=int(r_p["expanded"]), tree_nodes=int(r_p["tree_nodes"]), success=bool(r_p["success"]))
    cc = call_counts()
    if EXPAND:  # device-side expansion: one expansion call (and one sync) per A* iteration, no separate cost batches
        assert cc["eval_nodes_begin"] == 0 and cc["eval_edge_costs"] == 0 and cc["sync"] >= r_p["expanded"], cc
        assert expand_counts(ctx) == (r_p["expanded"], r_p["expanded"])
    else:
        assert cc["eval_nodes_begin"] == r_p["expanded"] and cc["sync"] >= r_p["expanded"], cc  # split form used
        assert expand_counts(ctx) == (0, 0)
    # replanning on the same objects, other start/goal (PTG state carried over on both sides)
    for s, g in (([4.0, 2.5, 0.8, 0.2, 0.1, 0.0], [3.0, -0.5, -1.0, 0, 0, 0]), (start, [1.5, 3.0, 2.0, 0, 0, 0])):
        r_p, r_o = prod.plan(s, g, lo, hi), ora.plan(s, g, lo, hi)
        same_plan(prod, ora, r_p, r_o)
    prod.close()

    # --- obstacles changed between plans: in place (same count, same buffers) and by another planner on the context ---
    for slot in (0, 1):
        register_evaluator(slot, None, 0)
    rng = np.random.default_rng(5)
    xs2 = (xs + rng.uniform(-0.3, 0.3, xs.size)).astype(np.float32)
    ys2 = (ys + rng.uniform(-0.3, 0.3, ys.size)).astype(np.float32)
    pa = capi.Planner(ctx, p_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    pb = capi.Planner(ctx, p_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    pa.add_obstacles(xs, ys)
    pb.add_obstacles(xs2, ys2)
    oa = orc.Planner(o_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    ob = orc.Planner(o_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    oa.add_obstacles(xs, ys)
    ob.add_obstacles(xs2, ys2)
    ra, rb = oa.plan(start, goal, lo, hi), ob.plan(start, goal, lo, hi)
    assert not np.array_equal(oa.tree(), ob.tree())  # the two clouds do give different searches
    for _ in range(2):  # interleaved on ONE context: each planner must notice that the other replaced the cloud
        same_plan(pa, oa, pa.plan(start, goal, lo, hi), ra)
        same_plan(pb, ob, pb.plan(start, goal, lo, hi), rb)
    pa.update_obstacles(0, xs2, ys2)  # same point count: the cloud's buffers keep their addresses
    same_plan(pa, ob, pa.plan(start, goal, lo, hi), rb)
    pa.close()
    pb.close()
    out["obstacles_changed_between_plans"] = True

    # --- no evaluators, expansion cap and point goal ---
    for slot in (0, 1):
        register_evaluator(slot, None, 0)
    for params, point in ((list(C1_PARAMS[:8]) + [7], False), (list(C1_PARAMS), True)):
        prod = capi.Planner(ctx, p_ptgs, params, C1_TIMESTAMPS)
        prod.add_obstacles(xs, ys)
        ora = orc.Planner(o_ptgs, params, C1_TIMESTAMPS)
        ora.add_obstacles(xs, ys)
        r_p = prod.plan(start, goal, lo, hi, goal_is_point=point)
        r_o = ora.plan(start, goal, lo, hi, goal_is_point=point)
        same_plan(prod, ora, r_p, r_o)
        prod.close()
    out["cap_and_point_goal"] = True

    # --- a batch of independent queries in lock-step (worker pool, per-node world boxes) == the oracle one by one ---
    extent = 40.0
    mx, my = wl.walls_cloud(6000, extent, wall_fraction=0.8, seed=33)
    qs = random_queries(np.random.default_rng(34), 6, extent, 3.0, 8.0)
    params = list(C1_PARAMS)
    params[8] = 30
    wp2 = orc.Waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True)
    register_evaluator(0, wp2, 1)
    prod = capi.Planner(ctx, p_ptgs, params, C1_TIMESTAMPS)
    prod.add_obstacles(mx, my)
    prod.add_waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True)
    batch = prod.plan_many(qs, n_threads=3)
    total = 0
    for i, q in enumerate(qs):
        ora = orc.Planner(o_ptgs, params, C1_TIMESTAMPS)
        ora.add_obstacles(mx, my)
        ora.add_waypoints(wp2)
        r_o = ora.plan(*q[:4])
        same_plan(prod, ora, batch[i], r_o, query=i)
        total += int(r_o["expanded"])
    assert total > 50
    out["batch"] = dict(queries=len(qs), expansions=total)
    prod.close()
    # --- C1: HolonomicBlend PTG with the obstacle costmap (candidate records, direction memo, split calls) ---
    ph = capi.holonomic_ptgs()
    oh = orc.holonomic_ptgs()
    sim_holo = orc.holonomic_ptgs()[0]  # the stand-in's own oracle object (the oracle planner keeps its own state)
    register_ptgs([])
    register_holo(capi.holonomic_ptgs()[0], sim_holo)
    register_evaluator(0, cm, 0)
    register_evaluator(1, None, 0)
    prod = capi.Planner(ctx, ph, C1_PARAMS, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    prod.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    ora = orc.Planner(oh, C1_PARAMS, C1_TIMESTAMPS)
    ora.add_obstacles(xs, ys)
    ora.add_costmap(cm)
    r_p, r_o = prod.plan(start, goal, lo, hi), ora.plan(start, goal, lo, hi)
    same_plan(prod, ora, r_p, r_o)
    assert r_p["success"] == 1.0 and r_p["expanded"] > 20
    out["c1"] = dict(expanded=int(r_p["expanded"]), tree_nodes=int(r_p["tree_nodes"]))
    for s_, g_ in (([4.0, 2.5, 0.8, 0.2, 0.1, 0.0], [3.0, -0.5, -1.0, 0, 0, 0]), (start, [1.5, 3.0, 2.0, 0, 0, 0])):
        r_p, r_o = prod.plan(s_, g_, lo, hi), ora.plan(s_, g_, lo, hi)
        same_plan(prod, ora, r_p, r_o)
    prod.close()
    # a holonomic batch in lock-step (per-node boxes, private PTG clones), no evaluators
    register_evaluator(0, None, 0)
    params = list(C1_PARAMS)
    params[8] = 15
    prod = capi.Planner(ctx, ph, params, C1_TIMESTAMPS)
    prod.add_obstacles(mx, my)
    batch = prod.plan_many(qs[:4], n_threads=2)
    for i, q in enumerate(qs[:4]):
        ora = orc.Planner(orc.holonomic_ptgs(), params, C1_TIMESTAMPS)
        ora.add_obstacles(mx, my)
        same_plan(prod, ora, batch[i], ora.plan(*q[:4]), query=i)
    out["holonomic_batch"] = True
    prod.close()
    # --- other TPS_Astar parameter sets (lattice resolution, trajectories, speeds, sample timestamps, interpolated
    #     segments, heuristic weights; averaged costmap, max-combined waypoints), both PTG kinds ---
    param_sets = {
        "defaults_of_the_class": ([0.20, np.deg2rad(5.0), 1.0, 0.1, 20, 3, 5, 0, 30], [1.0, 3.0, 5.0]),
        "coarse_two_speeds": ([0.50, np.deg2rad(45.0), 0.3, 0.0, 7, 2, 3, 0, 60], [0.4, 2.0]),
        "one_speed_many_paths": ([0.30, np.deg2rad(15.0), 0.1, 0.5, 31, 1, 8, 0, 40], [0.25, 0.75, 1.5, 3.0]),
    }
    cm2 = orc.CostMap(xs, ys, 0.05, 0.4, 2.0, True, 0.0, None)
    wp3 = orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, False)
    register_evaluator(0, cm2, 0)
    register_evaluator(1, wp3, 1)
    done = []
    for pname, (params, stamps) in sorted(param_sets.items()):
        for kind in ("ackermann", "holonomic"):
            if kind == "ackermann":
                register_ptgs(o_ptgs)
                pp, oo = p_ptgs, o_ptgs
            else:
                register_ptgs([])
                register_holo(capi.holonomic_ptgs()[0], sim_holo, n_speeds=int(params[5]))
                pp, oo = capi.holonomic_ptgs(), orc.holonomic_ptgs()
            prod = capi.Planner(ctx, pp, params, stamps)
            ora = orc.Planner(oo, params, stamps)
            for pl in (prod, ora):
                pl.add_obstacles(xs, ys)
            prod.add_costmap(xs, ys, 0.05, 0.4, 2.0, True, 0.0, None)
            ora.add_costmap(cm2)
            prod.add_waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, False)
            ora.add_waypoints(wp3)
            r_p, r_o = prod.plan(start, goal, lo, hi), ora.plan(start, goal, lo, hi)
            same_plan(prod, ora, r_p, r_o)
            assert r_p["expanded"] >= 5
            done.append(pname + "/" + kind)
            prod.close()
    out["parameter_sets"] = done
    out["expand"] = EXPAND
    print(json.dumps(out))


if __name__ == "__main__":
    main()
