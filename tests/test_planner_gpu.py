"""End-to-end parity of mpp::TPS_Astar: the product planner (host A* + GPU collision/cost batches, through
the C ABI) against the oracle's restatement of the reference planner on the same inputs.

Bar (north_star): free/blocked decisions and chosen k per edge bit-exact, final path identical. Here the whole
motion tree (node ids, parents, poses, costs) is compared for exact equality, which implies both."""
import os

import numpy as np
import pytest

from mrpt_path_planning_b200 import workloads as wl
from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS, c1_inputs

pytestmark = pytest.mark.gpu

WAYPOINTS = ([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0])  # share/mvsim-demo-waypoints01.yaml


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


def same_plan(prod, ora, r_p, r_o):
    for key in ("success", "tree_nodes", "tree_parents", "expanded", "tps_points", "edges_costed", "best_node",
                "goal_node"):
        assert r_p[key] == r_o[key], key
    assert r_p["path_cost"] == r_o["path_cost"]
    assert r_p["best_cost_to_goal"] == r_o["best_cost_to_goal"]
    assert np.array_equal(prod.tree(), ora.tree())
    n_p, e_p = prod.path()
    n_o, e_o = ora.path()
    assert np.array_equal(n_p, n_o) and np.array_equal(e_p, e_o)


def test_c1_holonomic_demo_identical_tree_and_path(ctx, capi, orc):
    """config C1: path-planner-cli holonomic PTG, obstacles_01.txt, start [0.5 0 0] goal [4 2.5 45], costmap"""
    xs, ys, start, goal, lo, hi = c1_inputs()
    ptgs = capi.holonomic_ptgs()
    prod = capi.Planner(ctx, ptgs, C1_PARAMS, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    prod.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    ora = orc.Planner(orc.holonomic_ptgs(), C1_PARAMS, C1_TIMESTAMPS)
    ora.add_obstacles(xs, ys)
    cm = orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    ora.add_costmap(cm)
    # the GPU-built costmap equals the reference's
    info, cells = prod.costmap_cells(0)
    assert info == cm.info() and np.array_equal(cells, cm.cells())
    r_p = prod.plan(start, goal, lo, hi)
    r_o = ora.plan(start, goal, lo, hi)
    assert r_p["success"] == 1.0
    same_plan(prod, ora, r_p, r_o)
    # the CLI's post-processing: refine_trajectory + plan_to_trajectory + save_to_txt
    prod.refine()
    ora.refine()
    n_p, e_p = prod.path()
    n_o, e_o = ora.path()
    assert np.array_equal(n_p, n_o) and np.array_equal(e_p, e_o)
    assert prod.trajectory_txt() == ora.trajectory_txt()
    st = prod.gpu_stats()
    assert st["collision_batches"] == r_p["expanded"] and st["holo_candidates"] > 0 and st["edges"] == r_p["edges_costed"]
    # decision guard: no free/blocked decision of the search was taken within 1e-9 relative of a tie, so the device's
    # acos/cos (the one non-IEEE ingredient of the HolonomicBlend distances) cannot have tipped one
    assert st["holo_near_ties"] == 0
    # planning again on the same planner object (resident cloud and tables) gives the same answer
    r_p2 = prod.plan(start, goal, lo, hi)
    assert r_p2["path_cost"] == r_p["path_cost"] and r_p2["tree_nodes"] == r_p["tree_nodes"]


def test_c2_ackermann_costmap_waypoints_identical(ctx, capi, orc, product_ackermann, oracle_ackermann):
    """config C2: Ackermann (DiffDrive_C) PTGs with costmap-obstacles + prefer-waypoints"""
    xs, ys, start, goal, lo, hi = c1_inputs()
    prod = capi.Planner(ctx, product_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    prod.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    prod.add_waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, True)
    ora = orc.Planner(oracle_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    ora.add_obstacles(xs, ys)
    ora.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
    ora.add_waypoints(orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, True))
    r_p = prod.plan(start, goal, lo, hi)
    r_o = ora.plan(start, goal, lo, hi)
    same_plan(prod, ora, r_p, r_o)
    assert r_p["expanded"] > 5
    prod.refine()
    ora.refine()
    assert np.array_equal(prod.path()[1], ora.path()[1])
    if r_p["success"] == 1.0:
        assert prod.trajectory_txt() == ora.trajectory_txt()


def test_point_goal_and_expansion_cap(ctx, capi, orc):
    xs, ys, start, goal, lo, hi = c1_inputs()
    params = list(C1_PARAMS)
    params[8] = 12  # stop after 12 expansions: unfinished plan, best node so far
    prod = capi.Planner(ctx, capi.holonomic_ptgs(), params, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    ora = orc.Planner(orc.holonomic_ptgs(), params, C1_TIMESTAMPS)
    ora.add_obstacles(xs, ys)
    r_p = prod.plan(start, goal, lo, hi)
    r_o = ora.plan(start, goal, lo, hi)
    assert r_p["expanded"] == 12 and r_p["success"] == 0.0
    same_plan(prod, ora, r_p, r_o)
    # R(2) point goal
    prod2 = capi.Planner(ctx, capi.holonomic_ptgs(), C1_PARAMS, C1_TIMESTAMPS)
    prod2.add_obstacles(xs, ys)
    ora2 = orc.Planner(orc.holonomic_ptgs(), C1_PARAMS, C1_TIMESTAMPS)
    ora2.add_obstacles(xs, ys)
    r_p = prod2.plan(start, goal, lo, hi, goal_is_point=True)
    r_o = ora2.plan(start, goal, lo, hi, goal_is_point=True)
    assert r_p["success"] == 1.0
    same_plan(prod2, ora2, r_p, r_o)


def _random_queries(rng, n, extent, dmin, dmax, margin=4.0):
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


@pytest.mark.parametrize("kind", ["ackermann", "holonomic"])
def test_batch_of_queries_equals_single_plans_and_oracle(ctx, capi, orc, kind, product_ackermann, oracle_ackermann):
    """config C4 in small: independent queries on one shared cloud, lock-step batch == one by one == oracle"""
    extent = 40.0
    xs, ys = wl.walls_cloud(6000, extent, wall_fraction=0.8, seed=33)
    qs = _random_queries(np.random.default_rng(34), 6, extent, 3.0, 8.0)
    params = list(C1_PARAMS)
    params[8] = 40  # cap the work per query
    p_ptgs = product_ackermann if kind == "ackermann" else capi.holonomic_ptgs()
    o_ptgs = oracle_ackermann if kind == "ackermann" else orc.holonomic_ptgs()
    prod = capi.Planner(ctx, p_ptgs, params, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    prod.add_waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True)
    batch = prod.plan_many(qs, n_threads=4)
    trees = [prod.tree(i) for i in range(len(qs))]
    paths = [prod.path(i) for i in range(len(qs))]
    assert sum(r["expanded"] for r in batch) > 50
    for i, q in enumerate(qs):
        single = capi.Planner(ctx, p_ptgs, params, C1_TIMESTAMPS)
        single.add_obstacles(xs, ys)
        single.add_waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True)
        r1 = single.plan(*q[:4])
        for key in ("success", "path_cost", "tree_nodes", "expanded", "edges_costed", "best_node"):
            assert r1[key] == batch[i][key], (i, key)
        assert np.array_equal(single.tree(), trees[i])
        if i < 3:
            ora = orc.Planner(o_ptgs, params, C1_TIMESTAMPS)
            ora.add_obstacles(xs, ys)
            ora.add_waypoints(orc.Waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True))
            r_o = ora.plan(*q[:4])
            assert r_o["path_cost"] == batch[i]["path_cost"] and r_o["expanded"] == batch[i]["expanded"]
            assert np.array_equal(ora.tree(), trees[i])
            n_o, e_o = ora.path()
            assert np.array_equal(n_o, paths[i][0]) and np.array_equal(e_o, paths[i][1])


def test_boxed_eval_equals_clipped_upload(ctx, capi, product_ackermann):
    """per-node world boxes on a shared cloud == uploading the cloud clipped to that box"""
    c = capi.Context(0)
    for p in product_ackermann:
        c.add_ptg(p)
    xs, ys = wl.uniform_cloud(30000, 40.0, seed=16)
    poses = wl.node_poses(32, 40.0, margin=8.0, seed=17)
    boxes = np.stack([poses[:, 0] - 3.5, poses[:, 1] - 2.0, poses[:, 0] + 4.0, poses[:, 1] + 6.0], 1).astype(np.float32)
    c.set_obstacles(xs, ys)
    got = c.eval_nodes_boxed(poses, boxes, 5.0)
    for i in range(len(poses)):
        c.set_obstacles_clipped(xs, ys, boxes[i])
        ref = c.eval_nodes(poses[i:i + 1], 5.0)
        assert np.array_equal(got[i], ref[0]), i
    c.close()


def test_planner_input_errors(ctx, capi):
    xs, ys, start, goal, lo, hi = c1_inputs()
    pl = capi.Planner(ctx, capi.holonomic_ptgs(), C1_PARAMS, C1_TIMESTAMPS)
    pl.add_obstacles(xs, ys)
    with pytest.raises(capi.MppbError):  # start outside the world box (ASSERT_ in TPS_Astar.cpp:96)
        pl.plan([100.0, 0, 0, 0, 0, 0], goal, lo, hi)
    with pytest.raises(capi.MppbError):  # empty world box
        pl.plan(start, goal, lo, lo)


@pytest.mark.parametrize("kind", ["holonomic", "ackermann"])
def test_successive_plans_on_one_planner(ctx, capi, orc, kind, product_ackermann):
    """Replanning with ONE planner object, as NavEngine does: the PTG objects carry their state (dynamic state, trimmed
    speed, HolonomicBlend's step-count cache) from plan to plan, the cloud and the evaluators stay resident on the GPU
    between plans, and an evaluator added later must be picked up. Every plan equals the checker's, which is driven
    through the same sequence on one planner object of its own."""
    xs, ys, start, goal, lo, hi = c1_inputs()
    p_ptgs = capi.holonomic_ptgs() if kind == "holonomic" else product_ackermann
    o_ptgs = orc.holonomic_ptgs() if kind == "holonomic" else orc.ackermann_ptgs()
    prod = capi.Planner(ctx, p_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    ora = orc.Planner(o_ptgs, C1_PARAMS, C1_TIMESTAMPS)
    for pl in (prod, ora):
        pl.add_obstacles(xs, ys)
    prod.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    ora.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
    goals = [goal, [3.0, -0.5, -1.0, 0, 0, 0], goal, [1.5, 3.0, 2.0, 0, 0, 0]]
    starts = [start, [4.0, 2.5, 0.8, 0.2, 0.1, 0.0], start, [3.0, -0.5, -1.0, 0, 0, 0]]
    for i, (s, g) in enumerate(zip(starts, goals)):
        if i == 2:  # a second evaluator appears between plans
            prod.add_waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, True)
            ora.add_waypoints(orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, True))
        r_p = prod.plan(s, g, lo, hi)
        r_o = ora.plan(s, g, lo, hi)
        same_plan(prod, ora, r_p, r_o)
        assert r_p["expanded"] > 0


PARAM_SETS = {
    "defaults_of_the_class": ([0.20, np.deg2rad(5.0), 1.0, 0.1, 20, 3, 5, 0, 30], [1.0, 3.0, 5.0]),
    "coarse_two_speeds": ([0.50, np.deg2rad(45.0), 0.3, 0.0, 7, 2, 3, 0, 60], [0.4, 2.0]),
    "one_speed_many_paths": ([0.30, np.deg2rad(15.0), 0.1, 0.5, 31, 1, 8, 0, 40], [0.25, 0.75, 1.5, 3.0]),
}


@pytest.mark.parametrize("kind", ["holonomic", "ackermann"])
@pytest.mark.parametrize("pname", sorted(PARAM_SETS))
def test_other_planner_parameters(ctx, capi, orc, kind, pname, product_ackermann):
    """Other TPS_Astar parameter sets than the demo's (lattice resolution, trajectories, speeds, sample timestamps,
    interpolated segments, heuristic weights; averaged costmap, max-combined waypoints): tree and path identical."""
    params, stamps = PARAM_SETS[pname]
    xs, ys, start, goal, lo, hi = c1_inputs()
    p_ptgs = capi.holonomic_ptgs() if kind == "holonomic" else product_ackermann
    o_ptgs = orc.holonomic_ptgs() if kind == "holonomic" else orc.ackermann_ptgs()
    prod = capi.Planner(ctx, p_ptgs, params, stamps)
    ora = orc.Planner(o_ptgs, params, stamps)
    for pl in (prod, ora):
        pl.add_obstacles(xs, ys)
    prod.add_costmap(xs, ys, 0.05, 0.4, 2.0, True, 0.0, None)
    ora.add_costmap(orc.CostMap(xs, ys, 0.05, 0.4, 2.0, True, 0.0, None))
    prod.add_waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, False)
    ora.add_waypoints(orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.5, 0.5, False))
    r_p = prod.plan(start, goal, lo, hi)
    r_o = ora.plan(start, goal, lo, hi)
    same_plan(prod, ora, r_p, r_o)
    assert r_p["expanded"] >= 5


@pytest.mark.parametrize("kind", ["ackermann", "holonomic"])
def test_batch_without_evaluators_equals_single_plans_and_oracle(ctx, capi, orc, kind, product_ackermann, oracle_ackermann):
    """No cost evaluator on the GPU (the shape of config C4: no cost batch at all, relaxation straight after C/D).
    Batch == one by one == oracle; queries that finish early must not disturb the others."""
    extent = 40.0
    xs, ys = wl.walls_cloud(6000, extent, wall_fraction=0.8, seed=35)
    qs = _random_queries(np.random.default_rng(36), 7, extent, 2.0, 9.0)
    params = list(C1_PARAMS)
    params[8] = 45
    p_ptgs = product_ackermann if kind == "ackermann" else capi.holonomic_ptgs()
    o_ptgs = oracle_ackermann if kind == "ackermann" else orc.holonomic_ptgs()
    prod = capi.Planner(ctx, p_ptgs, params, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    batch = prod.plan_many(qs, n_threads=3)
    trees = [prod.tree(i) for i in range(len(qs))]
    assert sum(r["expanded"] for r in batch) > 50
    for i, q in enumerate(qs):
        single = capi.Planner(ctx, p_ptgs, params, C1_TIMESTAMPS)
        single.add_obstacles(xs, ys)
        r1 = single.plan(*q[:4])
        for key in ("success", "path_cost", "tree_nodes", "expanded", "edges_costed", "best_node"):
            assert r1[key] == batch[i][key], (i, key)
        assert np.array_equal(single.tree(), trees[i])
        if i < 4:
            ora = orc.Planner(o_ptgs, params, C1_TIMESTAMPS)
            ora.add_obstacles(xs, ys)
            r_o = ora.plan(*q[:4])
            assert r_o["path_cost"] == batch[i]["path_cost"] and r_o["expanded"] == batch[i]["expanded"]
            assert np.array_equal(ora.tree(), trees[i])


def test_obstacles_changed_between_plans_and_interleaved_planners(ctx, capi, orc, product_ackermann, oracle_ackermann):
    """The reference re-reads ObstacleSource::obstacles() on every plan() (TPS_Astar.cpp:125-137). The product keeps
    the cloud resident on the GPU; it must notice (a) a cloud updated in place — same point count, same buffers —
    and (b) another planner on the same context that replaced the resident cloud in between."""
    xs, ys, start, goal, lo, hi = c1_inputs()
    rng = np.random.default_rng(5)
    xs2 = (xs + rng.uniform(-0.3, 0.3, xs.size)).astype(np.float32)
    ys2 = (ys + rng.uniform(-0.3, 0.3, ys.size)).astype(np.float32)
    pa = capi.Planner(ctx, product_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    pb = capi.Planner(ctx, product_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    pa.add_obstacles(xs, ys)
    pb.add_obstacles(xs2, ys2)
    oa = orc.Planner(oracle_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    ob = orc.Planner(oracle_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    oa.add_obstacles(xs, ys)
    ob.add_obstacles(xs2, ys2)
    ra, rb = oa.plan(start, goal, lo, hi), ob.plan(start, goal, lo, hi)
    assert not np.array_equal(oa.tree(), ob.tree())
    for _ in range(2):
        same_plan(pa, oa, pa.plan(start, goal, lo, hi), ra)
        same_plan(pb, ob, pb.plan(start, goal, lo, hi), rb)
    pa.update_obstacles(0, xs2, ys2)
    same_plan(pa, ob, pa.plan(start, goal, lo, hi), rb)
    pa.close()
    pb.close()


def test_interleaved_planners_with_different_ptg_tables_on_one_context(ctx, capi, orc, product_ackermann, oracle_ackermann):
    """Two planners on one context whose PTG sets have the same count but different tables: each plan must run on
    its own planner's tables (a PTG epoch on the context tells a planner that somebody else replaced them)."""
    xs, ys, start, goal, lo, hi = c1_inputs()
    p_small, o_small = capi.ackermann_ptgs(num_paths=61, ref_distance=4.0), orc.ackermann_ptgs(num_paths=61, ref_distance=4.0)
    pa = capi.Planner(ctx, product_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    pb = capi.Planner(ctx, p_small, C1_PARAMS, C1_TIMESTAMPS)
    oa = orc.Planner(oracle_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    ob = orc.Planner(o_small, C1_PARAMS, C1_TIMESTAMPS)
    for pl in (pa, pb, oa, ob):
        pl.add_obstacles(xs, ys)
    ra, rb = oa.plan(start, goal, lo, hi), ob.plan(start, goal, lo, hi)
    for _ in range(2):
        same_plan(pa, oa, pa.plan(start, goal, lo, hi), ra)
        same_plan(pb, ob, pb.plan(start, goal, lo, hi), rb)
    pa.close()
    pb.close()


def test_more_than_max_evaluators_and_direct_plugin_calls_leave_the_planner_intact(ctx, capi, orc, product_ackermann,
                                                                                   oracle_ackermann):
    """Evaluators beyond MPPB_MAX_EVALUATORS run as host plugins through operator(), which uses a private context:
    the planner's resident slots stay as they are and the costs are the reference's."""
    xs, ys, start, goal, lo, hi = c1_inputs()
    prod = capi.Planner(ctx, product_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    ora = orc.Planner(oracle_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    prod.add_obstacles(xs, ys)
    ora.add_obstacles(xs, ys)
    cm = orc.CostMap(xs, ys, 0.05, 0.4, 1.0, False, 0.0, None)
    for i in range(6):  # 4 GPU slots + 2 host-plugin evaluators
        if i % 2 == 0:
            prod.add_costmap(xs, ys, 0.05, 0.4, 1.0, False, 0.0, None)
            ora.add_costmap(cm)
        else:
            prod.add_waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.0 + 0.25 * i, 0.5, True)
            ora.add_waypoints(orc.Waypoints(WAYPOINTS[0], WAYPOINTS[1], 1.0 + 0.25 * i, 0.5, True))
    r_p, r_o = prod.plan(start, goal, lo, hi), ora.plan(start, goal, lo, hi)
    same_plan(prod, ora, r_p, r_o)
    prod.close()
