"""Pins the oracle restatement (oracle/orc_*.hpp, "port") against the REFERENCE'S OWN CODE: the
hot-path translation units of /root/reference compiled unmodified into oracle/_ref/libmpp_ref.so
(oracle/Makefile target `ref`; MRPT replaced by the stand-in oracle/ref/mrpt_shim.h).

Every comparison is exact (==): both sides are x86-64 code built with the reference's flags from
the same arithmetic, so any difference is a restatement error. CPU-only; runs wherever the prebuilt
library exists (it is built here from /root/reference and shipped to the GPU box)."""
import os

import numpy as np
import pytest

from mrpt_path_planning_b200 import workloads as wl

SHARE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "share")
WAYPOINTS01 = ([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0])


@pytest.fixture(scope="module")
def port():
    from oracle import orc

    orc.lib()
    return orc


@pytest.fixture(scope="module")
def refm():
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref/libmpp_ref.so not built and /root/reference absent")
    return ref.module()


@pytest.fixture(scope="module")
def ack(port, refm):
    return port.ackermann_ptgs(), refm.ackermann_ptgs()


@pytest.fixture(scope="module")
def holo(port, refm):
    return port.holonomic_ptgs()[0], refm.holonomic_ptgs()[0]


def test_reference_library_is_the_reference_build(refm):
    assert refm.KIND == "reference" and refm.lib().orc_is_reference_build() == 1


def test_diffdrive_tables_identical(ack):
    """Trajectories (DiffDriveCollisionGridBased.cpp:103-201) and collision grid (:615-764), all k,
    including the per-cell entry order."""
    for a, b in zip(*ack):
        assert a.num_paths == b.num_paths == 121
        assert a.ref_distance == b.ref_distance and a.max_radius == b.max_radius
        assert a.step_duration == b.step_duration
        for k in range(a.num_paths):
            assert a.step_count(k) == b.step_count(k)
            assert np.array_equal(a.path(k), b.path(k)), k
            assert a.init_dist(k) == b.init_dist(k)
        ga, gb = a.grid(), b.grid()
        for key in ("xmin", "ymin", "res", "nx", "ny"):
            assert ga[key] == gb[key]
        for key in ("offsets", "ks", "ds"):
            assert np.array_equal(ga[key], gb[key]), key


def test_ini_loader_gives_the_same_ptgs(refm, ack, holo):
    """TrajectoriesAndRobotShape::initFromConfigFile (the reference's loader, on the reference's own
    share/*.ini) yields the PTGs that the numeric constructors used everywhere else produce."""
    from oracle import ref

    ini = ref.ptgs_from_ini(os.path.join(SHARE, "ptgs_ackermann_vehicle.ini"))
    assert len(ini) == 2
    for a, b in zip(ack[1], ini):
        ga, gb = a.grid(), b.grid()
        for key in ("offsets", "ks", "ds"):
            assert np.array_equal(ga[key], gb[key])
        assert a.max_radius == b.max_radius
        for k in (0, 30, 60, 120):
            assert np.array_equal(a.path(k), b.path(k))
    (h,) = ref.ptgs_from_ini(os.path.join(SHARE, "ptgs_holonomic_robot.ini"))
    assert h.num_paths == 191 and h.ref_distance == 5.0 and h.max_radius == 0.15
    for k in (0, 47, 95, 190):
        for step in (1, 50, 150, 400):
            assert np.array_equal(h.path_pose(k, step), holo[1].path_pose(k, step))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_clip_and_transform_identical(port, refm, seed):
    rng = np.random.default_rng(seed)
    xs = rng.uniform(-20, 20, 5000).astype(np.float32)
    ys = rng.uniform(-20, 20, 5000).astype(np.float32)
    pose = [rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-np.pi, np.pi)]
    # points exactly on / next to the clip border
    xs[:4] = np.float32(pose[0] + 5.0), np.nextafter(np.float32(pose[0] + 5.0), np.float32(99)), np.float32(pose[0] - 5.0), 0
    a, b = port.local_obstacles(xs, ys, pose, 5.0), refm.local_obstacles(xs, ys, pose, 5.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(a[0]) > 10


@pytest.mark.parametrize("kind,density,seed", [("uniform", 0.05, 1), ("uniform", 0.5, 2), ("uniform", 5.0, 3),
                                               ("walls", 8.0, 4)])
def test_eval_nodes_identical(port, refm, ack, kind, density, seed):
    extent = 50.0
    n = int(density * extent * extent)
    xs, ys = (wl.uniform_cloud(n, extent, seed=seed) if kind == "uniform" else wl.walls_cloud(n, extent, seed=seed))
    poses = wl.node_poses(48, extent, margin=1.0, seed=seed + 50)
    o0, _ = port.eval_nodes(ack[0], xs, ys, poses, 5.0, mode=0, nthreads=0)
    r0, _ = refm.eval_nodes(ack[1], xs, ys, poses, 5.0, mode=0, nthreads=0)
    r1, _ = refm.eval_nodes(ack[1], xs, ys, poses, 5.0, mode=1, nthreads=0)
    assert np.array_equal(o0, r0) and np.array_equal(r0, r1)
    if density <= 0.5:
        assert (o0 == 5.0).any() and (o0 < 5.0).any()


def test_points_inside_robot_shape_identical(port, refm, ack):
    """COLL_BEH_BACK_AWAY branch of internal_TPObsDistancePostprocess: obstacles inside the polygon."""
    rng = np.random.default_rng(9)
    xs = rng.uniform(9.6, 10.4, 400).astype(np.float32)
    ys = rng.uniform(9.7, 10.3, 400).astype(np.float32)
    poses = np.array([[10.0, 10.0, a] for a in np.linspace(-3, 3, 13)])
    for i in range(0, 400, 40):  # few points per evaluation so that single points decide
        o, _ = port.eval_nodes(ack[0], xs[i:i + 3], ys[i:i + 3], poses, 5.0)
        r, _ = refm.eval_nodes(ack[1], xs[i:i + 3], ys[i:i + 3], poses, 5.0)
        assert np.array_equal(o, r)
        assert (o < 5.0).any()


def test_diffdrive_inverse_map_and_steps_identical(ack):
    rng = np.random.default_rng(5)
    for a, b in zip(*ack):
        for _ in range(300):
            x, y = rng.uniform(-5, 5, 2)
            if rng.random() < 0.1:
                y = 0.0
            tol = rng.choice([0.1, 0.4])
            assert a.inverse_map(x, y, tol) == b.inverse_map(x, y, tol)
            k = int(rng.integers(0, 121))
            d = rng.uniform(0, 6)
            assert a.step_for_dist(k, d) == b.step_for_dist(k, d)
            s = int(rng.integers(0, a.step_count(k)))
            assert np.array_equal(a.path_pose(k, s), b.path_pose(k, s))
            assert np.array_equal(a.path_twist(k, s), b.path_twist(k, s))


def test_holonomic_ptg_queries_identical(holo):
    """HolonomicBlend.cpp:301-840 under random dynamic states and trimmed speeds, including the
    never-invalidated step-count cache (same call sequence on both sides)."""
    a, b = holo
    rng = np.random.default_rng(7)
    for it in range(120):
        vel = [rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)] if it % 3 else [0.0, 0.0, 0.0]
        tgt = [rng.uniform(-6, 6), rng.uniform(-6, 6), rng.uniform(-3, 3)]
        trs = float(rng.choice([0.0, 0.5, 1.0]))
        speed = float(rng.choice([1.0 / 3.0, 1.0 / 3.0 + 1.0 / 3.0, 1.0]))
        for p in (a, b):
            p.set_trim_speed(speed)
            p.update_dyn_state(vel, tgt, trs)
        for _ in range(6):
            k = int(rng.integers(0, 191))
            na, nb = a.step_count(k), b.step_count(k)
            assert na == nb
            s = int(rng.integers(1, max(2, na)))
            assert np.array_equal(a.path_pose(k, s), b.path_pose(k, s))
            assert np.array_equal(a.path_twist(k, s), b.path_twist(k, s))
            assert a.init_dist(k) == b.init_dist(k)
            d = rng.uniform(0, 5)
            assert a.step_for_dist(k, d) == b.step_for_dist(k, d)
            x, y = rng.uniform(-4, 4, 2)
            ra, rb = a.inverse_map(x, y), b.inverse_map(x, y)
            assert ra[0] == rb[0] and (not ra[0] or ra == rb)
            ox, oy = rng.uniform(-3, 3, 2)
            ta, tb = a.update_tpobs_single(ox, oy, k, 5.0), b.update_tpobs_single(ox, oy, k, 5.0)
            assert ta == tb


def test_holonomic_candidates_identical(port, refm, holo):
    rng = np.random.default_rng(11)
    xs, ys = wl.uniform_cloud(2500, 30.0, seed=13)
    B, M = 24, 600
    nodes = np.c_[rng.uniform(3, 27, B), rng.uniform(3, 27, B), rng.uniform(-np.pi, np.pi, B),
                  rng.uniform(-1, 1, B), rng.uniform(-1, 1, B), rng.uniform(-0.5, 0.5, B)]
    nodes[::4, 3:] = 0
    rel = np.c_[rng.uniform(-8, 8, B), rng.uniform(-8, 8, B), rng.uniform(-3, 3, B)]
    cn = rng.integers(0, B, M)
    ck = rng.integers(0, 191, M)
    cs = rng.choice([1.0 / 3.0, 2.0 / 3.0, 1.0], M)
    ra = port.eval_candidates(holo[0], xs, ys, nodes, rel, cn, ck, cs, 5.0)
    rb = refm.eval_candidates(holo[1], xs, ys, nodes, rel, cn, ck, cs, 5.0)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    assert np.isinf(ra[0]).any() and (ra[0] < 5).any()


@pytest.mark.parametrize("use_average", [False, True])
def test_cost_evaluators_identical(port, refm, use_average):
    xs, ys = wl.walls_cloud(3000, 25.0, seed=21)
    rng = np.random.default_rng(22)
    E = 400
    frm = np.c_[rng.uniform(-2, 27, E), rng.uniform(-2, 27, E), rng.uniform(-np.pi, np.pi, E)]
    cnt = rng.integers(1, 8, E)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    rel = rng.uniform(-3, 3, (off[-1], 2))
    outs = []
    for M in (port, refm):
        cm = M.CostMap(xs, ys, 0.05, 0.4, 2.0, use_average, 0.0, None)
        cm2 = M.CostMap(xs, ys, 0.04, 0.5, 4.0, use_average, 6.0, [12.0, 12.0, 0.3])
        wp = M.Waypoints([3.0, 5.0, 9.0, 20.0], [2.0, 5.0, 8.0, 7.0], 1.5, 0.5, use_average)
        outs.append((cm.info(), cm.cells(), cm2.info(), cm2.cells(), M.eval_edge_costs(cm, frm, off, rel),
                     M.eval_edge_costs(cm2, frm, off, rel), M.eval_edge_costs(wp, frm, off, rel)))
    a, b = outs
    assert a[0] == b[0] and a[2] == b[2]
    for i in (1, 3, 4, 5, 6):
        assert np.array_equal(a[i], b[i]), i
    assert (a[4] > 0).any() and (a[6] < 0.5).any()


def _plan_both(port, refm, mk, params, obstacles, query, costmap=None, waypoints=False, goal_is_point=False):
    res = []
    for M in (port, refm):
        pl = M.Planner(getattr(M, mk)(), params, wl.DEMO_ASTAR_TIMESTAMPS)
        pl.add_obstacles(*obstacles)
        if costmap:
            pl.add_costmap(M.CostMap(obstacles[0], obstacles[1], *costmap))
        if waypoints:
            pl.add_waypoints(M.Waypoints(WAYPOINTS01[0], WAYPOINTS01[1], 1.5, 0.5, True))
        r = pl.plan(*query, goal_is_point=goal_is_point)
        tree = pl.tree()
        if r["best_node"] >= 0:
            pl.refine()
        nodes, edges = pl.path()
        txt = pl.trajectory_txt() if len(nodes) > 1 else ""
        res.append((r, tree, nodes, edges, txt))
    (ra, ta, na, ea, xa), (rb, tb, nb, eb, xb) = res
    for key in ("success", "path_cost", "tree_nodes", "tree_parents", "expanded", "tps_points", "edges_costed",
                "best_node", "goal_node", "best_cost_to_goal"):
        assert ra[key] == rb[key], (key, ra[key], rb[key])
    assert np.array_equal(ta, tb) and np.array_equal(na, nb) and np.array_equal(ea, eb) and xa == xb
    return ra


def _demo():
    obs = np.loadtxt(os.path.join(os.path.dirname(SHARE), "obstacles_01.txt")).astype(np.float32)
    xs, ys = obs[:, 0].copy(), obs[:, 1].copy()
    start, goal = [0.5, 0, 0, 0, 0, 0], [4, 2.5, np.deg2rad(45.0), 0, 0, 0]
    lo = np.minimum(np.minimum(obs.min(0), np.float32([0.5 - 2, 0 - 2])), np.float32([4 - 2, 2.5 - 2]))
    hi = np.maximum(np.maximum(obs.max(0), np.float32([0.5 + 2, 0 + 2])), np.float32([4 + 2, 2.5 + 2]))
    return (xs, ys), (start, goal, [float(lo[0]), float(lo[1]), -np.pi], [float(hi[0]), float(hi[1]), np.pi])


def test_c1_holonomic_plan_identical(port, refm):
    """BASELINE config C1 (README command line): whole motion tree, refined path and trajectory text."""
    obstacles, q = _demo()
    r = _plan_both(port, refm, "holonomic_ptgs", wl.DEMO_ASTAR_PARAMS, obstacles, q,
                   costmap=(0.04, 0.5, 4.0, False, 15.0, q[0][:3]))
    assert r["success"] == 1.0 and r["expanded"] > 10


def test_c2_ackermann_plan_identical(port, refm):
    obstacles, q = _demo()
    r = _plan_both(port, refm, "ackermann_ptgs", wl.DEMO_ASTAR_PARAMS, obstacles, q,
                   costmap=(0.04, 0.5, 4.0, False, 15.0, q[0][:3]), waypoints=True)
    assert r["success"] == 1.0


@pytest.mark.parametrize("mk", ["holonomic_ptgs", "ackermann_ptgs"])
def test_point_goal_plan_identical(port, refm, mk):
    """R(2) goal (PoseOrPoint holding a point): default_heuristic_R2 and the goal-cell redefinition."""
    obstacles, q = _demo()
    goal = [4.0, 2.5, 0.0, 0, 0, 0]
    _plan_both(port, refm, mk, wl.DEMO_ASTAR_PARAMS, obstacles, (q[0], goal, q[2], q[3]), goal_is_point=True)


@pytest.mark.parametrize("mk,cap", [("ackermann_ptgs", 60), ("holonomic_ptgs", 25)])
def test_capped_queries_on_wall_map_identical(port, refm, mk, cap):
    """C4-shaped queries (random start/goal on a wall map, per-query world box) under an expansion
    cap. On the reference side the cap is applied from outside through the planner's own progress
    callback (oracle/ref/ref_capi.cpp), the reference code is untouched."""
    xs, ys = wl.walls_cloud(40000, 60.0, wall_fraction=1.0, spacing=0.02, seed=33)
    params = list(wl.DEMO_ASTAR_PARAMS)
    params[8] = cap
    n_nonempty = 0
    for q in wl.planning_queries(6, extent=60.0, dmin=6.0, dmax=14.0, seed=34):
        try:
            r = _plan_both(port, refm, mk, params, (xs, ys), q[:4])
        except RuntimeError as e:  # both sides must fail alike is not checkable; only tolerate start-in-collision
            pytest.fail(str(e))
        n_nonempty += r["tree_nodes"] > 1
    assert n_nonempty >= 3


# [grid_xy, grid_yaw, SE2_metricAngleWeight, heuristic_heading_weight, max_trajs, max_speeds, interp segments, time, cap]
PARAM_SETS = {
    "defaults_of_the_class": ([0.20, np.deg2rad(5.0), 1.0, 0.1, 20, 3, 5, 0, 30], [1.0, 3.0, 5.0]),
    "coarse_two_speeds": ([0.50, np.deg2rad(45.0), 0.3, 0.0, 7, 2, 3, 0, 60], [0.4, 2.0]),
    "one_speed_many_paths": ([0.30, np.deg2rad(15.0), 0.1, 0.5, 31, 1, 8, 0, 40], [0.25, 0.75, 1.5, 3.0]),
}


@pytest.mark.parametrize("mk", ["holonomic_ptgs", "ackermann_ptgs"])
@pytest.mark.parametrize("pname", sorted(PARAM_SETS))
def test_other_planner_parameters_identical(port, refm, mk, pname):
    """TPS_Astar under other parameter sets than the demo's (lattice resolution, number of trajectories and speeds,
    sample timestamps, interpolated segments, heuristic weights), with costmap and waypoint evaluators."""
    params, stamps = PARAM_SETS[pname]
    obstacles, q = _demo()
    res = []
    for M in (port, refm):
        pl = M.Planner(getattr(M, mk)(), params, stamps)
        pl.add_obstacles(*obstacles)
        pl.add_costmap(M.CostMap(obstacles[0], obstacles[1], 0.05, 0.4, 2.0, True, 0.0, None))
        pl.add_waypoints(M.Waypoints(WAYPOINTS01[0], WAYPOINTS01[1], 1.5, 0.5, False))
        r = pl.plan(*q)
        res.append((r, pl.tree()))
    (ra, ta), (rb, tb) = res
    for key in ("success", "path_cost", "tree_nodes", "expanded", "tps_points", "edges_costed", "best_node"):
        assert ra[key] == rb[key], (key, ra[key], rb[key])
    assert np.array_equal(ta, tb) and ra["expanded"] >= 5


HOLO_VARIANTS = [
    # num_paths, ref_distance, T_ramp_max, v_max, w_max, radius, expr_V, expr_W
    (61, 4.0, 0.5, 0.8, np.deg2rad(90.0), 0.30, "", ""),
    (101, 6.0, 2.0, 1.5, np.deg2rad(30.0), 0.10, "V_MAX * max(0.2, trimmable_speed) * (1 - 0.3*abs(dir)/PI)",
     "W_MAX * min(1.0, abs(dir))"),
]


@pytest.mark.parametrize("variant", range(len(HOLO_VARIANTS)))
def test_holonomic_other_ptg_parameters_identical(port, refm, variant):
    """HolonomicBlend with other parameters and user expressions than the demo's: candidates and a capped plan."""
    v = HOLO_VARIANTS[variant]
    a, b = port.holonomic_blend(*v), refm.holonomic_blend(*v)
    rng = np.random.default_rng(40 + variant)
    xs, ys = wl.uniform_cloud(1500, 25.0, seed=41 + variant)
    B, M = 12, 300
    nodes = np.c_[rng.uniform(3, 22, B), rng.uniform(3, 22, B), rng.uniform(-np.pi, np.pi, B),
                  rng.uniform(-0.7, 0.7, B), rng.uniform(-0.7, 0.7, B), rng.uniform(-0.3, 0.3, B)]
    rel = np.c_[rng.uniform(-6, 6, B), rng.uniform(-6, 6, B), rng.uniform(-3, 3, B)]
    cn, ck, cs = rng.integers(0, B, M), rng.integers(0, v[0], M), rng.choice([0.5, 1.0], M)
    ra = port.eval_candidates(a, xs, ys, nodes, rel, cn, ck, cs, v[1])
    rb = refm.eval_candidates(b, xs, ys, nodes, rel, cn, ck, cs, v[1])
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and np.isfinite(ra[0]).any()
    obstacles, q = _demo()
    params = list(wl.DEMO_ASTAR_PARAMS)
    params[8] = 25
    res = []
    for M_, ptg in ((port, a), (refm, b)):
        pl = M_.Planner([ptg], params, wl.DEMO_ASTAR_TIMESTAMPS)
        pl.add_obstacles(*obstacles)
        r = pl.plan(*q)
        res.append((r["expanded"], r["path_cost"], pl.tree()))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2])
