"""Known-answer checks that pin the oracle itself (SURVEY.md section 8c list). The reference
ships no tests or golden vectors, so these analytic cases + second opinions are the pin."""
import numpy as np


def test_index2alpha_centres(oracle_ackermann):
    # K=121, k=60 -> alpha = 0: the straight path, x grows ~1 mm per step, y stays 0
    p = oracle_ackermann[0]
    tr = p.path(60)
    assert np.all(tr[:, 1] == 0) and np.all(tr[:, 2] == 0)
    assert abs(tr[1000, 0] - 1.0) < 1e-3 and abs(tr[1000, 4] - 1.0) < 1e-3
    assert 4990 <= len(tr) <= 5010
    back = oracle_ackermann[1].path(60)
    assert np.all(back[1:, 0] < 0)


def test_paths_end_at_ref_distance(oracle_ackermann):
    for p in oracle_ackermann:
        for k in range(p.num_paths):
            tr = p.path(k)
            assert tr[-1, 4] >= 5.0 and tr[-3, 4] < 5.0
            assert np.array_equal(tr[-1, :5], tr[-2, :5])  # final record is duplicated
            assert p.init_dist(k) == 5.0


def test_local_obstacles_second_opinion(orc):
    """stage (a) against an independent numpy restatement of the same formulas."""
    rng = np.random.default_rng(3)
    xs = rng.uniform(0, 30, 5000).astype(np.float32)
    ys = rng.uniform(0, 30, 5000).astype(np.float32)
    pose = np.array([14.2, 9.9, 0.77])
    ox, oy = orc.local_obstacles(xs, ys, pose, 5.0)
    gx, gy = xs.astype(np.float64), ys.astype(np.float64)
    keep = ~((np.abs(gx - pose[0]) > 5.0) | (np.abs(gy - pose[1]) > 5.0))
    c, s = np.cos(pose[2]), np.sin(pose[2])
    ix, iy = -pose[0] * c - pose[1] * s, pose[0] * s - pose[1] * c
    ex = (ix + gx[keep] * c) + gy[keep] * s
    ey = (iy - gx[keep] * s) + gy[keep] * c
    assert len(ox) == keep.sum()
    assert np.array_equal(ox, ex.astype(np.float32)) and np.array_equal(oy, ey.astype(np.float32))
    # geometric sanity: distances to the node are preserved
    d_glob = np.hypot(gx[keep] - pose[0], gy[keep] - pose[1])
    assert np.allclose(np.hypot(ox, oy), d_glob, atol=1e-5)


def test_obstacle_dead_ahead(orc, oracle_ackermann):
    """Forward PTG, straight path k=60, obstacle at (ox,0): freeDistance ~= ox - 0.3 (polygon nose at
    x = 0.3) minus at most one 0.05 m cell because of the 4-cell dilation."""
    fwd = oracle_ackermann[0]
    for ox in (1.0, 2.02, 3.51):
        out, _ = orc.eval_nodes([fwd], [ox], [0.0], [[0.0, 0.0, 0.0]], 5.0)
        d = out[0, 60]
        assert ox - 0.3 - 0.1 <= d <= ox - 0.3 + 1e-3, (ox, d)
    out, _ = orc.eval_nodes([fwd], [-1.0], [0.0], [[0.0, 0.0, 0.0]], 5.0)
    assert out[0, 60] == 5.0  # obstacle behind a forward path does not constrain it
    # obstacle inside the robot footprint: paths that would drive further into it get 0
    out, _ = orc.eval_nodes([fwd], [0.2], [0.0], [[0.0, 0.0, 0.0]], 5.0)
    assert fwd.inside_shape(0.2, 0.0) and (out[0] == 0.0).sum() + (out[0] == 5.0).sum() == 121


def test_pnpoly_second_opinion(oracle_ackermann):
    """Polygon containment against an independent winding-angle computation."""
    from oracle.orc import ACKERMANN_SHAPE_X as X, ACKERMANN_SHAPE_Y as Y

    rng = np.random.default_rng(9)
    p = oracle_ackermann[0]
    for _ in range(2000):
        x, y = rng.uniform(-0.4, 0.4, 2)
        vx, vy = X - x, Y - y
        ang = np.arctan2(vy, vx)
        d = np.diff(np.append(ang, ang[0]))
        d = (d + np.pi) % (2 * np.pi) - np.pi
        inside = abs(d.sum()) > np.pi
        edge_dist = min(abs(vx[i] * (Y[(i + 1) % 6] - Y[i]) - vy[i] * (X[(i + 1) % 6] - X[i])) /
                        np.hypot(X[(i + 1) % 6] - X[i], Y[(i + 1) % 6] - Y[i]) for i in range(6))
        if edge_dist > 1e-9:
            assert p.inside_shape(x, y) == inside


# ---- solve_poly* ([MRPT] poly34) against numpy.roots -------------------------------------
def _real_roots_np(coeffs):
    r = np.roots(coeffs)
    return np.sort(r[np.abs(r.imag) < 1e-9].real)


def test_solve_poly4_matches_numpy_roots(orc):
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(400):
        roots = rng.uniform(-3, 3, 4)
        if np.min(np.abs(np.subtract.outer(roots, roots)) + np.eye(4)) < 0.05:
            continue  # keep the roots separated: multiple roots are ill-conditioned for any solver
        a, b, c, d = np.poly(roots)[1:]
        n, x = orc.solve_poly4(a, b, c, d)
        assert n == 4
        assert np.allclose(np.sort(x), np.sort(roots), rtol=0, atol=1e-7)
        checked += 1
    assert checked > 200
    # two real + one complex pair, and no real root
    n, x = orc.solve_poly4(*np.poly([1.0, -2.0, 0.5 + 1j, 0.5 - 1j]).real[1:])
    assert n == 2 and np.allclose(np.sort(x[:2]), [-2.0, 1.0], atol=1e-9)
    n, x = orc.solve_poly4(*np.poly([1j, -1j, 2 + 1j, 2 - 1j]).real[1:])
    assert n == 0
    for _ in range(200):
        co = rng.uniform(-2, 2, 4)
        n, x = orc.solve_poly4(*co)
        ref = _real_roots_np([1.0, *co])
        assert n == len(ref)
        assert np.allclose(np.sort(x[:n]), ref, atol=1e-6)


def test_solve_poly3_and_poly2(orc):
    rng = np.random.default_rng(12)
    for _ in range(200):
        co = rng.uniform(-2, 2, 3)
        n, x = orc.solve_poly3(*co)
        ref = _real_roots_np([1.0, *co])
        real = np.sort(x[:n]) if n != 1 else x[:1]
        assert len(ref) == (3 if n == 3 else n)
        assert np.allclose(real, ref, atol=1e-6)
    n, r = orc.solve_poly2(1.0, -3.0, 2.0)
    assert n == 2 and np.allclose(r, [1.0, 2.0])
    assert orc.solve_poly2(0.0, 0.0, 1.0)[0] == 0
    n, r = orc.solve_poly2(0.0, 2.0, -1.0)
    assert n == 1 and r[0] == 0.5


# ---- HolonomicBlend (SURVEY 8c known answers) --------------------------------------------
def test_holo_index2alpha_and_straight_path(orc):
    h = orc.holonomic_ptgs()[0]
    h.update_dyn_state(force=True)
    # K=191, k=95 -> alpha = 0; zero initial velocity, T_ramp = 1, V = 1: x(t) = t^2/2 then (t-1)+0.5
    assert np.allclose(h.path_pose(95, 50), [0.125, 0, 0, 0.125], atol=1e-15)
    assert np.allclose(h.path_pose(95, 100), [0.5, 0, 0, 0.5], atol=1e-15)
    assert np.allclose(h.path_pose(95, 250), [2.0, 0, 0, 2.0], atol=1e-12)
    ok, step = h.step_for_dist(95, 5.0)
    assert ok and step == 550 and h.step_count(95) == 550  # t = 5.5 s
    assert h.init_dist(95) == 5.0 - 0.01  # dist at step 549 = 4.99 (< refDistance)


def test_holo_single_obstacle_free_distance(orc):
    h = orc.holonomic_ptgs()[0]
    h.update_dyn_state(force=True)
    R = 0.15
    for ox in (0.3, 0.45, 0.8, 2.0, 4.5):  # ramp phase (< 0.5 + R) and cruise phase
        got = h.update_tpobs_single(ox, 0.0, 95, 5.0)
        assert abs(got - (ox - R)) < 1e-9, (ox, got)
    # obstacle behind / sideways far away never constrains the straight path
    assert h.update_tpobs_single(-1.0, 0.0, 95, 5.0) == 5.0
    assert h.update_tpobs_single(1.0, 2.0, 95, 5.0) == 5.0
    # obstacle inside the robot disc (BACK_AWAY rule): the disc leaves it after d = ox + R;
    # d < R -> ignored, d >= R -> path blocked at distance 0
    assert h.update_tpobs_single(-0.05, 0.0, 95, 5.0) == 5.0
    assert h.update_tpobs_single(0.05, 0.0, 95, 5.0) == 0.0


def test_holo_speed_trim_mixes_vf_and_vmax(orc):
    """cruise-phase collision *time* uses the trimmed vf, the distance uses V_MAX (HolonomicBlend.cpp:805,835)."""
    h = orc.holonomic_ptgs()[0]
    h.set_trim_speed(0.5)
    h.update_dyn_state(force=True)
    # vf = 0.5: ramp covers 0.25 m in 1 s; obstacle edge at 2.0-0.15 -> t = 1 + (1.85-0.25)/0.5 = 4.2 s
    got = h.update_tpobs_single(2.0, 0.0, 95, 50.0)
    assert abs(got - ((4.2 - 1.0) * 1.0 + 0.25)) < 1e-9
    h.set_trim_speed(1.0)


# ---- cost evaluators ------------------------------------------------------------------------
def test_costmap_known_values(orc):
    xs = np.array([0.0, 3.0], dtype=np.float32)
    ys = np.array([0.0, 0.0], dtype=np.float32)
    cm = orc.CostMap(xs, ys, 0.1, 0.5, 4.0, False, 0.0, None)
    info = cm.info()
    assert (info["nx"], info["ny"]) == (40, 10) and abs(info["xmin"] + 0.5) < 1e-12
    cells = cm.cells()
    # cell centre (0.25, 0.05): d = sqrt(0.25^2+0.05^2) in float
    d = np.sqrt(np.float32(np.float32(0.25) ** 2 + np.float32(0.05) ** 2))
    cx, cy = int((0.25 - info["xmin"]) / 0.1), int((0.05 - info["ymin"]) / 0.1)
    exp = 4.0 * (-0.99999 + 1.0 / float(np.float32(d) / np.float32(0.5))) ** 0.4
    assert abs(cells[cy, cx] - exp) < 1e-6
    assert cm.eval_pose(1.5, 0.0) == 0.0  # farther than the clearance distance from both points
    assert cm.eval_pose(100.0, 0.0) == 0.0  # outside the grid
    assert (cells >= 0).all()


def test_waypoint_cost_known_values(orc):
    wp = orc.Waypoints([3.0, 5.0], [2.0, 5.0], 1.5, 0.5, True)
    assert wp.eval_pose(3.0, 2.0) == 0.0  # on a waypoint: scale - scale*1
    assert wp.eval_pose(0.0, 0.0) == 0.5  # nothing within R: full scale
    d = 0.75
    exp = 0.5 - 0.5 * (1.0 - np.sqrt(d / 1.5))
    assert abs(wp.eval_pose(3.0 + d, 2.0) - exp) < 1e-7


# ---- TPS_Astar on config C1 (holonomic demo of the reference README) -------------------------
def c1_inputs():
    import os
    obs = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "obstacles_01.txt")).astype(np.float32)
    xs, ys = obs[:, 0].copy(), obs[:, 1].copy()
    start = [0.5, 0, 0, 0, 0, 0]
    goal = [4, 2.5, np.deg2rad(45.0), 0, 0, 0]
    # world bbox = obstacle bbox U (start, goal +- 2 m) in float (path-planner-cli.cpp:219-247)
    lo = np.minimum(obs.min(0), np.float32([0.5 - 2, 0 - 2]))
    lo = np.minimum(lo, np.float32([4 - 2, 2.5 - 2]))
    hi = np.maximum(obs.max(0), np.float32([0.5 + 2, 0 + 2]))
    hi = np.maximum(hi, np.float32([4 + 2, 2.5 + 2]))
    return xs, ys, start, goal, [float(lo[0]), float(lo[1]), -np.pi], [float(hi[0]), float(hi[1]), np.pi]


C1_PARAMS = [0.40, np.deg2rad(25.0), 0.1, 0.1, 13, 3, 5, 0, 0]
C1_TIMESTAMPS = [0.5, 1.5, 4.0]


def test_oracle_plans_c1(orc):
    xs, ys, start, goal, lo, hi = c1_inputs()
    ptgs = orc.holonomic_ptgs()
    pl = orc.Planner(ptgs, C1_PARAMS, C1_TIMESTAMPS)
    pl.add_obstacles(xs, ys)
    pl.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
    r = pl.plan(start, goal, lo, hi)
    assert r["success"] == 1.0 and r["goal_node"] == 1.0
    nodes, edges = pl.path()
    assert np.allclose(nodes[0, 1:4], start[:3]) and np.allclose(nodes[-1, 1:4], goal[:3])
    assert abs(nodes[-1, 7] - r["path_cost"]) < 1e-12
    assert np.isclose(edges[:, 4].sum(), r["path_cost"])
    # the path keeps the robot disc clear of every obstacle point at its nodes
    for n in nodes:
        assert np.min(np.hypot(xs - n[1], ys - n[2])) > 0.15
    # deterministic: planning twice gives the identical tree
    pl2 = orc.Planner(orc.holonomic_ptgs(), C1_PARAMS, C1_TIMESTAMPS)
    pl2.add_obstacles(xs, ys)
    pl2.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
    r2 = pl2.plan(start, goal, lo, hi)
    assert r2["path_cost"] == r["path_cost"] and np.array_equal(pl2.tree(), pl.tree())


def test_lattice_truncation(orc):
    # x2idx truncates toward zero (TPS_Astar.h:224): -0.39 -> 0, 0.39 -> 0, 0.41 -> 1 at 0.4 m
    f = lambda x: int(np.float32(x) / 0.40)
    assert (f(-0.39), f(0.39), f(0.41)) == (0, 0, 1)
