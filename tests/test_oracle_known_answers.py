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
