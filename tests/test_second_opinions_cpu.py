"""Second opinions on the arithmetic the oracle takes from MRPT (not under /root/reference, restated in
oracle/ref/mrpt_shim.h from its published behaviour): independent re-derivations with scipy / numpy of
  * the nearest-obstacle distance behind CostEvaluatorCostMap::FromStaticPointObstacles (nanoflann's float L2,
    CostEvaluatorCostMap.cpp:91-107) -> scipy.spatial.cKDTree,
  * the radius search behind CostEvaluatorPreferredWaypoint::eval_single_pose (CostEvaluatorPreferredWaypoint.cpp:
    94-120) -> cKDTree.query_ball_point,
  * TPolygon2D::contains, the crossing-number point-in-polygon test of the robot shape -> a winding-number test and
    matplotlib-free ray casting written here from the textbook definition.
Runs against both checkers (the restatement and the reference's own translation units)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from mrpt_path_planning_b200 import workloads as wl


def test_costmap_cells_agree_with_a_kdtree_nearest_distance(orc):
    rng = np.random.default_rng(11)
    xs = rng.uniform(0, 12, 400).astype(np.float32)
    ys = rng.uniform(0, 9, 400).astype(np.float32)
    res, D, max_cost = 0.05, 0.5, 4.0
    cm = orc.CostMap(xs, ys, res, D, max_cost, False, 0.0, None)
    info, cells = cm.info(), cm.cells()
    # cell centres as CDynamicGrid::idx2x gives them, narrowed to float as the reference's query does
    cx = (info["xmin"] + (np.arange(info["nx"]) + 0.5) * info["res"]).astype(np.float32)
    cy = (info["ymin"] + (np.arange(info["ny"]) + 0.5) * info["res"]).astype(np.float32)
    gx, gy = np.meshgrid(cx, cy)
    d, _ = cKDTree(np.stack([xs, ys], 1).astype(np.float64)).query(np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float64))
    d = d.reshape(cells.shape)
    # the reference works with float squared distances: compare away from the d = D edge, where float and double agree
    inside = d < D * (1 - 1e-5)
    outside = d > D * (1 + 1e-5)
    assert inside.sum() > 1000 and outside.sum() > 1000
    assert np.all(cells[outside] == 0.0)
    # cost = maxCost * (-0.99999 + D/d)^0.4 (:99-101) is ill-conditioned towards d = D, so the comparison is made on
    # the distance each cell's cost implies: float (squared) distances in the reference, double here -> ~1e-7 relative
    c = cells[inside]
    assert np.all(c > 0)
    fin = np.isfinite(c)  # (a cell centre exactly on an obstacle point has infinite cost)
    implied = D / (np.power(c[fin] / max_cost, 1.0 / 0.4) + 0.99999)
    assert np.allclose(implied, d[inside][fin], rtol=2e-6, atol=1e-9)


def test_waypoint_cost_agrees_with_a_kdtree_radius_search(orc):
    rng = np.random.default_rng(12)
    wx, wy = rng.uniform(0, 20, 60), rng.uniform(0, 20, 60)
    R, scale = 2.5, 0.8
    wp = orc.Waypoints(wx, wy, R, scale, True)
    tree = cKDTree(np.stack([wx.astype(np.float32), wy.astype(np.float32)], 1).astype(np.float64))
    worst = 0.0
    for _ in range(300):
        x, y = rng.uniform(-1, 21, 2)
        got = wp.eval_pose(x, y)
        near = tree.query_ball_point([np.float32(x), np.float32(y)], R)
        cost = scale
        for i in near:
            dist = np.hypot(np.float32(x) - np.float32(wx[i]), np.float32(y) - np.float32(wy[i]))
            cost -= scale * (1.0 - np.sqrt(dist / R))
        want = max(cost, 0.0)
        worst = max(worst, abs(got - want))
    assert worst < 1e-5  # float squared distances in the reference, double here


def _pnpoly_textbook(vx, vy, px, py):
    """crossing number with the half-open rule: an edge counts if it straddles the horizontal line through the point
    and crosses it to the right of the point"""
    inside = False
    n = len(vx)
    for i in range(n):
        j = (i - 1) % n
        if (vy[i] > py) != (vy[j] > py):
            if px < (vx[j] - vx[i]) * (py - vy[i]) / (vy[j] - vy[i]) + vx[i]:
                inside = not inside
    return inside


def _winding(vx, vy, px, py):
    """winding number by summed signed angles (independent of the crossing rule away from the boundary)"""
    a = np.arctan2(np.asarray(vy) - py, np.asarray(vx) - px)
    d = np.diff(np.append(a, a[0]))
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return abs(d.sum()) > np.pi


def test_point_in_robot_shape_agrees_with_two_rederivations(orc, oracle_ackermann):
    from mrpt_path_planning_b200 import capi

    vx, vy = capi.ACKERMANN_SHAPE_X, capi.ACKERMANN_SHAPE_Y
    ptg = oracle_ackermann[0]
    rng = np.random.default_rng(13)
    n_in = 0
    for _ in range(4000):
        px, py = rng.uniform(-0.3, 0.5), rng.uniform(-0.3, 0.3)
        got = ptg.inside_shape(px, py)
        assert got == _pnpoly_textbook(vx, vy, px, py)
        # the winding number agrees wherever the point is not within 1e-9 of an edge
        edge_dist = min(abs((vx[i] - vx[i - 1]) * (vy[i - 1] - py) - (vx[i - 1] - px) * (vy[i] - vy[i - 1])) /
                        np.hypot(vx[i] - vx[i - 1], vy[i] - vy[i - 1]) for i in range(len(vx)))
        if edge_dist > 1e-9:
            assert got == _winding(vx, vy, px, py)
        n_in += got
    assert 200 < n_in < 3800
    # on-boundary conventions of the crossing rule (half-open in y): the vertices of the shape
    for i in range(len(vx)):
        assert ptg.inside_shape(vx[i], vy[i]) == _pnpoly_textbook(vx, vy, vx[i], vy[i])
