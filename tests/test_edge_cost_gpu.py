"""GPU parity of stage (c): cost evaluators over interpolated edge samples, libmpp_b200 (C ABI) vs the
oracle. Bar: bit-exact (no transcendental on the path; same operation order, no FMA contraction)."""
import math

import numpy as np
import pytest

from mrpt_path_planning_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


def random_edges(rng, n, extent, max_samples=7):
    frm = np.stack([rng.uniform(0, extent, n), rng.uniform(0, extent, n), rng.uniform(-np.pi, np.pi, n)], axis=1)
    counts = rng.integers(2, max_samples + 1, n)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    rel = rng.uniform(-3, 3, (int(off[-1]), 2))
    rel[off[:-1]] = 0.0  # first sample is the start pose itself
    return frm, off, rel


@pytest.mark.parametrize("use_average", [False, True])
def test_costmap_edge_costs_bit_exact(ctx, orc, use_average):
    xs, ys = wl.uniform_cloud(600, 20.0, seed=4)
    cm = orc.CostMap(xs, ys, 0.04, 0.5, 4.0, use_average, 0.0, None)
    info, cells = cm.info(), cm.cells()
    assert (cells > 0).any()
    ctx.clear_evaluators()
    ctx.set_costmap(0, cells, info["xmin"], info["ymin"], info["res"], use_average)
    frm, off, rel = random_edges(np.random.default_rng(8), 5000, 20.0)
    frm[:50, :2] += 100.0  # edges outside the grid -> 0
    got = ctx.eval_edge_costs(frm, off, rel)
    ref = orc.eval_edge_costs(cm, frm, off, rel)
    assert got.shape == (5000, 1)
    assert np.array_equal(got[:, 0], ref)
    assert (ref > 0).any() and (ref[:50] == 0).all()


@pytest.mark.parametrize("use_average", [True, False])
def test_waypoint_edge_costs_bit_exact(ctx, orc, use_average):
    rng = np.random.default_rng(21)
    wx, wy = rng.uniform(0, 20, 40), rng.uniform(0, 20, 40)
    wx[5], wy[5] = wx[4], wy[4]  # duplicated waypoint: equal distances
    wp = orc.Waypoints(wx, wy, 1.5, 0.5, use_average)
    ctx.clear_evaluators()
    ctx.set_waypoints(0, wx, wy, 1.5, 0.5, use_average)
    frm, off, rel = random_edges(rng, 4000, 20.0)
    got = ctx.eval_edge_costs(frm, off, rel)
    ref = orc.eval_edge_costs(wp, frm, off, rel)
    assert np.array_equal(got[:, 0], ref)
    assert (ref < 0.5).any() and (ref == 0.5).any()


def test_two_evaluators_in_slot_order(ctx, orc):
    xs, ys = wl.uniform_cloud(300, 12.0, seed=14)
    cm = orc.CostMap(xs, ys, 0.05, 0.4, 2.0, False, 0.0, None)
    wp = orc.Waypoints([3.0, 5.0, 9.0], [2.0, 5.0, 8.0], 1.5, 0.5, True)
    info = cm.info()
    ctx.clear_evaluators()
    ctx.set_costmap(0, cm.cells(), info["xmin"], info["ymin"], info["res"], False)
    ctx.set_waypoints(1, [3.0, 5.0, 9.0], [2.0, 5.0, 8.0], 1.5, 0.5, True)
    assert ctx.num_evaluators == 2
    frm, off, rel = random_edges(np.random.default_rng(2), 1000, 12.0)
    got = ctx.eval_edge_costs(frm, off, rel)
    assert np.array_equal(got[:, 0], orc.eval_edge_costs(cm, frm, off, rel))
    assert np.array_equal(got[:, 1], orc.eval_edge_costs(wp, frm, off, rel))


def test_no_evaluators_and_empty_batch(ctx):
    ctx.clear_evaluators()
    frm, off, rel = random_edges(np.random.default_rng(1), 3, 5.0)
    assert ctx.eval_edge_costs(frm, off, rel).shape == (3, 0)
    ctx.set_waypoints(0, [1.0], [1.0], 1.0, 1.0, True)
    assert ctx.eval_edge_costs(frm[:0], off[:1], rel[:0]).shape == (0, 1)


def test_costmap_nearest_distance_and_cost_cells(ctx, orc):
    """costmap construction: GPU nearest float squared distance + host pow() == the oracle's cells"""
    xs, ys = wl.walls_cloud(3000, 25.0, seed=3)
    D, max_cost, res = 0.5, 4.0, 0.04
    cm = orc.CostMap(xs, ys, res, D, max_cost, False, 6.0, [12.0, 12.0, 0.0])
    info, ref = cm.info(), cm.cells()
    d2 = ctx.costmap_nearest_d2(xs, ys, info["xmin"], info["ymin"], info["res"], info["nx"], info["ny"], D)
    d = np.sqrt(d2)  # float32 sqrt, IEEE exact
    inside = d < np.float32(D)
    assert np.array_equal(inside, ref > 0) or np.array_equal(inside & (ref > 0), ref > 0)
    got = np.zeros_like(ref)
    ratio = (d[inside] / np.float32(D)).astype(np.float64)
    got[inside] = [max_cost * math.pow(-0.99999 + 1.0 / r, 0.4) for r in ratio]  # libm pow, as the reference
    assert np.array_equal(got, ref)


def test_clip_box_matches_reference_clipping(capi, orc, oracle_ackermann, product_ackermann):
    """mppb_set_obstacles_clipped == ObstacleSource::apply_clipping_box followed by the plain upload"""
    c = capi.Context(0)
    for p in product_ackermann:
        c.add_ptg(p)
    xs, ys = wl.uniform_cloud(20000, 40.0, seed=6)
    box = np.float32([8.0, 9.5, 31.25, 30.0])
    xs[:4], ys[:4] = box[[0, 2, 0, 2]], box[[1, 1, 3, 3]]  # points exactly on the (inclusive) bounds
    keep = ~((xs < box[0]) | (xs > box[2]) | (ys < box[1]) | (ys > box[3]))
    poses = wl.node_poses(64, 40.0, margin=6.0, seed=7)
    c.set_obstacles_clipped(xs, ys, box)
    assert c.num_obstacles == int(keep.sum())
    a = c.eval_nodes(poses, 5.0)
    c.set_obstacles(xs[keep], ys[keep])
    b = c.eval_nodes(poses, 5.0)
    assert np.array_equal(a, b)
    c.close()
