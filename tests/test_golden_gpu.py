"""The product (CUDA path through the C ABI) against the committed regression vectors under tests/golden/."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from conftest import free_dist  # noqa: E402
from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS, c1_inputs  # noqa: E402

pytestmark = pytest.mark.gpu


def test_grid_free_distance_golden(gpu_ctx, product_ackermann):
    xs, ys, poses = mg.grid_inputs()
    gpu_ctx.set_obstacles(xs, ys)
    got = free_dist(gpu_ctx.eval_nodes(poses, mg.GRID_CASE["clip"]), product_ackermann)
    gold = np.load(os.path.join(HERE, "golden", "grid_free_distance.npz"))["free"]
    assert np.array_equal(got.astype(np.float32), gold)


def test_holo_candidates_golden(capi):
    from test_tpobs_holo_gpu import make_candidates

    xs, ys, nodes, rel, cn, ck, cs = mg.holo_inputs()
    ptg = capi.holonomic_ptgs()[0]
    ctx = capi.Context(0)
    ctx.add_ptg(ptg)
    ctx.set_obstacles(xs, ys)
    ks = [mg.HOLO_CASE["ks"]] * len(nodes)
    cands, cn2, ck2, cs2 = make_candidates(ptg, nodes, rel, ks, mg.HOLO_CASE["speeds"])
    assert cn2 == cn and ck2 == ck
    got = ctx.eval_holo_candidates(nodes[:, :3], cands, mg.HOLO_CASE["clip"])
    gold = np.load(os.path.join(HERE, "golden", "holo_candidates.npz"))["raw"]
    assert np.array_equal(np.isinf(got), np.isinf(gold))
    fin = np.isfinite(gold)
    assert np.max(np.abs(got[fin] - gold[fin]) / np.maximum(np.abs(gold[fin]), 1e-12)) <= 1e-9
    ctx.close()


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_plans_golden(capi, product_ackermann, name):
    gold = np.load(os.path.join(HERE, "golden", "plans.npz"))
    xs, ys, start, goal, lo, hi = c1_inputs()
    ctx = capi.Context(0)
    ptgs = capi.holonomic_ptgs() if name == "c1" else product_ackermann
    pl = capi.Planner(ctx, ptgs, C1_PARAMS, C1_TIMESTAMPS)
    pl.add_obstacles(xs, ys)
    pl.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
    if name == "c2":
        pl.add_waypoints([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0], 1.5, 0.5, True)
    r = pl.plan(start, goal, lo, hi)
    res = np.array([r[k] for k in ("success", "path_cost", "tree_nodes", "tree_parents", "expanded", "tps_points",
                                   "edges_costed", "best_node", "goal_node")])
    assert np.array_equal(res, gold[name + "_result"])
    assert np.array_equal(pl.tree(), gold[name + "_tree"])
    nodes, edges = pl.path()
    assert np.array_equal(nodes, gold[name + "_path_nodes"]) and np.array_equal(edges, gold[name + "_path_edges"])
    pl.close()
    ctx.close()
