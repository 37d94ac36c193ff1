"""The oracle against the committed regression vectors (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402


def test_grid_free_distance_golden(orc, oracle_ackermann):
    xs, ys, poses = mg.grid_inputs()
    free, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses, mg.GRID_CASE["clip"], mode=1, nthreads=0)
    gold = np.load(os.path.join(HERE, "golden", "grid_free_distance.npz"))["free"]
    assert np.array_equal(free.astype(np.float32), gold)
    assert (gold == 5.0).any() and (gold < 5.0).any()


def test_holo_candidates_golden(orc):
    xs, ys, nodes, rel, cn, ck, cs = mg.holo_inputs()
    raw, free, _ = orc.eval_candidates(orc.holonomic_ptgs()[0], xs, ys, nodes, rel, cn, ck, cs, mg.HOLO_CASE["clip"])
    gold = np.load(os.path.join(HERE, "golden", "holo_candidates.npz"))
    assert np.array_equal(raw, gold["raw"]) and np.array_equal(free, gold["free"])


def test_plans_golden(orc, oracle_ackermann):
    from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS, c1_inputs

    gold = np.load(os.path.join(HERE, "golden", "plans.npz"))
    xs, ys, start, goal, lo, hi = c1_inputs()
    pl = orc.Planner(oracle_ackermann, C1_PARAMS, C1_TIMESTAMPS)
    pl.add_obstacles(xs, ys)
    pl.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
    pl.add_waypoints(orc.Waypoints([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0], 1.5, 0.5, True))
    pl.plan(start, goal, lo, hi)
    assert np.array_equal(pl.tree(), gold["c2_tree"])
    assert gold["c1_result"][0] == 1.0
