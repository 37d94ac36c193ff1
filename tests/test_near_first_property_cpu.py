"""The property the near-first form of tpobs_grid_kernel rests on, checked on the CPU against the checker (the reference's
own updateTPObstacleSingle): every entry (k, d) of a collision-grid cell has d >= dist(robot origin, cell) - slack with
slack = the robot's largest radius, so an obstacle farther than r from the node cannot lower a path's free distance below
bound = r - cell diagonal - slack. The library measures slack on the tables it is given; here the same measurement is made
on the reference's tables, and the consequence — rows computed from the points within r of a node alone are final wherever
they are <= bound — is checked bit for bit against rows computed from the whole cloud."""
import numpy as np
import pytest

from mrpt_path_planning_b200 import workloads as wl


def _slack(g):
    nx, res = g["nx"], g["res"]
    cnt = np.diff(g["offsets"]).astype(np.int64)
    cell = np.repeat(np.arange(len(cnt)), cnt)
    x0 = g["xmin"] + (cell % nx) * res
    y0 = g["ymin"] + (cell // nx) * res
    near_x = np.clip(0.0, x0, x0 + res)
    near_y = np.clip(0.0, y0, y0 + res)
    return float((np.hypot(near_x, near_y) - g["ds"].astype(np.float64)).max())


def test_table_entries_are_bounded_below_by_the_cell_distance(capi):
    from oracle.orc import ACKERMANN_SHAPE_X as X, ACKERMANN_SHAPE_Y as Y

    w = 60.0 * np.pi / 180.0
    specs = [(121, 5.0, 0.05, 1.0, w, +1.0), (121, 5.0, 0.05, 1.0, w, -1.0), (31, 3.0, 0.1, 0.8, 1.3 * w, -1.0),
             (25, 2.35, 0.1, 0.6, 1.1 * w, +1.0)]
    for (n, rd, res, v, ww, K) in specs:
        ptg = capi.PTG.diffdrive_c(n, rd, res, v, ww, K, X, Y)
        g = ptg.grid()
        s = _slack(g)
        assert 0.0 < s <= g["max_radius"] + 1e-6, (s, g["max_radius"])
        assert np.diff(g["offsets"]).sum() == len(g["ds"]) and len(g["ds"]) > 1000


@pytest.mark.parametrize("density,seed", [(26.0, 5), (8.0, 6), (2.0, 7)])
def test_rows_from_near_points_are_final_below_the_bound(orc, oracle_ackermann, capi, density, seed):
    extent = 30.0
    xs, ys = wl.uniform_cloud(int(density * extent * extent), extent, seed=seed)
    poses = wl.node_poses(48, extent, margin=1.0, seed=seed + 20)
    slack = max(_slack(p.grid()) for p in capi.ackermann_ptgs())
    diag = 0.05 * np.sqrt(2.0)
    r = 4.0 * (slack + diag)  # what the library takes (near_window)
    bound = r - diag - slack - 1e-3
    full, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses, 5.0, mode=1, nthreads=0)
    decided_nodes = decided_cols = 0
    for i, p in enumerate(poses):
        near = (np.abs(xs.astype(np.float64) - p[0]) <= r) & (np.abs(ys.astype(np.float64) - p[1]) <= r)
        row, _ = orc.eval_nodes(oracle_ackermann, xs[near], ys[near], p[None, :], 5.0, mode=1, nthreads=1)
        row = row[0]
        assert (row >= full[i]).all()  # fewer points can only leave more room
        ok = row <= bound
        assert np.array_equal(row[ok], full[i][ok]), i  # below the bound no farther point changes anything
        decided_cols += int(ok.sum())
        decided_nodes += int(ok.all())
    assert decided_cols > 0
    if density >= 26.0:
        assert decided_nodes >= 40  # dense cloud: the near points decide (almost) every node outright
