"""path-planner-cli drop-in check: the headless CLI of the product, driven with the README command lines of
the reference (README.md:59-64) on the reference's own demo inputs (copied as data under tests/golden/share),
must report the plan the oracle computes and write the same interpolated-path CSV."""
import os
import subprocess

import numpy as np
import pytest

from mrpt_path_planning_b200 import build as _build
from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS, c1_inputs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SHARE = os.path.join(HERE, "golden", "share")
OBS = os.path.join(HERE, "golden", "obstacles_01.txt")


def run_cli(tmp_path, ptg_ini, extra):
    csv = str(tmp_path / "path.csv")
    cmd = [_build.CLI_PATH, "-s", "[0.5 0 0]", "-g", "[4 2.5 45]", "--planner", "mpp::TPS_Astar", "-c",
           os.path.join(SHARE, ptg_ini), "--obstacles", OBS, "--planner-parameters",
           os.path.join(SHARE, "mvsim-demo-astar-planner-params.yaml"), "--costmap-obstacles",
           os.path.join(SHARE, "costmap-obstacles.yaml"), "--save-interpolated-path", csv] + extra
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout, open(csv).read() if os.path.exists(csv) else None


def oracle_plan(orc, ptgs, waypoints):
    xs, ys, start, goal, lo, hi = c1_inputs()
    params = list(C1_PARAMS)
    params[7] = 10.0  # maximumComputationTime of the demo yaml
    pl = orc.Planner(ptgs, params, C1_TIMESTAMPS)
    pl.add_obstacles(xs, ys)
    pl.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
    if waypoints:
        pl.add_waypoints(orc.Waypoints([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0], 1.5, 0.5, True))
    r = pl.plan(start, goal, lo, hi)
    pl.refine()
    return pl, r


def test_cli_c1_holonomic(tmp_path, orc):
    out, csv = run_cli(tmp_path, "ptgs_holonomic_robot.ini", [])
    pl, r = oracle_plan(orc, orc.holonomic_ptgs(), False)
    assert "Success: YES" in out and r["success"] == 1.0
    assert "Plan has %d overall edges, %d nodes" % (r["tree_parents"], r["tree_nodes"]) in out
    assert "Obstacles  : 155 points" in out
    assert csv == pl.trajectory_txt(period=0.25)


def test_cli_c2_ackermann_costmap_waypoints(tmp_path, orc, oracle_ackermann):
    out, csv = run_cli(tmp_path, "ptgs_ackermann_vehicle.ini",
                       ["--waypoints", os.path.join(SHARE, "mvsim-demo-waypoints01.yaml"), "--waypoints-parameters",
                        os.path.join(SHARE, "costmap-prefer-waypoints.yaml")])
    pl, r = oracle_plan(orc, oracle_ackermann, True)
    assert ("Success: YES" in out) == (r["success"] == 1.0)
    assert "Plan has %d overall edges, %d nodes" % (r["tree_parents"], r["tree_nodes"]) in out
    if r["success"] == 1.0:
        assert csv == pl.trajectory_txt(period=0.25)


def test_cli_errors(tmp_path):
    r = subprocess.run([_build.CLI_PATH, "-s", "[0 0 0]"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Required argument missing" in r.stderr
    r = subprocess.run([_build.CLI_PATH, "-s", "[0.5 0 0]", "-g", "[4 2.5 45]", "-c",
                        os.path.join(SHARE, "ptgs_holonomic_robot.ini"), "-o", "/nonexistent.txt"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "nonexistent" in r.stderr
