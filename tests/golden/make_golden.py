#!/usr/bin/env python
"""Generates (or re-checks) the committed vectors under tests/golden/.

The generator is the REFERENCE BUILD when it is available — the reference's own translation units compiled from
/root/reference as oracle/_ref/libmpp_ref.so (see oracle/Makefile, target `ref`; MRPT replaced by a stand-in) — and the
line-by-line restatement (oracle/orc_*.hpp) otherwise; both produce the same arrays (tests/test_oracle_vs_reference.py,
and tests/test_golden_cpu.py runs against both). Inputs are regenerated from seeds, only expected outputs are stored.

    python tests/golden/make_golden.py            # regenerate in memory and compare with the committed files
    python tests/golden/make_golden.py --write    # rewrite the files
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mrpt_path_planning_b200 import workloads as wl  # noqa: E402
from oracle import orc as _port, ref as _ref  # noqa: E402

orc = _ref.module() if _ref.available() else _port  # the generator (same interface either way)

GRID_CASE = dict(n=9000, extent=40.0, cloud_seed=101, nodes=48, nodes_seed=102, margin=2.0, clip=5.0)
HOLO_CASE = dict(n=1500, extent=25.0, cloud_seed=111, nodes=10, seed=112, ks=list(range(0, 191, 10)),
                 speeds=[1 / 3, 2 / 3, 1.0], clip=5.0)


def grid_inputs():
    c = GRID_CASE
    xs, ys = wl.uniform_cloud(c["n"], c["extent"], seed=c["cloud_seed"])
    return xs, ys, wl.node_poses(c["nodes"], c["extent"], margin=c["margin"], seed=c["nodes_seed"])


def holo_inputs():
    c = HOLO_CASE
    xs, ys = wl.uniform_cloud(c["n"], c["extent"], seed=c["cloud_seed"])
    rng = np.random.default_rng(c["seed"])
    b = c["nodes"]
    nodes = np.zeros((b, 6))
    nodes[:, 0] = rng.uniform(1, c["extent"] - 1, b)
    nodes[:, 1] = rng.uniform(1, c["extent"] - 1, b)
    nodes[:, 2] = rng.uniform(-np.pi, np.pi, b)
    sp, th = rng.uniform(0, 1, b) * (rng.uniform(size=b) < 0.6), rng.uniform(-np.pi, np.pi, b)
    nodes[:, 3], nodes[:, 4] = sp * np.cos(th), sp * np.sin(th)
    rel = np.stack([rng.uniform(-6, 6, b), rng.uniform(-6, 6, b), rng.uniform(-3, 3, b)], 1)
    cn, ck, cs = [], [], []
    for i in range(b):
        for s in c["speeds"]:
            for k in c["ks"]:
                cn.append(i)
                ck.append(k)
                cs.append(s)
    return xs, ys, nodes, rel, cn, ck, cs


def generate():
    from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS, c1_inputs

    files = {}
    xs, ys, poses = grid_inputs()
    free, _ = orc.eval_nodes(orc.ackermann_ptgs(), xs, ys, poses, GRID_CASE["clip"], mode=0, nthreads=0)
    files["grid_free_distance.npz"] = dict(free=free.astype(np.float32))

    xs, ys, nodes, rel, cn, ck, cs = holo_inputs()
    raw, fr, _ = orc.eval_candidates(orc.holonomic_ptgs()[0], xs, ys, nodes, rel, cn, ck, cs, HOLO_CASE["clip"])
    files["holo_candidates.npz"] = dict(raw=raw, free=fr)

    xs, ys, start, goal, lo, hi = c1_inputs()
    out = {}
    for name, ptgs, wps in (("c1", orc.holonomic_ptgs(), False), ("c2", orc.ackermann_ptgs(), True)):
        pl = orc.Planner(ptgs, C1_PARAMS, C1_TIMESTAMPS)
        pl.add_obstacles(xs, ys)
        pl.add_costmap(orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3]))
        if wps:
            pl.add_waypoints(orc.Waypoints([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0], 1.5, 0.5, True))
        r = pl.plan(start, goal, lo, hi)
        nodes_, edges_ = pl.path()
        out[name + "_result"] = np.array([r[k] for k in ("success", "path_cost", "tree_nodes", "tree_parents", "expanded",
                                                         "tps_points", "edges_costed", "best_node", "goal_node")])
        out[name + "_tree"] = pl.tree()
        out[name + "_path_nodes"] = nodes_
        out[name + "_path_edges"] = edges_
    files["plans.npz"] = out
    return files


def main():
    files = generate()
    kind = getattr(orc, "KIND", "port")
    if "--write" in sys.argv:
        for fn, arrays in files.items():
            np.savez_compressed(os.path.join(HERE, fn), **arrays)
        print("wrote", sorted(files), "from the", kind)
        return
    bad = []
    for fn, arrays in files.items():
        have = np.load(os.path.join(HERE, fn))
        for k, v in arrays.items():
            if k not in have or not np.array_equal(have[k], v):
                bad.append("%s:%s" % (fn, k))
    print("generator: %s; committed vectors %s" % (kind, "REPRODUCED" if not bad else "DIFFER: %s" % bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
