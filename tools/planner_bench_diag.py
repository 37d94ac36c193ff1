#!/usr/bin/env python
"""Diagnostic: bench.planner_bench() alone (C1/C2 only matter here), with or without the CPU checker in the process.
   python tools/planner_bench_diag.py 0|1"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mrpt_path_planning_b200 import capi  # noqa: E402

with_cpu = len(sys.argv) > 1 and sys.argv[1] == "1"
args = argparse.Namespace(queries=16, points=1 << 16)
out = bench.planner_bench(lambda: capi.Context(0), 0, 1, args, with_cpu)
for k in ("c1_holonomic", "c2_ackermann"):
    print("with_cpu=%d %s plan %.3f ms phases %s" % (with_cpu, k, out[k]["plan_ms"], {a: round(b, 3) for a, b in out[k]["phase_ms"].items()}))
