#!/usr/bin/env python
"""Timing of the PTG table set-up: host builder vs GPU collision-grid build (run on the GPU box)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrpt_path_planning_b200 import capi  # noqa: E402

ctx = capi.Context(0)
out = {}
capi.ackermann_ptgs(ctx=ctx)  # warm-up: allocations, module load
for name, kw in (("host_builder", {}), ("gpu_builder", {"ctx": ctx})):
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        capi.ackermann_ptgs(**kw)
        ts.append(time.perf_counter() - t0)
    out[name + "_both_ptgs_s"] = min(ts)
out["last_gpu_build"] = ctx.colgrid_build_stats()
out["host_threads"] = os.cpu_count()
print(json.dumps(out))
