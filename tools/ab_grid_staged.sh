#!/bin/bash
# A/B on the GPU box: tpobs_grid_kernel with global-load point loop (default) vs the staged form (cp.async.bulk ring).
# Parity tests under both, then the C3 device step time of both, three processes each.
set -u
cd "$(dirname "$0")/.."
echo "== parity (staged) =="
MPPB_GRID_STAGED=1 python -m pytest tests/test_tpobs_grid_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -2
for v in 0 1 0 1 0 1; do
  MPPB_GRID_STAGED=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-planner 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('staged=$v', 'ms_per_step', round(d['ms_per_step'],5), 'structured', round(d['structured_variant']['ms_per_step'],5), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
