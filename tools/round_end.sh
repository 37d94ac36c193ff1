#!/bin/bash
# Round-end check on the GPU box (one GPU): the GPU test suite, smoke(), the default bench, the reference arm and the
# profile captures, in the order the driver runs them. Outputs land in gpurun_out/<tag>_*; copy what is to be judged
# into profiles/ afterwards (tools/ncu_summary.py for the .ncu-rep files).
#   usage: gpurun --timeout 1200 -- 'bash tools/round_end.sh r2a'
TAG=${1:-rN}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err; echo "reference arm rc=$?"
bash tools/profile_round2b.sh ${TAG} > gpurun_out/${TAG}_profile.log 2>&1; echo "profile rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json")); p = d.get("planner") or {}
print("C3 %.4f ms resident, %.4f ms from host buffers; C1 %.3f ms, C2 %.3f ms, C4 %.0f plans/s" % (
    d["ms_per_step"], d["e2e"]["ms_per_step"], p["c1_holonomic"]["plan_ms"], p["c2_ackermann"]["plan_ms"],
    p["c4_batch"]["plans_per_s"]))
PY
