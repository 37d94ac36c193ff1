#!/bin/bash
# Round-end profile capture on the GPU box (one GPU). Each ncu pass follows a plain run of the same command that
# exited 0. Outputs land in gpurun_out/ (scratch); summaries are written to profiles/ by tools/ncu_summary.py.
#   usage: bash tools/profile_round.sh <tag>
TAG=${1:-r1}
BENCH="python bench.py --no-planner --no-cpu-baseline --steps 3 --warmup 3"
set -x
$BENCH > gpurun_out/${TAG}_plain_bench.log 2>&1 || exit 1
# launch list of the bench command (durations only, clocks untouched)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
# full captures of the dominant kernel (two launches of the timed 4096-node batch) and of the other kernels
ncu --set full --clock-control none --import-source on -k regex:tpobs_grid_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_grid \
    $BENCH > gpurun_out/${TAG}_ncu_grid.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:colgrid_sweep_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_colgrid \
    $BENCH > gpurun_out/${TAG}_ncu_colgrid.log 2>&1
python tools/profile_targets.py holo > gpurun_out/${TAG}_plain_holo.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tpobs_holo_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_holo \
    python tools/profile_targets.py holo > gpurun_out/${TAG}_ncu_holo.log 2>&1
python tools/profile_targets.py cost > gpurun_out/${TAG}_plain_cost.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:edge_cost_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_cost \
    python tools/profile_targets.py cost > gpurun_out/${TAG}_ncu_cost.log 2>&1
ls -la gpurun_out/${TAG}_*
