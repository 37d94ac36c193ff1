#!/bin/bash
# Round-2 final profile capture on the GPU box (one GPU). Each ncu pass follows a plain run of the same command that
# exited 0. Outputs land in gpurun_out/ (scratch); tools/make_traffic.py and tools/ncu_summary.py turn them into
# the files under profiles/.   usage: bash tools/profile_round2b.sh <tag>
TAG=${1:-r2t}
BENCH="python bench.py --no-planner --no-cpu-baseline --no-structured --steps 3 --warmup 3"
set -x
$BENCH > gpurun_out/${TAG}_plain_bench.log 2>&1 || exit 1
# launch list of the whole bench command (per-launch durations; cold-cache and serialised under ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
# the two launches of one timed near-first step: near pass and rest pass (-s: warm-up launches skipped)
ncu --set full --clock-control none --import-source on -k regex:tpobs_grid_kernel -s 8 -c 2 -f -o gpurun_out/${TAG}_grid \
    $BENCH > gpurun_out/${TAG}_ncu_grid.log 2>&1
# the plain one-launch form of the same step, for comparison
MPPB_NO_GRID_NEAR=1 $BENCH > gpurun_out/${TAG}_plain_bench_nonear.log 2>&1 && \
MPPB_NO_GRID_NEAR=1 ncu --set full --clock-control none --import-source on -k regex:tpobs_grid_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_grid_plain \
    $BENCH > gpurun_out/${TAG}_ncu_grid_plain.log 2>&1
python tools/profile_targets.py expand > gpurun_out/${TAG}_plain_expand.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:expand_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_expand \
    python tools/profile_targets.py expand > gpurun_out/${TAG}_ncu_expand.log 2>&1
python tools/profile_targets.py holo > gpurun_out/${TAG}_plain_holo.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tpobs_holo_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_holo \
    python tools/profile_targets.py holo > gpurun_out/${TAG}_ncu_holo.log 2>&1
for k in grid grid_plain expand holo; do python tools/ncu_summary.py gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_${k}_ncu_full.txt 2>&1; done
python tools/make_traffic.py gpurun_out/${TAG}_grid.ncu-rep gpurun_out/${TAG}_traffic.json
ls -la gpurun_out/${TAG}_* | head -40
