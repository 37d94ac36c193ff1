#!/bin/bash
# Times prebuilt kernel variants (variants/lib_*.so, built with MPPB_EXTRA_NVCC) on the GPU box: copies each over
# the in-tree library and runs the device-timed part of bench.py (uniform and structured C3, and a 2^23-point cloud
# = one C5 shard) plus, with DENSITIES set, the single-node latency diagnostic. Experiments only.
cp mrpt_path_planning_b200/lib/libmpp_b200.so /tmp/lib_orig.so
for f in variants/lib_*.so; do
  cp $f mrpt_path_planning_b200/lib/libmpp_b200.so
  touch mrpt_path_planning_b200/lib/libmpp_b200.so
  echo "== $f"
  python bench.py --no-planner --no-cpu-baseline --steps 60 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms=%.4f e2e_ms=%.4f structured_ms=%.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['structured_variant']['ms_per_step']))"
  python bench.py --no-planner --no-cpu-baseline --no-structured --points 8388608 --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dense (2^23 points) kernel_ms=%.4f'%(d['ms_per_step']))"
  if [ -n "$DENSITIES" ]; then python tools/node_latency_diag.py $DENSITIES; fi
done
cp /tmp/lib_orig.so mrpt_path_planning_b200/lib/libmpp_b200.so
