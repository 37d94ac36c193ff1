#!/bin/bash
# A/B of prebuilt libraries (variants/lib_*.so) on the near-first evaluation of the C3 batch. Experiments only.
cp mrpt_path_planning_b200/lib/libmpp_b200.so /tmp/lib_orig.so
for f in variants/lib_*.so; do
  cp $f mrpt_path_planning_b200/lib/libmpp_b200.so
  touch mrpt_path_planning_b200/lib/libmpp_b200.so
  echo "== $f"
  NEAR_MODES=2,2 python tools/near_diag.py 2>&1 | tail -2
done
cp /tmp/lib_orig.so mrpt_path_planning_b200/lib/libmpp_b200.so
