#!/bin/bash
# A/B of two prebuilt libraries (variants/lib_base.so, variants/lib_new.so) on the C4-shaped batch. Experiments only.
cp mrpt_path_planning_b200/lib/libmpp_b200.so /tmp/lib_orig.so
for rep in 1 2; do
for v in base new; do
  cp variants/lib_$v.so mrpt_path_planning_b200/lib/libmpp_b200.so
  echo "== $v (rep $rep)"; python tools/c4_diag.py 2>&1 | tail -2
done
done
cp /tmp/lib_orig.so mrpt_path_planning_b200/lib/libmpp_b200.so
