#!/bin/bash
# A/B of two prebuilt libraries (variants/lib_base.so, variants/lib_new.so) on the single-query plans. Experiments only.
cp mrpt_path_planning_b200/lib/libmpp_b200.so /tmp/lib_orig.so
for rep in 1 2; do
for v in base new; do
  cp variants/lib_$v.so mrpt_path_planning_b200/lib/libmpp_b200.so
  for which in c1 c2; do echo "== $v $which (rep $rep)"; python tools/plan_latency_diag.py $which 2>&1 | tail -2; done
done
done
cp /tmp/lib_orig.so mrpt_path_planning_b200/lib/libmpp_b200.so
