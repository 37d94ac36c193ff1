#!/bin/bash
# A/B of one of the library's measurement switches on the single-query plans:  bash tools/ab_env.sh MPPB_NO_DONE_WORD
# (the variable set = the feature off). Experiments only.
VAR=${1:?switch name}
for rep in 1 2; do
for which in c1 c2; do
  echo "== $which, $VAR=1 (rep $rep)"; env $VAR=1 python tools/plan_latency_diag.py $which 2>&1 | tail -2
  echo "== $which, default (rep $rep)"; python tools/plan_latency_diag.py $which 2>&1 | tail -2
done
done
