#!/bin/bash
# A/B of the single-query A/B overlap (MPPB_NO_OVERLAP=1 restores the blocking collision calls). Experiments only.
for rep in 1 2; do
for which in c1 c2; do
  echo "== $which blocking (rep $rep)"; MPPB_NO_OVERLAP=1 python tools/plan_latency_diag.py $which 2>&1 | tail -2
  echo "== $which overlapped (rep $rep)"; python tools/plan_latency_diag.py $which 2>&1 | tail -2
done
done
