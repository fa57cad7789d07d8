#!/bin/bash
# round 2, call 35: full 1000-step e2e of the default bench on the four-stream schedule (and the round-1 layout beside it)
mkdir -p gpurun_out
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["job_seconds"])'
for v in 1 2; do
  echo "== BRANCH_PARALLEL=$v"; CNB_BRANCH_PARALLEL=$v timeout 600 python bench.py --no-cpu --no-other --no-dropin 2>/dev/null | python -c "$pick"
done
