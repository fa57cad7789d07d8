#!/bin/bash
# round 2, call 43: short-K rule A/B inside the step graph, same box, interleaved
mkdir -p gpurun_out
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["kernel_families"]["conv_tc"]["ms"])'
for rep in 1 2; do for v in 0 1; do
  echo "== B=1024 SHORTK=$v"; CNB_CONV_SHORTK=$v timeout 600 python bench.py --batch 1024 --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | python -c "$pick"
done; done
for v in 0 1; do
  echo "== B=512 SHORTK=$v"; CNB_CONV_SHORTK=$v timeout 600 python bench.py --batch 512 --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | python -c "$pick"
done
