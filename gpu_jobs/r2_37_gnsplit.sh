#!/bin/bash
# round 2, call 37: cluster-split register GroupNorm at the strong-scaling batch sizes
mkdir -p gpurun_out
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["kernel_families"]["groupnorm"])'
for b in 128 256; do for w in 1 2 4 8; do
  echo "== B=$b GN_SPLIT=$w"; CNB_GN_SPLIT=$w timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | python -c "$pick"
done; done
