#!/bin/bash
# round 2, call 34: config 3 (CelebHQ LDM ControlNet, B = 256) - batch halves (default) vs the four-stream branch schedule
mkdir -p gpurun_out
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"))'
for v in 2 1; do
  echo "== config 3 SAMPLER_SPLIT=$v"; CNB_SAMPLER_SPLIT=$v timeout 900 python bench.py --config 3 --no-cpu 2>gpurun_out/r2_34_v$v.err | tee gpurun_out/r2_34_cfg3_v$v.json | python -c "$pick"
done
