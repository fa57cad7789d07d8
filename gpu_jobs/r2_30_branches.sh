#!/bin/bash
# round 2, call 30: full GPU suite on the restored tree + the four-stream ControlNet schedule (A/B against the round-1 layout)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_30_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/r2_30_tests.log; grep -E "^FAILED|^ERROR" gpurun_out/r2_30_tests.log | head
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["launches_per_step"], d["e2e"]["value"])'
for b in 128 256 1024; do for v in 2 1; do
  echo "== bench B=$b BRANCH_PARALLEL=$v"; CNB_BRANCH_PARALLEL=$v timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 20 2>gpurun_out/r2_30_b${b}_v${v}.err | tee gpurun_out/r2_30_b${b}_v${v}.json | python -c "$pick"
done; done
