#!/bin/bash
# round 2, call 40: full GPU suite + default bench + small-batch points on the rebuilt library (no 7-warp attention CTAs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_40_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/r2_40_tests.log; grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2_40_tests.log | head -20
timeout 900 python bench.py > gpurun_out/r2_40_bench.json 2> gpurun_out/r2_40_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_40_bench.json; echo
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["launches_per_step"])'
for b in 128 256 512; do
  echo "== B=$b"; timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | tee gpurun_out/r2_40_b$b.json | python -c "$pick"
done
