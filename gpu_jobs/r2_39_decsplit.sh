#!/bin/bash
# round 2, call 39: ConvTranspose phases as parallel branches; decoder as two batch halves from 512 samples on
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -x -k "branch or split or graph or sharding or config2" 2>&1 | tail -3
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["launches_per_step"])'
for b in 128 256; do
  echo "== B=$b (ConvT phases parallel)"; timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | python -c "$pick"
done
for b in 256 512 1024; do for d in 0 2; do
  echo "== B=$b DECODER_SPLIT=$d"; CNB_DECODER_SPLIT=$d timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 30 2>/dev/null | python -c "$pick"
done; done
