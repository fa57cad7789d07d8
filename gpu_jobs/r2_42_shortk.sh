#!/bin/bash
# round 2, call 42: short-K 1x1 rule (N tiles <= 128, two CTAs per SM) - kernel tests, shapes, step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv" 2>&1 | tail -2
CB_VARIANT=f16 CB_ONLY=8,10,14,15 timeout 300 python tests/conv_bench.py conv 7 2>&1 | grep "^conv"
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -2
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["kernel_families"]["conv_tc"])'
for b in 1024 128; do
  echo "== B=$b"; timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | python -c "$pick"
done
