#!/bin/bash
# round 2, call 41: short-K 1x1 convs (attention in_proj / out_proj): N tiles <= 128 with two CTAs per SM vs the default
mkdir -p gpurun_out
for v in 2 3; do
  echo "== CNB_CONV_2CTA=$v"; CB_VARIANT=f16 CNB_CONV_2CTA=$v CB_ONLY=8,9,10,11,14,15,16,17 timeout 300 python tests/conv_bench.py conv 7 2>&1 | grep "^conv"
done
