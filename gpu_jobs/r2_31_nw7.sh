#!/bin/bash
# round 2, call 31: 7-warp (112-query) attention CTAs at 28x28 tokens; fused EDM kernels test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention or layout" 2>&1 | tail -3
for v in 0 1; do
  echo "== NW7=$v B=1024"; CNB_ATTN_NW7=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
  echo "== NW7=$v B=128"; CB_BATCH=128 CNB_ATTN_NW7=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
done
