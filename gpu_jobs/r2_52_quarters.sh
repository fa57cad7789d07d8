#!/bin/bash
# round 2, call 52: 16-key quarters + 64-register build (8 CTAs per SM) for d = 16 in the mma.sync attention kernel
mkdir -p gpurun_out
CNB_ATTN_QUARTERS=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention_f16" 2>&1 | tail -2
for v in 0 1; do
  echo "== QUARTERS=$v B=1024"; CNB_ATTN_QUARTERS=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
  echo "== QUARTERS=$v B=128"; CB_BATCH=128 CNB_ATTN_QUARTERS=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
done
