#!/bin/bash
# round 2, call 51: two 16-row query tiles per warp (2-warp CTAs) in the mma.sync attention kernel
mkdir -p gpurun_out
CNB_ATTN_RT2=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention_f16" 2>&1 | tail -2
for v in 0 1; do
  echo "== RT2=$v B=1024"; CNB_ATTN_RT2=$v CB_ONLY_ATTN=0,1,3 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
  echo "== RT2=$v B=128"; CB_BATCH=128 CNB_ATTN_RT2=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
done
