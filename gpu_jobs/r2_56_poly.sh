#!/bin/bash
# round 2, call 56: polynomial 2^x on the FMA pipe for one / two of every four exponentials of the mma.sync attention kernel
mkdir -p gpurun_out
for v in 1 2; do CNB_ATTN_MMA_POLY=$v timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention_f16" 2>&1 | tail -1; done
for v in 0 1 2; do
  echo "== POLY=$v B=1024"; CNB_ATTN_MMA_POLY=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
  echo "== POLY=$v B=128"; CB_BATCH=128 CNB_ATTN_MMA_POLY=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
done
