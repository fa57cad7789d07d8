#!/bin/bash
# round 2, call 12: tmem attention (v7) vs mma.sync on every attention shape of the three model families -> routing table
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -2
SH="784,64,4;784,16,4;196,128,4;196,32,4;49,256,4;49,128,4;49,64,4;1024,128,4;1024,16,4;256,256,4;256,64,4;64,256,4;64,128,4;1024,384,16,256;1024,128,16,256;256,512,16,256;256,256,16,256;64,768,16,256;64,384,16,256;16,512,16,256"
for k in mma tmem; do
  echo "== kernel $k"; CB_SHAPES="$SH" CB_ATTN_KERNEL=$k timeout 600 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_12_table_$k.log
done
for B in 128 16; do for k in mma tmem; do
  echo "== kernel $k batch $B"; CB_BATCH=$B CB_SHAPES="784,64,4;784,16,4;196,128,4;196,32,4;49,256,4;49,128,4" CB_ATTN_KERNEL=$k timeout 600 python tests/conv_bench.py attn 20 2>&1 | grep -v "^$" | tee gpurun_out/r2_12_table_${k}_b$B.log
done; done
