#!/bin/bash
# round 2, call 20: mma.sync attention, 64-key tiles processed as two 32-key halves (register pressure -> less address re-derivation)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2_20_tests.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -3 gpurun_out/r2_20_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^E  " gpurun_out/r2_20_tests.log | head -30; fi
SH="784,64,4;784,16,4;196,128,4;196,32,4;49,256,4;1024,64,4;784,32,4"
for h in 1 0; do
  echo "== HALVES=$h"; CNB_ATTN_HALVES=$h CB_SHAPES="$SH" CB_ATTN_KERNEL=mma timeout 600 python tests/conv_bench.py attn 7 2>&1 | grep -v "^$"
done
