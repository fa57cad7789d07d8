#!/bin/bash
# round 2, call 21: attention halves at 7 vs 8 CTAs / SM; per-shape step table; ncu of the halves kernel
mkdir -p gpurun_out
SH="784,64,4;784,16,4;196,32,4;1024,64,4;1024,16,4;784,32,4;256,64,4"
for h in 1 2; do
  echo "== HALVES=$h"; CNB_ATTN_HALVES=$h CB_SHAPES="$SH" CB_ATTN_KERNEL=mma timeout 600 python tests/conv_bench.py attn 7 2>&1 | grep -v "^$"
done
CNB_ATTN_HALVES=2 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -2
timeout 600 python tests/step_profile.py > gpurun_out/r2_21_step_profile.log 2>&1; head -45 gpurun_out/r2_21_step_profile.log
CB_SHAPES="784,64,4" CB_ATTN_KERNEL=mma python tests/conv_bench.py attn 2 > gpurun_out/plain21.log 2>&1 &&
CB_SHAPES="784,64,4" CB_ATTN_KERNEL=mma ncu --set full --clock-control none --import-source on -k "regex:attention_f16_kernel" -s 1 -c 1 -o gpurun_out/r2_21_attn python tests/conv_bench.py attn 2 > gpurun_out/ncu21.log 2>&1
echo "ncu rc=$?"
