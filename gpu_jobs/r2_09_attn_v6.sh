#!/bin/bash
# round 2, call 9: tmem attention v6 (two Q tiles per CTA, 4 softmax warps per SM sub-partition)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2_09_tests.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -3 gpurun_out/r2_09_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED" gpurun_out/r2_09_tests.log | head -20; fi
timeout 300 python tests/attn_stress.py 5
for cfg in "2 0 0" "2 96 0" "2 80 0" "2 64 0" "2 0 1" "2 0 2"; do
  set -- $cfg
  echo "== tmem NBUF=$1 BK=$2 POLY=$3"; CNB_ATTN_TMEM_NBUF=$1 CNB_ATTN_TMEM_BK=$2 CNB_ATTN_POLY=$3 CB_ATTN_KERNEL=tmem CB_ONLY=0,1,2,3,7,8 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_09_bench_n$1_bk$2_p$3.log
done
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 python tests/conv_bench.py attn 2 > gpurun_out/plain_attn.log 2>&1 &&
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 ncu --set full --clock-control none --import-source on -k regex:attention_tmem -s 2 -c 1 -o gpurun_out/r2_09_attn_tmem python tests/conv_bench.py attn 2 > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
