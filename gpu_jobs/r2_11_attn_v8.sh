#!/bin/bash
# round 2, call 11: tmem attention v8 (tile B one key tile behind tile A)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2_11_tests.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -2 gpurun_out/r2_11_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED" gpurun_out/r2_11_tests.log | head -20; fi
timeout 300 python tests/attn_stress.py 10 | tail -4
for cfg in "0 0" "80 0" "64 0" "0 1" "0 2"; do
  set -- $cfg
  echo "== tmem BK=$1 POLY=$2"; CNB_ATTN_TMEM_BK=$1 CNB_ATTN_POLY=$2 CB_ATTN_KERNEL=tmem CB_ONLY=0,1,2,3,7,8 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_11_bench_bk$1_p$2.log
done
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 python tests/conv_bench.py attn 2 > gpurun_out/plain_attn.log 2>&1 &&
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 ncu --set full --clock-control none --import-source on -k regex:attention_tmem -s 2 -c 1 -o gpurun_out/r2_11_attn_tmem python tests/conv_bench.py attn 2 > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
