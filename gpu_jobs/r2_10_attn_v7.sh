#!/bin/bash
# round 2, call 10: tmem attention v7 (one MMA issuer per CTA, PV_g(j) then QK_g(j+1) in issue order): is the in-order
# hand-over correct under stress, and what does it buy?
mkdir -p gpurun_out
for safe in 0 1; do
echo "=== SAFE=$safe"
CNB_ATTN_SAFE=$safe timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2_10_tests_s$safe.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -2 gpurun_out/r2_10_tests_s$safe.log
if [ $rc -ne 0 ]; then grep -E "^FAILED" gpurun_out/r2_10_tests_s$safe.log | head -20; fi
CNB_ATTN_SAFE=$safe timeout 300 python tests/attn_stress.py 20 | tail -4
for cfg in "0 0" "80 0" "64 0" "48 0" "0 1" "0 2"; do
  set -- $cfg
  echo "== tmem SAFE=$safe BK=$1 POLY=$2"; CNB_ATTN_SAFE=$safe CNB_ATTN_TMEM_BK=$1 CNB_ATTN_POLY=$2 CB_ATTN_KERNEL=tmem CB_ONLY=0,1,2,3,7,8 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_10_bench_s${safe}_bk$1_p$2.log
done
done
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 python tests/conv_bench.py attn 2 > gpurun_out/plain_attn.log 2>&1 &&
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 ncu --set full --clock-control none --import-source on -k regex:attention_tmem -s 2 -c 1 -o gpurun_out/r2_10_attn_tmem python tests/conv_bench.py attn 2 > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
