#!/bin/bash
# round 2, call 3: tmem attention v3 (NBUF S buffers, separate S / O issuer threads): tests, microbench, one ncu capture
mkdir -p gpurun_out
python controlnet-pytorch_b200/build.py > /dev/null 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2_03_tests.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -3 gpurun_out/r2_03_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^E  " gpurun_out/r2_03_tests.log | head -40; fi
for cfg in "3 0" "2 0" "3 48" ; do
  set -- $cfg
  echo "== tmem NBUF=$1 BK=$2"; CNB_ATTN_TMEM_NBUF=$1 CNB_ATTN_TMEM_BK=$2 CB_ATTN_KERNEL=tmem CB_ONLY=0,1,2,3,7,8 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_03_bench_n$1_bk$2.log
done
for p in 1 2; do
  echo "== tmem POLY=$p"; CNB_ATTN_POLY=$p CB_ATTN_KERNEL=tmem CB_ONLY=0,1 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_03_bench_p$p.log
done
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 python tests/conv_bench.py attn 2 > gpurun_out/plain_attn.log 2>&1 &&
CB_BATCH=128 CB_ATTN_KERNEL=tmem CB_ONLY=0 ncu --set full --clock-control none --import-source on -k regex:attention_tmem -s 2 -c 1 -o gpurun_out/r2_03_attn_tmem python tests/conv_bench.py attn 2 > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
