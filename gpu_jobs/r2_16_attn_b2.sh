#!/bin/bash
# round 2, call 16: tmem attention geometry B2 (one Q tile, two S buffers, 192 threads, 4 CTAs / SM)
mkdir -p gpurun_out
CNB_ATTN_TMEM_GEOM=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/r2_16_tests.log 2>&1
rc=$?; echo "attention tests (B2) rc=$rc"; tail -2 gpurun_out/r2_16_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED" gpurun_out/r2_16_tests.log | head -20; fi
CNB_ATTN_TMEM_GEOM=1 timeout 300 python tests/attn_stress.py 10 | tail -11
SH="784,64,4;784,16,4;196,128,4;196,32,4;1024,128,4;1024,16,4;256,64,4;1024,384,16,256;1024,128,16,256;256,512,16,256;256,256,16,256"
for p in 0 1 2; do
  echo "== tmem B2 POLY=$p"; CNB_ATTN_TMEM_GEOM=1 CNB_ATTN_POLY=$p CB_SHAPES="$SH" CB_ATTN_KERNEL=tmem timeout 600 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_16_b2_p$p.log
done
CB_BATCH=128 CNB_ATTN_TMEM_GEOM=1 CB_ATTN_KERNEL=tmem CB_ONLY=0 python tests/conv_bench.py attn 2 > gpurun_out/plain_attn.log 2>&1 &&
CB_BATCH=128 CNB_ATTN_TMEM_GEOM=1 CB_ATTN_KERNEL=tmem CB_ONLY=0 ncu --set full --clock-control none --import-source on -k regex:attention_tmem -s 2 -c 1 -o gpurun_out/r2_16_attn_b2 python tests/conv_bench.py attn 2 > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
