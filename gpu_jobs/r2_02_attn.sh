#!/bin/bash
# round 2, call 2: first run of the tcgen05 / TMEM attention kernel: parity tests, then the microbench (mma vs tmem, POLY sweep)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r2_02_tests.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -25 gpurun_out/r2_02_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
for k in mma tmem; do
  echo "== kernel $k"; CB_ATTN_KERNEL=$k timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_02_bench_$k.log
done
for p in 1 2 3; do
  echo "== tmem POLY=$p"; CNB_ATTN_POLY=$p CB_ATTN_KERNEL=tmem CB_ONLY=0,1,2,3 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_02_bench_tmem_p$p.log
done
for bk in 96 80 64; do
  echo "== tmem BK=$bk"; CNB_ATTN_TMEM_BK=$bk CB_ATTN_KERNEL=tmem CB_ONLY=0,1,2,7 timeout 300 python tests/conv_bench.py attn 2>&1 | grep -v "^$" | tee gpurun_out/r2_02_bench_tmem_bk$bk.log
done
