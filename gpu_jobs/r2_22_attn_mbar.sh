#!/bin/bash
# round 2, call 22: mma.sync attention with the mbarrier-decoupled K/V ring
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention or conv" > gpurun_out/r2_22_tests.log 2>&1
rc=$?; echo "attention tests rc=$rc"; tail -3 gpurun_out/r2_22_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^E  " gpurun_out/r2_22_tests.log | head -30; fi
SH="784,64,4;784,16,4;196,32,4;1024,64,4;1024,16,4;784,32,4;256,64,4;196,64,4"
for m in 1 0; do
  echo "== MBAR=$m"; CNB_ATTN_MBAR=$m CB_SHAPES="$SH" CB_ATTN_KERNEL=mma timeout 600 python tests/conv_bench.py attn 7 2>&1 | grep -v "^$"
done
for b in 128; do for m in 1 0; do
  echo "== B=$b MBAR=$m"; CB_BATCH=$b CNB_ATTN_MBAR=$m CB_SHAPES="784,64,4;784,16,4" CB_ATTN_KERNEL=mma timeout 600 python tests/conv_bench.py attn 7 2>&1 | grep -v "^$"
done; done
for occ in 0 4 3 2; do
  echo "== GN warp kernel, CTAs/SM limit $occ"; CNB_GN_WARP_OCC=$occ CB_GN16=1 timeout 300 python tests/conv_bench.py gn 7 2>&1 | grep "C=256\|C=128 @7"
done
