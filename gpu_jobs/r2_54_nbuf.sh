#!/bin/bash
# round 2, call 54: is the dense epilogue waiting for its staging buffers?  two vs three buffers per epilogue warp
for v in 2 0; do
  echo "== CNB_CONV_EPI_NBUF=$v (0 = auto: three where they fit)"; CB_VARIANT=f16 CNB_CONV_EPI_NBUF=$v CB_ONLY=8,9,10,11,14,16 timeout 300 python tests/conv_bench.py conv 7 2>&1 | grep "^conv"
done
echo "== with residual"
for v in 2 0; do
  echo "== CNB_CONV_EPI_NBUF=$v"; CB_RES=1 CB_VARIANT=f16 CNB_CONV_EPI_NBUF=$v CB_ONLY=9,11,16 timeout 300 python tests/conv_bench.py conv 7 2>&1 | grep "^conv"
done
