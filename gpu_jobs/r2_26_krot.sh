#!/bin/bash
# round 2, call 26: do the few-tile / 7x7 im2col layers wait on the SAME weight lines?  rotate each tile's k-block order
for v in 0 7 13; do
  echo "== KROT=$v B=1024"; CNB_CONV_KROT=$v CB_VARIANT=f16 CB_ONLY=0,1,2,4,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
  echo "== KROT=$v B=128"; CB_BATCH=128 CNB_CONV_KROT=$v CB_VARIANT=f16 CB_ONLY=0,1,2,4,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
done
