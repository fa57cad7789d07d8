#!/bin/bash
# round 2, call 48: timing probe - the mma.sync attention kernel without its per-tile block barrier (results invalid):
# the upper bound of what decoupling the warps of a CTA could give
for v in 0 1; do
  echo "== NOBAR=$v"; CNB_ATTN_PROBE_NOBAR=$v CB_ONLY_ATTN=0,1 timeout 300 python tests/conv_bench.py attn 7 2>&1 | grep "^attn"
done
