#!/bin/bash
# round 2, call 6: stress the tmem attention kernel (v3) for intermittent errors, 2 vs 3 S buffers
for nb in 3 2; do echo "=== NBUF=$nb"; CNB_ATTN_TMEM_NBUF=$nb timeout 300 python tests/attn_stress.py 10; done
echo "=== NBUF=3 PDL=0"; CNB_PDL=0 CNB_ATTN_TMEM_NBUF=3 timeout 300 python tests/attn_stress.py 10
