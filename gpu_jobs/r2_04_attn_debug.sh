#!/bin/bash
# round 2: locate the wrong rows of tmem attention (NBUF = 3)
CNB_ATTN_TMEM_NBUF=3 timeout 120 python tests/attn_debug.py 5 257 256 16 2
CNB_ATTN_TMEM_NBUF=3 timeout 120 python tests/attn_debug.py 16 784 16 4 2
CNB_ATTN_TMEM_NBUF=3 CNB_ATTN_TMEM_BK=48 timeout 120 python tests/attn_debug.py 5 257 256 16 1
