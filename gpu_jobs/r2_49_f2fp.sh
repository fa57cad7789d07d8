#!/bin/bash
# round 2, call 49: which pipe does F2FP.F16.F32.PACK_AB (the P pack of the softmax) use, and at what rate next to MUFU.EX2?
cd tests/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe mufu_probe.cu && PROBE_ONLY_HMMA=1 /tmp/mufu_probe
