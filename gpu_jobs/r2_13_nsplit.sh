#!/bin/bash
# round 2, call 13: N-split of few-tile convolutions (strong-scaling regime): step time at per-GPU batch 64 / 128 / 256
mkdir -p gpurun_out
for B in 128 64 256 1024; do
for cfg in "4 2" "1 2" "1 1" "2 1"; do
  set -- $cfg
  CNB_CONV_NSPLIT_ENTER=$1 CNB_CONV_NSPLIT_STOP=$2 python bench.py --batch $B --no-cpu --e2e-steps 10 --steps 30 > gpurun_out/r2_13_b${B}_e$1_s$2.json 2> gpurun_out/r2_13_b${B}_e$1_s$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_13_b${B}_e$1_s$2.json")); print("B=$B enter=$1 stop=$2 ms/step", d["ms_per_step"], "value", d["value"])
except Exception as e: print("B=$B failed", e)
PY
done; done
