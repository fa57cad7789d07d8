#!/bin/bash
# round 2, call 1: validate the host-side changes, baseline numbers, small-batch sweep, launch list at B=128
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_01_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_01_tests.log
python bench.py > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_01_ref.json 2> gpurun_out/r2_01_ref.err; echo "ref rc=$?"
for B in 64 128 256; do
for cfg in "1 0" "2 0" "2 1" "4 1"; do
  set -- $cfg
  CNB_SAMPLER_SPLIT=$1 CNB_SPLIT_KEEP_BRANCH=$2 python bench.py --batch $B --no-cpu --e2e-steps 20 --steps 30 > gpurun_out/r2_01_b${B}_s$1_k$2.json 2> gpurun_out/r2_01_b${B}_s$1_k$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_01_b${B}_s$1_k$2.json")); print("B=$B split=$1 keepbranch=$2 ms/step", d["ms_per_step"], "value", d["value"], "launches", d["launches_per_step"])
except Exception as e: print("B=$B split=$1 keep=$2 failed", e)
PY
done; done
python bench.py --batch 128 --steps 2 --warmup 1 --no-cpu --e2e-steps 2 > gpurun_out/plain128.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_01_launches_b128.csv python bench.py --batch 128 --steps 2 --warmup 1 --no-cpu --e2e-steps 2 > gpurun_out/ncu128.log 2>&1
echo "ncu rc=$?"
