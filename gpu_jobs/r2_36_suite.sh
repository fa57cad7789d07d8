#!/bin/bash
# round 2, call 36: full GPU suite (new: branch layouts, fused EDM kernels) + config 3 with the four-stream default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_36_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/r2_36_tests.log; grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2_36_tests.log | head -20
timeout 900 python bench.py --config 3 --no-cpu > gpurun_out/r2_36_config3.json 2> gpurun_out/r2_36_config3.err; echo "config3 rc=$?"; tail -c 700 gpurun_out/r2_36_config3.json
