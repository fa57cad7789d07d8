#!/bin/bash
# round 2, call 55: final tree - full GPU suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_55_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -2 gpurun_out/r2_55_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_55_bench.json 2> gpurun_out/r2_55_bench.err; echo "bench rc=$?"; python -c '
import json; d=json.loads(open("gpurun_out/r2_55_bench.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])'
