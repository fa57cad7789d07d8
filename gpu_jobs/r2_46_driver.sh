#!/bin/bash
# round 2, call 46: what the driver runs at round end - smoke(), the reference arm, the default bench line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_46_ref.json 2> gpurun_out/r2_46_ref.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/r2_46_ref.json; echo
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_46_bench.json 2> gpurun_out/r2_46_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; python -c '
import json; d=json.loads(open("gpurun_out/r2_46_bench.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["roofline"]["frac"], d["cpu_baseline"]["value"])'
