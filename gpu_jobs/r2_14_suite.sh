#!/bin/bash
# round 2, call 14: whole GPU test suite, default bench, the other BASELINE configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_14_tests.log 2>&1
rc=$?; echo "gpu tests rc=$rc"; tail -3 gpurun_out/r2_14_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/r2_14_tests.log | head -40; fi
python bench.py > gpurun_out/r2_14_bench.json 2> gpurun_out/r2_14_bench.err; echo "bench rc=$?"
for c in 3 4 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r2_14_config$c.json 2> gpurun_out/r2_14_config$c.err; echo "config $c rc=$?"; tail -c 600 gpurun_out/r2_14_config$c.json; echo; tail -3 gpurun_out/r2_14_config$c.err
done
python bench.py --mode fp32 --no-cpu --e2e-steps 20 > gpurun_out/r2_14_bench_fp32.json 2> gpurun_out/r2_14_bench_fp32.err; echo "bench fp32 rc=$?"
