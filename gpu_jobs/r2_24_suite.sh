#!/bin/bash
# round 2, call 24: full GPU suite, default bench line, DRAM traffic per launch (ncu, two metrics)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_24_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/r2_24_tests.log; grep -E "^FAILED" gpurun_out/r2_24_tests.log | head
timeout 900 python bench.py > gpurun_out/r2_24_bench.json 2> gpurun_out/r2_24_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_24_bench.json
python profiles/traffic_capture.py 1024 > gpurun_out/plain24.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cnb --csv --log-file gpurun_out/traffic.csv python profiles/traffic_capture.py 1024 > gpurun_out/ncu24.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu24.log; wc -l gpurun_out/traffic.csv
