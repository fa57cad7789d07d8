#!/bin/bash
# round 2, call 25: second producer warp for the weight tiles (CNB_CONV_SPLIT_B); DRAM traffic capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv" > gpurun_out/r2_25_tests.log 2>&1
rc=$?; echo "conv kernel tests rc=$rc"; tail -2 gpurun_out/r2_25_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^E  " gpurun_out/r2_25_tests.log | head -30; fi
for v in 1 0; do
  echo "== SPLIT_B=$v B=1024"; CNB_CONV_SPLIT_B=$v CB_VARIANT=f16 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
  echo "== SPLIT_B=$v B=128"; CB_BATCH=128 CNB_CONV_SPLIT_B=$v CB_VARIANT=f16 CB_ONLY=0,1,2,4,8,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
done
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -2
for b in 1024 128; do for v in 1 0; do
  echo "== bench B=$b SPLIT_B=$v"; CNB_CONV_SPLIT_B=$v timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernel_families'].items()})"
done; done
fi
python profiles/traffic_capture.py 1024 > gpurun_out/plain25.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic.csv python profiles/traffic_capture.py 1024 > gpurun_out/ncu25.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu25.log; wc -l gpurun_out/traffic.csv
