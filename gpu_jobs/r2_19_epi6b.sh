#!/bin/bash
# round 2, call 19: EPI 6 + partial channel chunk: kernel + model parity, step time with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv" > gpurun_out/r2_19_tests.log 2>&1
rc=$?; echo "conv kernel tests rc=$rc"; tail -3 gpurun_out/r2_19_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^E  " gpurun_out/r2_19_tests.log | head -30; fi
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -q > gpurun_out/r2_19_models.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2_19_models.log; grep -E "^FAILED" gpurun_out/r2_19_models.log | head
for v in 1 0; do
  echo "== bench EPI_TMA=$v"; CNB_CONV_EPI_TMA=$v timeout 600 python bench.py --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernel_families'].items()})"
done
for b in 128 256; do for v in 1 0; do
  echo "== B=$b EPI_TMA=$v"; CNB_CONV_EPI_TMA=$v timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
done; done
