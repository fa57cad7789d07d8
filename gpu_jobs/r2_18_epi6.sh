#!/bin/bash
# round 2, call 18: TMA-store epilogue (EPI 6): parity, micro-benchmarks against the old epilogue, step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv" > gpurun_out/r2_18_tests.log 2>&1
rc=$?; echo "conv kernel tests rc=$rc"; tail -3 gpurun_out/r2_18_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|Error|error" gpurun_out/r2_18_tests.log | head -20; fi
for v in 0 1; do
  for nbuf in 2 3; do
    [ $v = 0 ] && [ $nbuf = 3 ] && continue
    echo "== EPI_TMA=$v NBUF=$nbuf, no residual"; CNB_CONV_EPI_NBUF=$nbuf CNB_CONV_EPI_TMA=$v CB_VARIANT=f16 CB_ONLY=0,2,8,9,10,11,12 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
    echo "== EPI_TMA=$v NBUF=$nbuf, residual";    CNB_CONV_EPI_NBUF=$nbuf CNB_CONV_EPI_TMA=$v CB_RES=1 CB_VARIANT=f16 CB_ONLY=0,9,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
    echo "== EPI_TMA=$v NBUF=$nbuf, B=128";  CNB_CONV_EPI_NBUF=$nbuf CNB_CONV_EPI_TMA=$v CB_BATCH=128 CB_VARIANT=f16 CB_ONLY=0,8,9,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
  done
done
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -x > gpurun_out/r2_18_models.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2_18_models.log
for v in 1 0; do
  echo "== bench EPI_TMA=$v"; CNB_CONV_EPI_TMA=$v timeout 600 python bench.py --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernel_families'].items()})"
done
fi
