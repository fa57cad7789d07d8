#!/bin/bash
# round 2, call 29: 8-channel-vector staged GroupNorm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "groupnorm or gn" > gpurun_out/r2_29_tests.log 2>&1
rc=$?; echo "gn kernel tests rc=$rc"; tail -2 gpurun_out/r2_29_tests.log
if [ $rc -ne 0 ]; then grep -E "^FAILED|^E  " gpurun_out/r2_29_tests.log | head -30; fi
for v in 1 0; do
  echo "== VEC8=$v"; CNB_GN_VEC8=$v CB_GN16=1 timeout 300 python tests/conv_bench.py gn 7 2>&1 | grep "^gn"
  echo "== VEC8=$v B=128"; CB_BATCH=128 CNB_GN_VEC8=$v CB_GN16=1 timeout 300 python tests/conv_bench.py gn 7 2>&1 | grep "^gn"
done
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -2
for b in 1024 128; do for v in 1 0; do
  echo "== bench B=$b VEC8=$v"; CNB_GN_VEC8=$v timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernel_families'].items()}, d['roofline']['by_family']['groupnorm']['time_weighted'])"
done; done
