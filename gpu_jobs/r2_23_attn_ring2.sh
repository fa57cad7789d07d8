#!/bin/bash
# round 2, call 23: attention two-group ring (2 tiles per barrier); step time after tensormap prefetch
mkdir -p gpurun_out
SH="784,64,4;784,16,4;196,32,4;1024,64,4;1024,16,4;784,32,4;256,64,4;196,64,4"
for h in 1 2; do
  echo "== HALVES=$h (2 = + two-group ring)"; CNB_ATTN_HALVES=$h CB_SHAPES="$SH" CB_ATTN_KERNEL=mma timeout 600 python tests/conv_bench.py attn 7 2>&1 | grep -v "^$"
done
for h in 1 2; do
  echo "== B=128 HALVES=$h"; CB_BATCH=128 CNB_ATTN_HALVES=$h CB_SHAPES="784,64,4;784,16,4" CB_ATTN_KERNEL=mma timeout 600 python tests/conv_bench.py attn 7 2>&1 | grep -v "^$"
done
CNB_ATTN_HALVES=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -2
for b in 1024 128; do for h in 1 2; do
  echo "== bench B=$b HALVES=$h"; CNB_ATTN_HALVES=$h timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernel_families'].items()})"
done; done
