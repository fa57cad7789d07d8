#!/bin/bash
# round 2, call 17: where does the dense conv epilogue spend its time (short-K 1x1 layers)?  + step A/B of the attention route
mkdir -p gpurun_out
echo "== 1x1 family, no residual"; CB_VARIANT=f16 CB_ONLY=8,9,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
echo "== 1x1 family, residual";    CB_RES=1 CB_VARIANT=f16 CB_ONLY=8,9,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
echo "== B=128"; CB_BATCH=128 CB_VARIANT=f16 CB_ONLY=0,8,9,10,11 timeout 300 python tests/conv_bench.py conv 5 2>&1 | grep -v "^$"
for t in 1 0; do
  echo "== bench CNB_ATTN_TMEM=$t"; CNB_ATTN_TMEM=$t timeout 600 python bench.py --no-cpu --no-other --no-dropin --e2e-steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernel_families'].items()})"
done
CB_VARIANT=f16 CB_ONLY=10,11 CB_ONLY_ATTN=0 python tests/conv_bench.py all 2 > gpurun_out/plain17.log 2>&1 &&
CB_VARIANT=f16 CB_ONLY=10,11 CB_ONLY_ATTN=0 ncu --set full --clock-control none --import-source on -k "regex:conv_tma_kernel|attention_f16_kernel" -c 9 -o gpurun_out/r2_17_epi python tests/conv_bench.py all 2 > gpurun_out/ncu17.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu17.log
