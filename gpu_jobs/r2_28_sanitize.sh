#!/bin/bash
# round 2, call 28: compute-sanitizer memcheck over the new conv epilogue / attention paths (small shapes)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "dense_f16_epilogue or conv_f16_operands or k_concat or (attention_f16 and not growing)" > gpurun_out/r2_28_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/r2_28_memcheck.log
