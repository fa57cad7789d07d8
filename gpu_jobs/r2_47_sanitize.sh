#!/bin/bash
# round 2, call 47: compute-sanitizer memcheck over the fused EDM layout kernels and the four-stream step graph (tiny models)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -x -k "layout_and_scale or sampler_graph_follows or consistency_student_vs or graphed_students" > gpurun_out/r2_47_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2_47_memcheck.log
