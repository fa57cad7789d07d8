#!/bin/bash
# round 2, call 32: full per-shape tables of the instrumented step at B = 1024 and 128
mkdir -p gpurun_out; rm -f gpurun_out/r2_32_shapes.jsonl
for b in 1024 128; do
  CNB_BENCH_SHAPES=gpurun_out/r2_32_shapes.jsonl timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 20 > gpurun_out/r2_32_b$b.json 2>gpurun_out/r2_32_b$b.err; echo "rc=$?"
done
