#!/bin/bash
# round 2, call 45: configs 4 and 5 on the final build (fused EDM layout kernels in the consistency student)
mkdir -p gpurun_out
for c in 4 5; do
  timeout 900 python bench.py --config $c --no-cpu > gpurun_out/r2_45_config$c.json 2> gpurun_out/r2_45_config$c.err; echo "config $c rc=$?"; tail -c 1500 gpurun_out/r2_45_config$c.json; echo
done
