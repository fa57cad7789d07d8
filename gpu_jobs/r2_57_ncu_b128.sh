#!/bin/bash
# round 2, call 57: ncu launch list of the step at 128 samples per GPU on the final build (the strong-scaling regime at 8 GPUs)
mkdir -p gpurun_out
CMD="python bench.py --batch 128 --steps 2 --warmup 3 --no-cpu --no-dropin --no-other --e2e-steps 2"
$CMD > gpurun_out/r2_57_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_57_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_57_launches_b128.csv $CMD > gpurun_out/r2_57_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2_57_launches_b128.csv
gzip -f gpurun_out/r2_57_launches_b128.csv
