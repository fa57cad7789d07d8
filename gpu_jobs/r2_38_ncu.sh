#!/bin/bash
# round 2, call 38: ncu launch list of the bench command on the final build + full captures of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-dropin --no-other --e2e-steps 2"
$CMD > gpurun_out/r2_38_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_38_plain.log; exit 1; }
tail -c 300 gpurun_out/r2_38_plain.log; echo
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_38_launches.csv $CMD > gpurun_out/r2_38_ncu1.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2_38_launches.csv
python profiles/traffic_capture.py 1024 > gpurun_out/r2_38_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_f16_kernel -c 3 -f -o gpurun_out/r2_38_attn python profiles/traffic_capture.py 1024 > gpurun_out/r2_38_ncu2.log 2>&1; echo "attn full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 30 -c 12 -f -o gpurun_out/r2_38_conv python profiles/traffic_capture.py 1024 > gpurun_out/r2_38_ncu3.log 2>&1; echo "conv full rc=$?"
ls -la gpurun_out/*.ncu-rep
gzip -f gpurun_out/r2_38_launches.csv
