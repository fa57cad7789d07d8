#!/bin/bash
# round 2, call 15: 2-GPU strong-scaling bench (the driver's launch line) + the data-parallel gather check
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_15_scale2.json 2> gpurun_out/r2_15_scale2.err; echo "bench N=2 rc=$?"
tail -c 1500 gpurun_out/r2_15_scale2.json; echo; tail -5 gpurun_out/r2_15_scale2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_15_ref2.json 2> gpurun_out/r2_15_ref2.err; echo "ref N=2 rc=$?"; head -c 300 gpurun_out/r2_15_ref2.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dp_gather_check.py > gpurun_out/r2_15_dpcheck.log 2>&1; echo "dp check rc=$?"; tail -4 gpurun_out/r2_15_dpcheck.log
timeout 300 python -m pytest tests/test_models_gpu.py -m gpu -q -k "activation_range" -s 2>&1 | grep -E "rel-L2|passed|failed"
