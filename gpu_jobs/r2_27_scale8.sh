#!/bin/bash
# round 2, call 27: 8- and 4-GPU strong-scaling bench (the driver's launch line) + the data-parallel gather check at world 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-dropin > gpurun_out/r2_27_scale8.json 2> gpurun_out/r2_27_scale8.err; echo "bench N=8 rc=$?"
tail -c 900 gpurun_out/r2_27_scale8.json; echo; tail -3 gpurun_out/r2_27_scale8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tests/dp_gather_check.py > gpurun_out/r2_27_dpcheck.log 2>&1; echo "dp check N=8 rc=$?"; tail -4 gpurun_out/r2_27_dpcheck.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu --no-dropin --no-other > gpurun_out/r2_27_scale4.json 2> gpurun_out/r2_27_scale4.err; echo "bench N=4 rc=$?"
tail -c 600 gpurun_out/r2_27_scale4.json; echo
