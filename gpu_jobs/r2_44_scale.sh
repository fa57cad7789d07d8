#!/bin/bash
# round 2, call 44: strong-scaling bench on the final build at N GPUs (the driver's launch line)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-dropin > gpurun_out/r2_44_scale$N.json 2> gpurun_out/r2_44_scale$N.err; echo "bench N=$N rc=$?"
python -c '
import sys,json
for l in open(sys.argv[1]):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l); print(d.get("scaling"), d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("batch_per_gpu"))
' gpurun_out/r2_44_scale$N.json
tail -2 gpurun_out/r2_44_scale$N.err
