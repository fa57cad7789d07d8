#!/bin/bash
# round 2, call 53: the driver's exact N = 2 launch line (default flags: weak record, e2e_dropin on rank 0)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_53_scale2.json 2> gpurun_out/r2_53_scale2.err ) 2>&1 | grep real; echo "rc=$?"
python -c '
import sys,json
for l in open("gpurun_out/r2_53_scale2.json"):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l); print(d.get("scaling"), d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("e2e_dropin"), list(d.keys())[-6:])
    w=d.get("weak") or d.get("other")
    print("other record:", json.dumps(w)[:300] if w else None)
'
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_53_ref2.json 2> gpurun_out/r2_53_ref2.err ) 2>&1 | grep real; echo "ref rc=$?"; tail -c 300 gpurun_out/r2_53_ref2.json
