#!/bin/bash
# round 2, call 33: PDL work threshold at the strong-scaling batch sizes (four-stream schedule)
mkdir -p gpurun_out
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"])'
for b in 128 256 512; do for w in 4194304 8388608 16777216 67108864; do
  echo "== B=$b PDL_WORK=$w"; CNB_PDL_WORK=$w timeout 600 python bench.py --batch $b --no-cpu --no-other --no-dropin --e2e-steps 2 --steps 40 2>/dev/null | python -c "$pick"
done; done
