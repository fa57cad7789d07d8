"""Multi-GPU check (GPU box, torchrun; not collected by pytest): sample_data_parallel over all ranks must reproduce
the single-rank result bit for bit (Philox noise keyed by global element index) and gather it with one NCCL
all_gather.   torchrun --nproc-per-node N tests/dp_gather_check.py"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
syn = importlib.import_module("controlnet-pytorch_b200.utils.synthetic")
cn = importlib.import_module("controlnet-pytorch_b200.models.controlnet")
sch = importlib.import_module("controlnet-pytorch_b200.scheduler.linear_noise_scheduler")
S = importlib.import_module("controlnet-pytorch_b200.sampler")

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
cfg = syn.MNIST_PARAMS
model = cn.ControlNet(cfg)
model.load_state_dict(syn.det_state_dict(model.state_dict()))
model = model.to(dev).eval()
sched = sch.LinearNoiseScheduler(**syn.MNIST_DIFFUSION)
B, steps = 40, 6        # 40 = 8 x 5: equal shards; 37: ragged; 1500: shards large enough to change the conv plan
for total in (B, 37, 1500):
    hint_all = syn.det_hint(total, 28)

    def hint_fn(lo, hi):
        return hint_all[lo:hi].to(dev)

    out = S.sample_data_parallel(model, sched, (total, 1, 28, 28), hint_fn, steps=steps, seed=11)
    if rank == 0:
        smp = S.DDPMSampler(model, sched, seed=11, use_graph=True)
        x_T = smp.draw_xT((total, 1, 28, 28), dev, elem_offset=0)
        ref, _ = smp.sample(x_T, hint_all.to(dev), steps=steps, elem_offset=0)
        same = torch.equal(out, ref)
        err = float((out - ref).abs().max())
        print(f"world={world} total={total}: gathered {tuple(out.shape)} bit-identical to 1-rank run: {same} (max abs diff {err:.3e})",
              flush=True)
        assert out.shape[0] == total
        assert same, err         # every kernel is batch-invariant
dist.barrier()
dist.destroy_process_group()
