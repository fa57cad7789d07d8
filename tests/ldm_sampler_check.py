"""Development check for BASELINE config 3 end to end (GPU box; not collected by pytest): CelebHQ LDM ControlNet sampling
loop replayed from the CUDA graph (sampler.LDMSampler) + VAE decode of the final latents, batch 256 by default.
Reports ms per timestep under graph replay and the extrapolated time / throughput of a 1000-step run."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import syn  # noqa: E402

rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
rt.set_mode("f16")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
CN = importlib.import_module("controlnet-pytorch_b200.models.controlnet_ldm").ControlNet
VAE = importlib.import_module("controlnet-pytorch_b200.models.vae").VAE
S = importlib.import_module("controlnet-pytorch_b200.sampler")
sch = importlib.import_module("controlnet-pytorch_b200.scheduler.linear_noise_scheduler")
m = CN(4, syn.CELEBHQ_LDM_PARAMS, down_sample_factor=32)
m.load_state_dict(syn.det_state_dict(m.state_dict(), 0))
m = m.cuda().eval()
vae = VAE(3, syn.CELEBHQ_VAE_PARAMS)
vae.load_state_dict(syn.det_state_dict(vae.state_dict(), 1))
vae = vae.cuda().eval()
sched = sch.LinearNoiseScheduler(ldm_scheduler=True, **syn.CELEBHQ_DIFFUSION)
smp = S.LDMSampler(m, sched, vae, seed=3, use_graph=True)
hint = (torch.rand(B, 1, 1024, 1024, device="cuda") < 0.05).float().expand(B, 3, 1024, 1024).contiguous()
x_T = smp.draw_xT((B, 4, 32, 32), torch.device("cuda"))
t0 = time.time()
with torch.no_grad():
    smp.sample(x_T, hint, steps=K)             # capture (hint pyramid + warm-up + graph) and one run
torch.cuda.synchronize()
print(f"capture + first {K}-step run: {time.time() - t0:.1f} s, {smp.launches_per_step} launches per step", flush=True)
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
with torch.no_grad():
    e0.record()
    smp.replay_steps(K, reset=True)
    e1.record()
    ims = smp.decode(smp.xt)
    e2.record()
torch.cuda.synchronize()
step = e0.elapsed_time(e1) / K
dec = e1.elapsed_time(e2)
total = 1000 * step + dec
print(f"B={B}: {step:.2f} ms per timestep (graph replay) = {64.64 * B / step:.0f} TFLOP/s model; decode {dec:.1f} ms; "
      f"1000 steps + decode = {total / 1e3:.2f} s -> {B / total * 1e3:.2f} images/s per GPU; "
      f"images {tuple(ims.shape)} finite={bool(torch.isfinite(ims).all())} flag={rt.lib().cnb_tc_error_flag()} "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
