"""Development check for BASELINE configs[0] (GPU box; not collected by pytest): MNIST DDPM ControlNet, batch 16,
50 steps (t = 49 .. 0).  Prints the PSNR / max-abs of the final x_0 against the unmodified reference's fixture
(tests/golden/config1_mnist_b16_50step.npz) in both compute modes with the per-step z injected, the wall time of the
graph-replayed public sampler call for the same job, and the oracle's time for it on the box's host cores."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden, inputs, max_abs, psnr, rel_l2, syn  # noqa: E402

rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
S = importlib.import_module("controlnet-pytorch_b200.sampler")
cn = importlib.import_module("controlnet-pytorch_b200.models.controlnet")
sch = importlib.import_module("controlnet-pytorch_b200.scheduler.linear_noise_scheduler")

cfg = syn.MNIST_PARAMS
m = cn.ControlNet(cfg)
sd = syn.det_state_dict(m.state_dict(), 0)
m.load_state_dict(sd)
m = m.cuda().eval()
x, hint = inputs("config1", 16, 1, 28)
zs = [syn.det_noise(f"config1:z{k}", tuple(x.shape)) for k in range(50)]
g = golden("config1_mnist_b16_50step")
sched = sch.LinearNoiseScheduler(**syn.MNIST_DIFFUSION)
xc, hc = x.cuda(), hint.cuda()
for mode in ("fp32", "f16"):
    rt.set_mode(mode)
    xt, x0 = S.DDPMSampler(m, sched, use_graph=False).sample_eager(xc, hc, steps=50, zs=zs)
    print(f"[{mode}] final x0: PSNR {psnr(x0.cpu(), g['x0']):.1f} dB, max-abs {max_abs(x0.cpu(), g['x0']):.3e}, "
          f"rel-L2 {rel_l2(x0.cpu(), g['x0']):.3e}; x_t-1 rel-L2 {rel_l2(xt.cpu(), g['xt']):.3e}", flush=True)

rt.set_mode("f16")
smp = S.DDPMSampler(m, sched, seed=1, use_graph=True)
xh, hh = x.pin_memory(), hint.pin_memory()
out = torch.empty_like(x).pin_memory()
smp.sample(xh.cuda(), hc, steps=50)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    hc.copy_(hh, non_blocking=True)
    a, _ = smp.sample(xh.to("cuda", non_blocking=True), hc, steps=50)
    out.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
best = min(ts)
print(f"config 1 on one B200 (public sampler, host buffers in / out): {best * 1e3:.2f} ms for 16 samples x 50 steps = "
      f"{best / 50 * 1e3:.3f} ms/step, {16 / best:.0f} samples/s (50-step samples)", flush=True)

if "--cpu" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cn_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    so = O.SchedulerOracle(**syn.MNIST_DIFFUSION)
    t0 = time.perf_counter()
    with torch.no_grad():
        xt_o, x0_o = O.ddpm_sample(lambda a, t, h: O.controlnet_ddpm_forward(sd, cfg, a, t, h), so, x, hint, 50, zs)
    dt = time.perf_counter() - t0
    print(f"oracle on {torch.get_num_threads()} host threads: {dt:.1f} s for the same job = {16 / dt:.2f} samples/s; "
          f"oracle vs fixture PSNR {psnr(x0_o, g['x0']):.1f} dB", flush=True)
