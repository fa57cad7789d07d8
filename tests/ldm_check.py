"""Development check for BASELINE config 3 (GPU box; not collected by pytest): the FULL-SIZE CelebHQ LDM ControlNet
(config/celebhq.yaml ldm_params: 190.5 M parameters, latent 4x32x32, canny hint 3x1024x1024, down_sample_factor 32).
  1. parity: eps at batch 1 against oracle/cn_oracle.py (torch CPU fp32) on deterministic weights, both modes;
  2. timing: ms per denoising step at a few batches (eager, hint feature cached), per-family kernel time.
    python tests/ldm_check.py [batches...]"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, syn  # noqa: E402

rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
cfg = syn.CELEBHQ_LDM_PARAMS
CN = importlib.import_module("controlnet-pytorch_b200.models.controlnet_ldm").ControlNet
t0 = time.time()
m = CN(4, cfg, down_sample_factor=32)
sd = syn.det_state_dict(m.state_dict(), 0)
m.load_state_dict(sd)
m = m.cuda().eval()
print(f"built: {sum(p.numel() for p in m.parameters()) / 1e6:.1f} M params in {time.time() - t0:.1f} s", flush=True)

if os.environ.get("LDM_PARITY", "1") == "1":
    import cn_oracle as O
    x = syn.det_noise("ldm_full:x", (1, 4, 32, 32))
    hint = syn.det_hint(1, 1024, p=0.05)
    t = torch.tensor([500])
    t1 = time.time()
    with torch.no_grad():
        want = O.controlnet_ldm_forward(sd, cfg, x, t, hint)
    print(f"oracle (CPU fp32, batch 1): {time.time() - t1:.1f} s", flush=True)
    for mode in ("fp32", "f16"):
        rt.set_mode(mode)
        with torch.no_grad():
            got = m(x.cuda(), t.cuda(), hint.cuda())
        print(f"mode {mode}: eps rel-L2 vs oracle = {rel_l2(got.cpu(), want):.3e}  flag={rt.lib().cnb_tc_error_flag()}", flush=True)

rt.set_mode("f16")
for B in [int(a) for a in sys.argv[1:]] or [16, 64]:
    x = torch.randn(B, 4, 32, 32, device="cuda")
    hint = (torch.rand(B, 1, 1024, 1024, device="cuda") < 0.05).float().expand(B, 3, 1024, 1024).contiguous()
    t = torch.tensor([500], device="cuda")
    with torch.no_grad():
        torch.cuda.synchronize()
        e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record()
        m(x, t, hint)                       # includes the hint pyramid (cached afterwards)
        e1.record()
        for _ in range(3):
            m(x, t, hint)
        e2.record()
        torch.cuda.synchronize()
    first, step = e0.elapsed_time(e1), e1.elapsed_time(e2) / 3
    gf = 64.64 * B
    print(f"B={B}: first call (hint pyramid + step) {first:.1f} ms, cached-hint step {step:.2f} ms = "
          f"{gf / step:.1f} TFLOP/s model, {B / step * 1e3:.1f} sample-steps/s, "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB  flag={rt.lib().cnb_tc_error_flag()}", flush=True)
    if os.environ.get("LDM_PROFILE", "0") == "1":
        import opprof
        opprof.table(lambda: m(x, t, hint), title=f"CelebHQ LDM ControlNet step B={B}")
    del x, hint
    torch.cuda.empty_cache()
