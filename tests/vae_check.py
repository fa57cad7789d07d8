"""Development check (GPU box; not collected by pytest): CelebHQ VAE.decode (config/celebhq.yaml autoencoder_params,
latent 4x32x32 -> image 3x128x128; 62.4 GFLOP per sample) timing at a few batches with a per-kernel table.
    python tests/vae_check.py [batches...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import syn  # noqa: E402

rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
VAE = importlib.import_module("controlnet-pytorch_b200.models.vae").VAE
m = VAE(3, syn.CELEBHQ_VAE_PARAMS)
m.load_state_dict(syn.det_state_dict(m.state_dict(), 0))
m = m.cuda().eval()
rt.set_mode("f16")
for B in [int(a) for a in sys.argv[1:]] or [16, 64]:
    z = torch.randn(B, 4, 32, 32, device="cuda")
    with torch.no_grad():
        m.decode(z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            img = m.decode(z)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"B={B}: decode {ms:.2f} ms = {62.42 * B / ms:.1f} TFLOP/s model, {B / ms * 1e3:.0f} images/s, "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB flag={rt.lib().cnb_tc_error_flag()} "
          f"finite={bool(torch.isfinite(img).all())}", flush=True)
    if os.environ.get("VAE_PROFILE", "1") == "1":
        import opprof
        opprof.table(lambda: m.decode(z), title=f"VAE.decode B={B}")
