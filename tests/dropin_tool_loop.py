"""Run in a subprocess by the tests with PYTHONSAFEPATH=1 and PYTHONPATH=<repo>/controlnet-pytorch_b200/dropin: the
reference tool's own import lines and sampling loop (tools/sample_ddpm_controlnet.py:9-12, :43-51) resolve to the
B200 drop-ins.  `--check-imports` stops after the import checks (CPU); otherwise 3 steps run on cuda:0 with the
per-step z injected the way oracle/make_golden.py injects it into the reference, and the result is compared with the
reference's fixture tests/golden/controlnet_mnist.npz."""
import importlib
import os
import sys

import numpy as np
import torch

from models.controlnet import ControlNet                                       # the tool's import lines, verbatim
from scheduler.linear_noise_scheduler import LinearNoiseScheduler
from models.controlnet_ldm import ControlNet as ControlNetLDM
from models.vae import VAE
from models.consistency_controlnet_distilled import ConsistencyControlNetDistilled
from models.distribution_matching_controlnet import DistributionMatchingControlNetDistilled
from models.unet_base import Unet

PKG = "controlnet-pytorch_b200"
for cls, mod in ((ControlNet, "models.controlnet"), (LinearNoiseScheduler, "scheduler.linear_noise_scheduler"),
                 (ControlNetLDM, "models.controlnet_ldm"), (VAE, "models.vae"), (Unet, "models.unet_base"),
                 (ConsistencyControlNetDistilled, "models.consistency_controlnet_distilled"),
                 (DistributionMatchingControlNetDistilled, "models.distribution_matching_controlnet")):
    assert cls.__module__ == f"{PKG}.{mod}", (cls, cls.__module__)
    assert getattr(importlib.import_module(f"{PKG}.{mod}"), cls.__name__) is cls
if "--check-imports" in sys.argv:
    print("dropin imports ok")
    sys.exit(0)

syn = importlib.import_module(PKG + ".utils.synthetic")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "controlnet_mnist.npz"))
device = torch.device("cuda")
model_config = syn.MNIST_PARAMS
model = ControlNet(model_config).to(device)
model.load_state_dict(syn.det_state_dict(model.state_dict(), 0))
model.eval()
scheduler = LinearNoiseScheduler(**syn.MNIST_DIFFUSION)
x = syn.det_noise("mnist:x", (2, 1, 28, 28))
hints = syn.det_hint(2, 28).to(device)
zs = [syn.det_noise(f"mnist:z{k}", tuple(x.shape)) for k in range(3)]
real_randn = torch.randn
torch.randn = lambda *a, **k: zs.pop(0)          # linear_noise_scheduler.py:71 draws z on the CPU generator
try:
    xt = x.to(device)
    with torch.no_grad():
        for i in reversed(range(3)):
            noise_pred = model(xt, torch.as_tensor(i).unsqueeze(0).to(device), hints)
            xt, x0_pred = scheduler.sample_prev_timestep(xt, noise_pred, torch.as_tensor(i).to(device))
            ims = (torch.clamp(xt, -1., 1.).detach().cpu() + 1) / 2
finally:
    torch.randn = real_randn


def rel(a, b):
    a, b = a.double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm())


e_xt, e_x0 = rel(xt.cpu(), g["traj3_xt"]), rel(x0_pred.cpu(), g["traj3_x0"])
print(f"dropin tool loop: x_t-1 rel-L2 {e_xt:.3e}, x0 rel-L2 {e_x0:.3e}")
assert e_xt < 3e-2 and e_x0 < 3e-2, (e_xt, e_x0)
assert ims.shape == (2, 1, 28, 28) and len(zs) == 1      # z is not drawn at t == 0
