"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Weights are utils/synthetic.det_state_dict fills of the reference modules' own state_dict keys, inputs are
det_noise / det_hint, so the vectors can be reproduced on any box from (key, shape, seed) alone.  Only the
OUTPUTS of the reference are stored.  TEST INFRASTRUCTURE ONLY.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CNB_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
syn = importlib.import_module("controlnet-pytorch_b200.utils.synthetic")
sys.path.insert(0, REF)

from models.controlnet import ControlNet as RefControlNet                       # noqa: E402
from models.controlnet_ldm import ControlNet as RefControlNetLDM                # noqa: E402
from models.unet_base import Unet as RefUnet                                    # noqa: E402
from models.consistency_controlnet_distilled import ConsistencyControlNet as RefCons       # noqa: E402
from models.distribution_matching_controlnet import DistributionMatchingControlNet as RefDM  # noqa: E402
from scheduler.linear_noise_scheduler import LinearNoiseScheduler as RefSched   # noqa: E402
import scheduler.linear_noise_scheduler as ref_sched_mod                        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fill(model, seed=0):
    model.load_state_dict(syn.det_state_dict(model.state_dict(), seed))
    return model.eval()


def inputs(name, B, C, S, hint_size=None, p=0.1):
    x = syn.det_noise(name + ":x", (B, C, S, S))
    hint = syn.det_hint(B, hint_size or S, p=p)
    return x, hint


class _InjectZ:
    """Replace torch.randn inside the reference scheduler module so z is the injected tensor
    (linear_noise_scheduler.py:71 draws it from the CPU default generator)."""

    def __init__(self, zs):
        self.zs = list(zs)

    def __enter__(self):
        self._orig = ref_sched_mod.torch.randn
        outer = self

        def fake(*a, **k):
            return outer.zs.pop(0)
        ref_sched_mod.torch.randn = fake
        return self

    def __exit__(self, *a):
        ref_sched_mod.torch.randn = self._orig


def trajectory(model, sched, x, hint, steps, name):
    zs = [syn.det_noise(f"{name}:z{k}", tuple(x.shape)) for k in range(steps)]
    xt = x
    with _InjectZ(zs[:steps - 1]):
        for t in reversed(range(steps)):
            eps = model(xt, torch.as_tensor(t).unsqueeze(0), hint)
            xt, x0 = sched.sample_prev_timestep(xt, eps, torch.as_tensor(t))
    return xt, x0


@torch.no_grad()
def make_config1():
    """BASELINE.json configs[0]: MNIST DDPM ControlNet, batch 16, 50 steps (t = 49 .. 0, the semantics of
    tools/compare_all_controlnet_models.py) with the per-step z injected; the final x_0 / x_{t-1} of
    the unmodified reference are the fixture the PSNR / max-abs bound of the parity tests is stated against."""
    cfg = syn.MNIST_PARAMS
    m = fill(RefControlNet(cfg))
    x, hint = inputs("config1", 16, cfg["im_channels"], cfg["im_size"])
    sched = RefSched(**syn.MNIST_DIFFUSION)
    xt, x0 = trajectory(m, sched, x, hint, 50, "config1")
    np.savez_compressed(os.path.join(OUT, "config1_mnist_b16_50step.npz"), xt=xt.numpy(), x0=x0.numpy())
    print("config1", tuple(xt.shape), float(x0.abs().max()), float(x0.std()))


@torch.no_grad()
def make_teachers():
    """Teacher-side inference of the distillation wrappers: get_teacher_prediction with per-sample t
    and get_ddpm_teacher_prediction / sigma_to_timestep; the teacher checkpoint is a det_state_dict ControlNet."""
    import tempfile
    from models.distribution_matching_controlnet import DistributionMatchingControlNetDistilled as RefDMD
    from models.consistency_controlnet_distilled import ConsistencyControlNetDistilled as RefConsD
    cfg = syn.TINY_PARAMS
    teacher = fill(RefControlNet(cfg), seed=3)
    with tempfile.TemporaryDirectory() as d:
        ck = os.path.join(d, "teacher.pth")
        torch.save(teacher.state_dict(), ck)
        dm = RefDMD(cfg, ck, device=torch.device("cpu")).eval()
        cs = RefConsD(cfg, ck, device=torch.device("cpu")).eval()
    x, hint = inputs("teacher_tiny", 3, cfg["im_channels"], cfg["im_size"])
    t = torch.tensor([999, 412, 3])
    sig = torch.tensor([60.0, 2.5, 0.05])
    rec = {"dm_x0": dm.get_teacher_prediction(x, t, hint).numpy(),
           "dm_x0_shared_t": dm.get_teacher_prediction(x[:1], torch.tensor(77), hint[:1]).numpy(),
           "cs_t": cs.sigma_to_timestep(sig).numpy(),
           "cs_x0": cs.get_ddpm_teacher_prediction(x, sig, hint).numpy()}
    np.savez_compressed(os.path.join(OUT, "teachers_tiny.npz"), **rec)
    print("teachers", {k: v.shape for k, v in rec.items()})


@torch.no_grad()
def make_vae():
    """models/vae.py decode / encode on the tiny VAE (decoder MidBlock + attention UpBlock) and, at batch 1, on the
    CelebHQ autoencoder_params (config/celebhq.yaml:27-38): output statistics + a strided sample of the image."""
    from models.vae import VAE as RefVAE
    cfg = syn.TINY_VAE_PARAMS
    m = fill(RefVAE(3, cfg))
    z = syn.det_noise("vae_tiny:z", (2, 4, 8, 8))
    x = syn.det_noise("vae_tiny:x", (2, 3, 32, 32))
    rec = {"dec": m.decode(z).numpy()}
    noise = syn.det_noise("vae_tiny:n", (2, 4, 8, 8))
    import models.vae as ref_vae_mod
    orig = ref_vae_mod.torch.randn
    ref_vae_mod.torch.randn = lambda *a, **k: noise
    try:
        sample, enc = m.encode(x)
    finally:
        ref_vae_mod.torch.randn = orig
    rec["enc_out"], rec["enc_sample"] = enc.numpy(), sample.numpy()
    np.savez_compressed(os.path.join(OUT, "vae_tiny.npz"), **rec)
    m = fill(RefVAE(3, syn.CELEBHQ_VAE_PARAMS))
    z = syn.det_noise("vae_celebhq:z", (1, 4, 32, 32))
    img = m.decode(z)
    np.savez_compressed(os.path.join(OUT, "vae_celebhq.npz"), dec_strided=img[:, :, ::4, ::4].numpy(),
                        dec_sum=img.double().sum().numpy(), dec_sqsum=(img.double() ** 2).sum().numpy())
    import json
    with open(os.path.join(OUT, "state_dict_manifest_vae.json"), "w") as f:
        json.dump({"vae_tiny": {k: list(v.shape) for k, v in RefVAE(3, cfg).state_dict().items()},
                   "vae_celebhq": {k: list(v.shape) for k, v in m.state_dict().items()}}, f)
    print("vae", img.shape)


@torch.no_grad()
def make_fullsize():
    """Full-size BASELINE configs at batch 1: the 190 M-parameter CelebHQ LDM ControlNet eps
    (config/celebhq.yaml ldm_params, hint 3 x 1024 x 1024, down_sample_factor 32), the CelebHQ-latent consistency student
    (dict(ldm_params, im_channels=4, im_size=32)), and the multi-step generate() of the consistency
    wrapper (consistency_controlnet_distilled.py:390-409) with its noise draws injected."""
    cfg = syn.CELEBHQ_LDM_PARAMS
    m = fill(RefControlNetLDM(4, cfg, down_sample_factor=32))
    x, hint = inputs("celebhq_ldm", 1, 4, 32, hint_size=1024, p=0.05)
    rec = {f"eps_{t}": m(x, torch.as_tensor(t).unsqueeze(0), hint).numpy() for t in (500, 3)}
    np.savez_compressed(os.path.join(OUT, "controlnet_celebhq_ldm.npz"), **rec)
    print("celebhq ldm", {k: (v.shape, float(np.abs(v).max())) for k, v in rec.items()})
    del m
    lat = dict(cfg, im_channels=4, im_size=32)
    m = fill(RefCons(lat))
    x, hint = inputs("cons_celebhq_latent", 1, 4, 32)
    rec = {"x0_max": m(x, torch.full((1,), 80.0), hint).numpy(), "x0_mid": m(x, torch.full((1,), 1.7), hint).numpy()}
    np.savez_compressed(os.path.join(OUT, "consistency_celebhq_latent.npz"), **rec)
    print("celebhq-latent student", {k: v.shape for k, v in rec.items()})
    del m
    # multi-step generate: x_T and the re-noising draws come from torch.randn / randn_like (:383, :398) - injected
    from models.consistency_controlnet_distilled import ConsistencyControlNetDistilled as RefConsD
    import models.consistency_controlnet_distilled as ref_cd_mod
    cfg = syn.TINY_PARAMS
    w = RefConsD(cfg).eval()
    w.student.load_state_dict(syn.det_state_dict(w.student.state_dict(), 0))
    hint = syn.det_hint(2, 16)
    shape = (2, 1, 16, 16)
    rec = {}
    for steps in (1, 4):
        draws = [syn.det_noise(f"gen{steps}:n{k}", shape) for k in range(steps)]
        orig_randn, orig_like = ref_cd_mod.torch.randn, ref_cd_mod.torch.randn_like
        ref_cd_mod.torch.randn = lambda *a, **k: draws.pop(0)
        ref_cd_mod.torch.randn_like = lambda *a, **k: draws.pop(0)
        try:
            rec[f"gen_{steps}"] = w.generate(hint, shape, num_steps=steps).numpy()
        finally:
            ref_cd_mod.torch.randn, ref_cd_mod.torch.randn_like = orig_randn, orig_like
        assert not draws
    np.savez_compressed(os.path.join(OUT, "consistency_generate_tiny.npz"), **rec)
    print("generate", {k: (v.shape, float(np.abs(v).max())) for k, v in rec.items()})


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if "--only-vae" in sys.argv:
        return make_vae()
    if "--only-teachers" in sys.argv:
        return make_teachers()
    if "--only-config1" in sys.argv:
        return make_config1()
    if "--only-fullsize" in sys.argv:
        return make_fullsize()

    # ---- DDPM ControlNet: tiny / mnist / cifar
    for name, cfg, B, ts in (("tiny", syn.TINY_PARAMS, 2, (999, 37, 0)),
                             ("mnist", syn.MNIST_PARAMS, 2, (999, 500, 0)),
                             ("cifar", syn.CIFAR_PARAMS, 1, (500,))):
        m = fill(RefControlNet(cfg))
        x, hint = inputs(name, B, cfg["im_channels"], cfg["im_size"])
        rec = {"x_sum": x.double().sum().numpy(), "hint_sum": hint.double().sum().numpy()}
        for t in ts:
            rec[f"eps_{t}"] = m(x, torch.as_tensor(t).unsqueeze(0), hint).numpy()
        if name != "cifar":
            sched = RefSched(**syn.MNIST_DIFFUSION)
            xt, x0 = trajectory(m, sched, x, hint, 3, name)
            rec["traj3_xt"], rec["traj3_x0"] = xt.numpy(), x0.numpy()
        # per-sample t of shape (B,) as used by the training / compare tools
        if name == "tiny":
            rec["eps_pers"] = m(x, torch.tensor([10, 700]), hint).numpy()
        np.savez_compressed(os.path.join(OUT, f"controlnet_{name}.npz"), **rec)
        print("controlnet", name, {k: getattr(v, "shape", None) for k, v in rec.items()})

    # ---- plain U-Net forward (tools/sample_ddpm.py path)
    m = fill(RefUnet(syn.TINY_PARAMS))
    x, _ = inputs("unet_tiny", 2, 1, 16)
    np.savez_compressed(os.path.join(OUT, "unet_tiny.npz"),
                        eps_123=m(x, torch.as_tensor(123).unsqueeze(0)).numpy())

    # ---- LDM ControlNet, tiny config (hint 64x64 -> latent 8x8)
    cfg = syn.TINY_LDM_PARAMS
    m = fill(RefControlNetLDM(4, cfg, down_sample_factor=8))
    x, hint = inputs("tiny_ldm", 2, 4, 8, hint_size=64, p=0.05)
    rec = {}
    for t in (999, 3):
        rec[f"eps_{t}"] = m(x, torch.as_tensor(t).unsqueeze(0), hint).numpy()
    sched = RefSched(ldm_scheduler=True, **syn.CELEBHQ_DIFFUSION)
    xt, x0 = trajectory(m, sched, x, hint, 3, "tiny_ldm")
    rec["traj3_xt"], rec["traj3_x0"] = xt.numpy(), x0.numpy()
    np.savez_compressed(os.path.join(OUT, "controlnet_tiny_ldm.npz"), **rec)

    # ---- students
    for name, cfg, sigma in (("tiny", syn.TINY_PARAMS, 80.0), ("mnist", syn.MNIST_PARAMS, 80.0),
                             ("cifar", syn.CIFAR_PARAMS, 5.0)):
        B = 2 if name != "cifar" else 1
        m = fill(RefCons(cfg))
        x, hint = inputs("cons_" + name, B, cfg["im_channels"], cfg["im_size"])
        rec = {"x0_max": m(x, torch.full((B,), sigma), hint).numpy(),
               "x0_mid": m(x, torch.full((B,), 1.7), hint).numpy(),
               "x0_min": m(x, torch.full((B,), 0.001), hint).numpy()}
        np.savez_compressed(os.path.join(OUT, f"consistency_{name}.npz"), **rec)
        m = fill(RefDM(cfg))
        x, hint = inputs("dm_" + name, B, cfg["im_channels"], cfg["im_size"])
        rec = {"x0_999": m(x, torch.full((B,), 999), hint).numpy()}
        if B == 2:
            rec["x0_pers"] = m(x, torch.tensor([5, 400]), hint).numpy()
        np.savez_compressed(os.path.join(OUT, f"dm_{name}.npz"), **rec)
        print("students", name)

    # ---- scheduler: tables + steps
    for name, kw in (("ddpm", dict(syn.MNIST_DIFFUSION)),
                     ("ldm", dict(syn.CELEBHQ_DIFFUSION, ldm_scheduler=True))):
        s = RefSched(**kw)
        rec = dict(betas=s.betas.numpy(), alphas=s.alphas.numpy(), alpha_cum_prod=s.alpha_cum_prod.numpy(),
                   sqrt_alpha_cum_prod=s.sqrt_alpha_cum_prod.numpy(),
                   sqrt_one_minus_alpha_cum_prod=s.sqrt_one_minus_alpha_cum_prod.numpy())
        xt = syn.det_noise("sched:xt", (3, 4, 9, 7))
        eps = syn.det_noise("sched:eps", (3, 4, 9, 7))
        z = syn.det_noise("sched:z", (3, 4, 9, 7))
        for t in (999, 500, 1, 0):
            with _InjectZ([z]):
                a, b = s.sample_prev_timestep(xt, eps, torch.as_tensor(t))
            rec[f"prev_{t}"], rec[f"x0_{t}"] = a.numpy(), b.numpy()
        np.savez_compressed(os.path.join(OUT, f"scheduler_{name}.npz"), **rec)
    # ---- state_dict layout manifest (key -> shape) of every reference module on the path
    import json
    from models.unet_cond_base import Unet as RefUnetLDM
    from models.consistency_controlnet_distilled import ConsistencyControlNetDistilled as RefConsD
    from models.distribution_matching_controlnet import DistributionMatchingControlNetDistilled as RefDMD
    man = {}
    for tag, mod in (("controlnet_mnist", RefControlNet(syn.MNIST_PARAMS)),
                     ("unet_mnist", RefUnet(syn.MNIST_PARAMS)),
                     ("unet_mnist_noup", RefUnet(syn.MNIST_PARAMS, use_up=False)),
                     ("controlnet_ldm_tiny", RefControlNetLDM(4, syn.TINY_LDM_PARAMS, down_sample_factor=8)),
                     ("unet_ldm_tiny", RefUnetLDM(4, syn.TINY_LDM_PARAMS)),
                     ("consistency_mnist", RefCons(syn.MNIST_PARAMS)),
                     ("dm_mnist", RefDM(syn.MNIST_PARAMS)),
                     ("consistency_distilled_tiny", RefConsD(syn.TINY_PARAMS)),
                     ("dm_distilled_tiny", RefDMD(syn.TINY_PARAMS, "missing.pth", device=None))):
        man[tag] = {k: list(v.shape) for k, v in mod.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_manifest.json"), "w") as f:
        json.dump(man, f)
    make_vae()
    make_teachers()
    make_config1()
    make_fullsize()
    print("done ->", OUT)


if __name__ == "__main__":
    main()
