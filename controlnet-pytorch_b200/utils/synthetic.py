"""Deterministic synthetic weights / inputs shared by tests, smoke() and bench.py.

There is no network for datasets or checkpoints, so every parity vector in this repo is built from
weights that are a pure function of (state_dict key, shape, seed).  The same fill is applied to the
reference modules (when generating tests/golden/*.npz in the build container) and to the drop-in
modules on the GPU box, so both sides hold bit-identical fp32 parameters without shipping checkpoints.

Default initialisation would leave every zero-conv (reference models/controlnet.py:7-10,85-107) at
exactly zero and hide the whole control branch from parity tests (SURVEY.md section 4 "trap"); this
fill makes them live.
"""
import zlib

import numpy as np
import torch


def _rng(key, seed):
    return np.random.default_rng([zlib.crc32(key.encode("utf-8")), int(seed)])


def det_tensor(key, shape, seed=0):
    """fp32 tensor that depends only on (key, shape, seed)."""
    shape = tuple(int(s) for s in shape)
    g = _rng(key, seed)
    n = int(np.prod(shape)) if len(shape) else 1
    v = g.standard_normal(n, dtype=np.float32)
    if len(shape) >= 2:
        if "up_sample_conv" in key:            # ConvTranspose2d weight (in, out, kh, kw), stride 2
            fan_in = shape[0] * 4
        else:
            fan_in = int(np.prod(shape[1:]))
        v *= np.float32(1.0 / np.sqrt(fan_in))
    elif key.endswith("weight"):               # GroupNorm scale
        v = np.float32(1.0) + np.float32(0.1) * v
    else:                                      # every bias
        v *= np.float32(0.05)
    return torch.from_numpy(v.reshape(shape).astype(np.float32))


def det_state_dict(template, seed=0):
    """Return {key: deterministic tensor} for every floating entry of ``template`` (a state_dict)."""
    out = {}
    for k, v in template.items():
        if torch.is_floating_point(v):
            out[k] = det_tensor(k, v.shape, seed)
        else:
            out[k] = v.clone()
    return out


def det_noise(key, shape, seed=0):
    g = _rng("noise:" + key, seed)
    return torch.from_numpy(g.standard_normal(tuple(shape), dtype=np.float32))


def det_hint(batch, size, p=0.1, seed=0, channels=3):
    """Canny-like hint: Bernoulli(p) in {0,1}, replicated over channels
    (reference dataset/mnist_dataset.py:56-63: cv2.Canny -> 3 identical channels -> ToTensor)."""
    g = _rng("hint", seed)
    m = (g.random((batch, 1, size, size)) < p).astype(np.float32)
    return torch.from_numpy(np.repeat(m, channels, axis=1).copy())


MNIST_PARAMS = dict(im_channels=1, im_size=28, hint_channels=3, down_channels=[32, 64, 128, 256],
                    mid_channels=[256, 256, 128], down_sample=[True, True, False], time_emb_dim=128,
                    num_down_layers=2, num_mid_layers=2, num_up_layers=2, num_heads=4)

CIFAR_PARAMS = dict(im_channels=3, in_channels=3, im_size=32, hint_channels=3,
                    down_channels=[64, 128, 256, 512], mid_channels=[512, 512, 256],
                    down_sample=[True, True, False], time_emb_dim=128, num_down_layers=2,
                    num_mid_layers=2, num_up_layers=2, num_heads=4, sigma_data=1.0, sigma_min=0.002,
                    sigma_max=5.0)

TINY_PARAMS = dict(im_channels=1, im_size=16, hint_channels=3, down_channels=[16, 32, 64, 64],
                   mid_channels=[64, 64, 64], down_sample=[True, True, False], time_emb_dim=32,
                   num_down_layers=1, num_mid_layers=1, num_up_layers=1, num_heads=4)

CELEBHQ_LDM_PARAMS = dict(hint_channels=3, down_channels=[256, 384, 512, 768], mid_channels=[768, 512],
                          down_sample=[True, True, True], attn_down=[True, True, True], time_emb_dim=512,
                          norm_channels=32, num_heads=16, conv_out_channels=128, num_down_layers=2,
                          num_mid_layers=2, num_up_layers=2)

TINY_LDM_PARAMS = dict(hint_channels=3, down_channels=[32, 64, 64, 128], mid_channels=[128, 64],
                       down_sample=[True, True, True], attn_down=[True, False, True], time_emb_dim=64,
                       norm_channels=8, num_heads=4, conv_out_channels=32, num_down_layers=1,
                       num_mid_layers=1, num_up_layers=1)

# config/celebhq.yaml:27-38 (im_channels 3, im_size 128 -> latent 4 x 32 x 32)
CELEBHQ_VAE_PARAMS = dict(z_channels=4, codebook_size=8192, down_channels=[128, 256, 384], mid_channels=[384],
                          down_sample=[True, True], attn_down=[False, False], norm_channels=32, num_heads=4,
                          num_down_layers=2, num_mid_layers=2, num_up_layers=2)

# small VAE that exercises what the CelebHQ one skips: a decoder / encoder MidBlock and an attention UpBlock
TINY_VAE_PARAMS = dict(z_channels=4, codebook_size=16, down_channels=[16, 32, 64], mid_channels=[64, 64],
                       down_sample=[True, True], attn_down=[False, True], norm_channels=8, num_heads=4,
                       num_down_layers=1, num_mid_layers=1, num_up_layers=1)

MNIST_DIFFUSION = dict(num_timesteps=1000, beta_start=0.0001, beta_end=0.02)
CELEBHQ_DIFFUSION = dict(num_timesteps=1000, beta_start=0.0015, beta_end=0.0195)
