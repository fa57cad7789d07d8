"""Shared helpers for the parity tests (model builders over a flat state_dict, metrics)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
syn = importlib.import_module("controlnet-pytorch_b200.utils.synthetic")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a, b, peak_to_peak=2.0):
    """PSNR in dB of images in [-1, 1] (peak-to-peak 2, as written by the sampling tools before the [0, 1] map)."""
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0.0 else 10.0 * float(np.log10(peak_to_peak ** 2 / mse))


def max_abs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs(name, B, C, S, hint_size=None, p=0.1):
    return syn.det_noise(name + ":x", (B, C, S, S)), syn.det_hint(B, hint_size or S, p=p)


_SD_CACHE = {}


def det_sd_for(module_factory, tag, seed=0):
    """state_dict with deterministic weights for a freshly built (CPU) module; cached per tag."""
    if tag not in _SD_CACHE:
        m = module_factory()
        _SD_CACHE[tag] = syn.det_state_dict(m.state_dict(), seed)
    return _SD_CACHE[tag]
