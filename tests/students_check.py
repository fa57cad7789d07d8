"""Development check for BASELINE configs 4 and 5 (GPU box; not collected by pytest): single-step throughput of the
distilled students.  Config 4: consistency student, MNIST shapes, batch 4096 (sigma = 80) and the CelebHQ-latent variant
(dict(ldm_params, im_channels=4, im_size=32)); config 5: DM student on CIFAR 32x32, batch sweep 1..8192.
Times the public forward (NCHW fp32 in, NCHW fp32 out, hint feature cached after the first call)."""
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
rt.set_mode("f16")


def fill(m):
    m.load_state_dict(syn.det_state_dict(m.state_dict(), 0))
    return m.cuda().eval()


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def hints(B, s):
    return (torch.rand(B, 1, s, s, device="cuda") < 0.1).float().expand(B, 3, s, s).contiguous()


with torch.no_grad():
    CS = importlib.import_module("controlnet-pytorch_b200.models.consistency_controlnet_distilled").ConsistencyControlNet
    DM = importlib.import_module("controlnet-pytorch_b200.models.distribution_matching_controlnet").DistributionMatchingControlNet
    m = fill(CS(syn.MNIST_PARAMS))
    for B in (1024, 4096):
        x, h, sg = torch.randn(B, 1, 28, 28, device="cuda"), hints(B, 28), torch.full((B,), 80.0, device="cuda")
        ms = timeit(lambda: m(x, sg, h))
        print(f"consistency MNIST  B={B:5d}: {ms:8.2f} ms  {B / ms * 1e3:10.0f} samples/s  {2.119 * B / ms:7.1f} TFLOP/s model", flush=True)
    del m
    lat = dict(syn.CELEBHQ_LDM_PARAMS, im_channels=4, im_size=32)
    m = fill(CS(lat))
    for B in (256, 1024):
        x, h, sg = torch.randn(B, 4, 32, 32, device="cuda"), hints(B, 32), torch.full((B,), 80.0, device="cuda")
        ms = timeit(lambda: m(x, sg, h), reps=3)
        print(f"consistency CelebHQ-latent B={B:5d}: {ms:8.2f} ms  {B / ms * 1e3:10.0f} samples/s  {33.85 * B / ms:7.1f} TFLOP/s model", flush=True)
    del m
    torch.cuda.empty_cache()
    m = fill(DM(syn.CIFAR_PARAMS))
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
        x, h, t = torch.randn(B, 3, 32, 32, device="cuda"), hints(B, 32), torch.full((B,), 999, device="cuda")
        ms = timeit(lambda: m(x, t, h), reps=5 if B <= 1024 else 2)
        print(f"DM CIFAR           B={B:5d}: {ms:8.2f} ms  {B / ms * 1e3:10.0f} samples/s  {9.394 * B / ms:7.1f} TFLOP/s model", flush=True)
    S = importlib.import_module("controlnet-pytorch_b200.sampler")
    g = S.GraphedStudent(m)
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        x, h, t = torch.randn(B, 3, 32, 32, device="cuda"), hints(B, 32), torch.full((B,), 999, device="cuda")
        ms = timeit(lambda: g(x, t, h), reps=10)
        print(f"DM CIFAR (graph)   B={B:5d}: {ms:8.3f} ms  {B / ms * 1e3:10.0f} samples/s  {9.394 * B / ms:7.1f} TFLOP/s model", flush=True)
print("flag", rt.lib().cnb_tc_error_flag(), "mem GiB", torch.cuda.max_memory_allocated() / 2**30)
