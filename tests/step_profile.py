"""Development tool (GPU box; not collected by pytest): per-layer-shape time table of one eager MNIST ControlNet
denoising step at batch B, CUDA events around every libcnb200 launch.   python tests/step_profile.py [batch]"""
import collections
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ops = importlib.import_module("controlnet-pytorch_b200.ops")
rt = importlib.import_module("controlnet-pytorch_b200.runtime")
S = importlib.import_module("controlnet-pytorch_b200.sampler")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
rt.lib()
rt.set_mode(os.environ.get("CNB_MODE", "f16"))
cfg, model, sched, hint_host = bench.build_problem(B, dev)
hint = hint_host.to(dev)
x = torch.randn(B, 1, 28, 28, device=dev)
rec = []
orig = {k: getattr(ops, k) for k in ("conv", "groupnorm", "attention", "sched_step", "copy_channels", "linear_small",
                                     "nchw_to_nhwc", "nhwc_to_nchw", "time_embedding")}


def wrap(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        t = a[0]
        if name == "conv":
            kind = a[2] if a[2] is not None else "convT"
            key = f"conv {kind:6s} {t.shape[3] - k.get('in_coff', 0):4d}->{a[3]:4d} @{t.shape[1]:3d} in={str(t.dtype)[6:]:8s} " \
                  f"out={'f16' if k.get('out_f16') else 'f32'} res={'y' if k.get('residual') is not None else 'n'}"
            flops = 2.0 * t.shape[0] * out.shape[1] * out.shape[2] * (a[1].shape[-2] if a[1].dim() == 3 else 4) * \
                (t.shape[3] - k.get('in_coff', 0)) * a[3] / (4 if a[2] is None else 1)
            byts = t.numel() * t.element_size() + out.shape[0] * out.shape[1] * out.shape[2] * a[3] * \
                (2 if k.get('out_f16') else 4) * (2 if k.get('residual') is not None and not k.get('out_f16') else 1)
        elif name == "groupnorm":
            key = f"gn   C={t.shape[3]:4d} @{t.shape[1]:3d} out={'f16' if k.get('out_f16') else 'f32'}"
            flops, byts = 0.0, t.numel() * (6.0 if k.get('out_f16') else 8.0)
        elif name == "attention":
            key = f"attn L={t.shape[1] * t.shape[2]:4d} E={t.shape[3] // 3:4d} {str(t.dtype)[6:]}"
            flops, byts = 4.0 * t.shape[0] * (t.shape[1] * t.shape[2]) ** 2 * (t.shape[3] // 3), t.numel() * t.element_size() * 4 / 3
        else:
            key, flops, byts = name, 0.0, 0.0
        rec.append((key, flops, byts, e0, e1))
        return out
    return w


for k_, v in orig.items():
    setattr(ops, k_, wrap(k_, v))
smp = S.DDPMSampler(model, sched, seed=3, use_graph=False)
with torch.no_grad():
    smp.sample_eager(x, hint, steps=1)
    rec.clear()
    smp.sample_eager(x, hint, steps=2)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, fl, by, e0, e1 in rec:
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += fl
    a[3] += by
tot = sum(a[1] for a in agg.values())
print(f"B={B}: {tot / 2:.3f} ms per step in libcnb200 launches ({sum(a[0] for a in agg.values()) // 2} calls)")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ms = a[1] / 2
    print(f"{key:62s} n={a[0] // 2:3d} {ms * 1e3:9.1f} us {100 * a[1] / tot:5.1f}%  "
          f"{a[2] / a[1] / 1e9:8.1f} TF/s {a[3] / a[1] / 1e6:8.1f} GB/s  ({ms * 1e3 / (a[0] // 2):7.1f} us each)")
