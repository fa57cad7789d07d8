"""Development helper (GPU box): CUDA events around every libcnb200 launch made by a callable, aggregated per layer
shape.  Used by tests/vae_check.py and tests/ldm_check.py; tests/step_profile.py is the MNIST-step original."""
import collections
import importlib

import torch

ops = importlib.import_module("controlnet-pytorch_b200.ops")
NAMES = ("conv", "groupnorm", "attention", "sched_step", "copy_channels", "linear_small", "nchw_to_nhwc",
         "nhwc_to_nchw", "time_embedding")


def table(fn, reps=2, title=""):
    rec = []
    orig = {k: getattr(ops, k) for k in NAMES}

    def wrap(name, f):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = f(*a, **k)
            e1.record()
            t = a[0]
            if name == "conv":
                kind = a[2] if a[2] is not None else "convT"
                cin = t.shape[3] - k.get('in_coff', 0)
                key = f"conv {kind:6s} {cin:4d}->{a[3]:4d} @{t.shape[1]:4d} in={str(t.dtype)[6:]:8s} out={str(out.dtype)[6:]:8s}"
                taps = a[1].shape[-2] if a[1].dim() == 3 else 4
                oh, ow = (t.shape[1], t.shape[2]) if a[2] is None else (out.shape[1], out.shape[2])
                flops = 2.0 * t.shape[0] * oh * ow * taps * cin * a[3]
                byts = t.numel() * t.element_size() + t.shape[0] * oh * ow * a[3] * out.element_size()
            elif name == "groupnorm":
                key = f"gn   C={t.shape[3]:4d} @{t.shape[1]:4d} in={str(t.dtype)[6:]:8s} out={'f16' if k.get('out_f16') else 'f32'}"
                flops, byts = 0.0, t.numel() * (t.element_size() + (2.0 if k.get('out_f16') else 4.0))
            elif name == "attention":
                L = t.shape[1] * t.shape[2]
                key = f"attn L={L:4d} E={t.shape[3] // 3:4d} {str(t.dtype)[6:]}"
                flops, byts = 4.0 * t.shape[0] * L * L * (t.shape[3] // 3), t.numel() * t.element_size() * 4 / 3
            else:
                key, flops, byts = name, 0.0, 0.0
            rec.append((key, flops, byts, e0, e1))
            return out
        return w

    for k_, v in orig.items():
        setattr(ops, k_, wrap(k_, v))
    try:
        with torch.no_grad():
            fn()
            rec.clear()
            for _ in range(reps):
                fn()
        torch.cuda.synchronize()
    finally:
        for k_, v in orig.items():
            setattr(ops, k_, v)
    agg = collections.OrderedDict()
    for key, fl, by, e0, e1 in rec:
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += fl
        a[3] += by
    tot = sum(a[1] for a in agg.values())
    print(f"{title}: {tot / reps:.3f} ms in libcnb200 launches ({sum(a[0] for a in agg.values()) // reps} calls)")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ms = a[1] / reps
        print(f"{key:62s} n={a[0] // reps:3d} {ms * 1e3:9.1f} us {100 * a[1] / tot:5.1f}%  "
              f"{a[2] / a[1] / 1e9:8.1f} TF/s {a[3] / a[1] / 1e6:8.1f} GB/s  ({ms * 1e3 / (a[0] // reps):7.1f} us each)")
