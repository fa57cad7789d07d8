"""Development micro-benchmark (GPU box; not collected by pytest): times individual hot shapes of the MNIST
ControlNet step at batch 1024 with CUDA events so a kernel change can be judged in one short gpurun call.
    python tests/conv_bench.py [conv|attn|gn|all] [reps]"""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("controlnet-pytorch_b200.ops")
rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = int(os.environ.get("CB_BATCH", "1024"))
B_LDM = int(os.environ.get("CB_BATCH_LDM", "256"))


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


if what in ("conv", "all"):
    shapes = [("3x3", 256, 256, 7), ("3x3", 128, 128, 7), ("3x3", 128, 128, 14), ("3x3", 64, 64, 28),
              ("3x3", 64, 128, 14), ("3x3", 32, 64, 28), ("3x3", 16, 16, 28), ("3x3", 64, 16, 28),
              ("1x1", 64, 192, 28), ("1x1", 64, 64, 28), ("1x1", 256, 768, 7), ("1x1", 256, 256, 7),
              ("4x4s2", 64, 64, 28), ("3x3", 128, 32, 28), ("1x1", 128, 384, 14), ("1x1", 128, 384, 7),
              ("1x1", 128, 128, 14), ("1x1", 128, 128, 7)]
    only = os.environ.get("CB_ONLY")
    if only:
        shapes = [shapes[int(i)] for i in only.split(",")]
    variants = [("f16", rt.MODE_F16, True), ("tf32", rt.MODE_F16, False)]
    if os.environ.get("CB_VARIANT"):
        variants = [v for v in variants if v[0] == os.environ["CB_VARIANT"]]
    if os.environ.get("CB_FP32", "0") == "1":
        variants.append(("fp32", rt.MODE_F32, False))
    for name, mode, half in variants:
        for kind, cin, cout, hw in shapes:
            k = {"3x3": 9, "1x1": 1, "4x4s2": 16}[kind]
            if half and (cin % 8 or cout % 16):
                continue
            x = torch.randn(B, hw, hw, cin, device="cuda")
            w = torch.randn(cout, k, cin, device="cuda") / math.sqrt(k * cin)
            wl = ops.cast_f16(w) if half else None
            if half:
                x = x.half()
            bias = torch.randn(cout, device="cuda")
            oh = ops.out_size(kind, hw)
            out = torch.empty(B, oh, oh, cout, device="cuda", dtype=torch.float16 if (half and os.environ.get("CB_OUT16", "1") == "1") else torch.float32)
            res = torch.randn_like(out) if os.environ.get("CB_RES", "0") == "1" else None
            ms = timeit(lambda: ops.conv(x, w, kind, cout, bias=bias, out=out, mode=mode, weight_lp=wl, residual=res,
                                      act=int(os.environ.get('CB_ACT', '0'))))
            fl = 2.0 * B * oh * oh * k * cin * cout
            by = x.numel() * x.element_size() + out.numel() * out.element_size() + w.numel() * (2 if half else 4)
            print(f"conv[{name}] {kind} {cin}->{cout} @{hw}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s  "
                  f"{by / ms / 1e6:8.1f} GB/s(min-traffic)", flush=True)

if what in ("attn", "all"):
    for mode in ((rt.MODE_F16, rt.MODE_F32) if os.environ.get("CB_FP32", "0") == "1" else (rt.MODE_F16,)):
        ashapes = [(784, 64, 4), (784, 16, 4), (196, 128, 4), (196, 32, 4), (49, 256, 4), (49, 128, 4), (49, 64, 4),
                   (1024, 384, 16), (1024, 128, 16), (256, 512, 16), (64, 768, 16)]   # last four: CelebHQ LDM levels
        if os.environ.get("CB_SHAPES"):                        # "L,E,heads[,batch];..." (batch defaults to CB_BATCH)
            ashapes = [tuple(int(v) for v in t.split(",")) for t in os.environ["CB_SHAPES"].split(";")]
        if os.environ.get("CB_ONLY_ATTN", os.environ.get("CB_ONLY")):
            ashapes = [ashapes[int(i)] for i in os.environ.get("CB_ONLY_ATTN", os.environ.get("CB_ONLY")).split(",")]
        for shp in ashapes:
            L, E, heads = shp[:3]
            side = int(math.isqrt(L))
            Bq = shp[3] if len(shp) > 3 else (B_LDM if heads == 16 else B)
            qkv = torch.randn(Bq, side, side, 3 * E, device="cuda")
            if os.environ.get("CB_ATTN_F16", "1") == "1" and mode != rt.MODE_F32:
                qkv = qkv.half()
            ms = timeit(lambda: ops.attention(qkv, heads, mode=mode, kernel=os.environ.get('CB_ATTN_KERNEL') or None))
            fl = 4.0 * Bq * L * L * E
            print(f"attn[{rt.mode_name(mode)}] L={L} E={E} d={E // heads}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.2f} TFLOP/s  "
                  f"{Bq * heads * L * L / ms / 1e6:8.2f} Gscore/s", flush=True)

if what in ("gn", "all"):
    for C, hw, G in [(64, 28, 8), (32, 28, 8), (16, 28, 8), (128, 14, 8), (256, 7, 8), (128, 7, 8)]:
        x = torch.randn(B, hw, hw, C, device="cuda")
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        for in16, f16 in ((False, False), (False, True), (True, True)):
            xi = x.half() if in16 else x
            if os.environ.get('CB_GN16', '0') == '1' and not (in16 and f16):
                continue
            ms = timeit(lambda: ops.groupnorm(xi, g, b, G, True, out_f16=f16))
            by = ((2.0 if in16 else 4.0) + (2.0 if f16 else 4.0)) * x.numel()
            print(f"gn C={C} @{hw} in={'f16' if in16 else 'f32'} out={'f16' if f16 else 'f32'}: {ms * 1e3:8.1f} us  "
                  f"{by / ms / 1e6:8.1f} GB/s", flush=True)
if what in ("elt", "all"):
    # scheduler update / Philox noise at sizes where HBM bandwidth shows (in situ the 3 MB tensors of a step do not)
    coef = torch.tensor([0.9, 0.43, 0.01, 0.995, 0.09, 1.0], device="cuda")
    for n_mi in (1, 16, 64, 192):
        n = n_mi << 20
        xt, eps, z = (torch.randn(n, device="cuda") for _ in range(3))
        prev, x0 = torch.empty_like(xt), torch.empty_like(xt)
        for name, kw, by in (("z given, x0 out", dict(z=z, x0_out=x0), 20), ("philox, x0 out", dict(x0_out=x0), 16),
                             ("philox, no x0", dict(want_x0=False), 12)):
            ms = timeit(lambda: ops.sched_step(xt, eps, coef, out=prev, seed=3, step=7, **kw))
            print(f"sched_step n={n_mi}Mi [{name}]: {ms * 1e3:8.1f} us  {by * n / ms / 1e6:8.1f} GB/s", flush=True)
        ms = timeit(lambda: ops.philox_normal((n,), "cuda", 3, 9))
        print(f"philox_normal n={n_mi}Mi: {ms * 1e3:8.1f} us  {4 * n / ms / 1e6:8.1f} GB/s (write only)", flush=True)
        del xt, eps, z, prev, x0
print("flag", rt.lib().cnb_tc_error_flag())
