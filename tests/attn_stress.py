"""Development aid (GPU box; not collected by pytest): repeat the tcgen05 / TMEM attention kernel on several shapes and
count runs that differ from the mma.sync kernel / from the first run (intermittent races show up as a non-zero count).
    python tests/attn_stress.py [repeats]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("controlnet-pytorch_b200.ops")
rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 10
shapes = [(5, 257, 256, 16), (64, 784, 64, 4), (2, 1024, 128, 16), (2, 784, 64, 4), (16, 784, 16, 4), (8, 196, 128, 4),
          (3, 130, 64, 4), (32, 196, 32, 4), (4, 1024, 384, 16)]
tot_bad = 0
for B, L, E, heads in shapes:
    torch.manual_seed(B * 1000 + L)
    qkv = torch.randn(B, L, 1, 3 * E, device="cuda").half()
    ref = ops.attention(qkv, heads, kernel="mma").float()
    first, bad, nondet, worst = None, 0, 0, 0.0
    for r in range(rep):
        got = ops.attention(qkv, heads, kernel="tmem").float()
        rel = float((got - ref).norm() / ref.norm())
        worst = max(worst, rel)
        if not (rel < 1.5e-3):
            bad += 1
        if first is None:
            first = got
        elif not torch.equal(first, got):
            nondet += 1
    tot_bad += bad + nondet
    print(f"B={B} L={L} E={E} h={heads} d={E // heads}: {bad}/{rep} wrong, {nondet} differ from run 0, worst rel {worst:.2e}, "
          f"flag {rt.lib().cnb_tc_error_flag()}", flush=True)
print("TOTAL BAD", tot_bad)
