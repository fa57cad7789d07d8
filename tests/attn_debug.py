"""Development aid (GPU box; not collected by pytest): where does the tcgen05 / TMEM attention kernel differ from the
mma.sync kernel?  Prints the error per (sample, head, 128-row query tile) and whether two runs agree bit for bit.
    python tests/attn_debug.py B L E heads [repeat]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("controlnet-pytorch_b200.ops")
rt = importlib.import_module("controlnet-pytorch_b200.runtime")
rt.lib()
B, L, E, heads = (int(a) for a in sys.argv[1:5])
rep = int(sys.argv[5]) if len(sys.argv) > 5 else 2
torch.manual_seed(5)
qkv = torch.randn(B, L, 1, 3 * E, device="cuda").half()
ref = ops.attention(qkv, heads, kernel="mma").float()
d = E // heads
outs = []
for r in range(rep):
    got = ops.attention(qkv, heads, kernel="tmem").float()
    torch.cuda.synchronize()
    outs.append(got)
    err = (got - ref).reshape(B, L, heads, d)
    nq = (L + 127) // 128
    bad = []
    for q in range(nq):
        e = err[:, q * 128:(q + 1) * 128].abs().amax(dim=(1, 3))        # (B, heads)
        idx = (e > 5e-3).nonzero()
        for b, h in idx.tolist():
            bad.append((b, h, q, float(e[b, h])))
    rel = float((got - ref).norm() / ref.norm())
    print(f"run {r}: rel-L2 vs mma {rel:.3e}; {len(bad)} bad (sample, head, q-tile) of {B * heads * nq}; flag {rt.lib().cnb_tc_error_flag()}")
    for b, h, q, e in bad[:24]:
        rows = err[b, q * 128:(q + 1) * 128, h].abs().amax(dim=1)
        badrows = (rows > 5e-3).nonzero().flatten().tolist()
        print(f"   b={b} h={h} qtile={q} max|err|={e:.3e} bad rows: {len(badrows)} first {badrows[:6]} last {badrows[-3:]}")
if rep > 1:
    print("runs identical:", all(torch.equal(outs[0], o) for o in outs[1:]))
