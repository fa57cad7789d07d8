"""Development probe for the tcgen05 GEMM pipeline (run on the GPU box; not collected by pytest).
Runs cnb_tc_gemm_selftest on a few shapes, prints error statistics and, on mismatch, dumps operands/results to
gpurun_out/ for offline layout forensics."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rt = importlib.import_module("controlnet-pytorch_b200.runtime")
L = rt.lib()
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
print("has_tcgen05", L.cnb_has_tcgen05(), torch.cuda.get_device_name(0), flush=True)

ok_all = True
for (M, N, K) in [(128, 16, 32), (128, 64, 32), (128, 128, 32), (128, 64, 64), (256, 64, 128), (300, 48, 36),
                  (1000, 128, 288), (4096, 256, 2304), (50000, 64, 576)]:
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    ref = (a.double() @ w.double().t()).float()
    ad, wd = a.cuda(), w.cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = L.cnb_tc_gemm_selftest(ad.data_ptr(), wd.data_ptr(), out.data_ptr(), M, N, K, rt.MODE_F16, 0)
    flag = L.cnb_tc_error_flag()
    o = out.cpu()
    nan = int(torch.isnan(o).sum())
    err = float((o - ref).norm() / ref.norm()) if nan == 0 else float("nan")
    good = rc == 0 and flag == 0 and nan == 0 and err < 2e-3
    ok_all &= good
    print(f"M={M} N={N} K={K} rc={rc} flag={flag} nan={nan} rel_l2={err:.3e} {'OK' if good else 'FAIL'}", flush=True)
    if not good and M * N <= 300 * 64:
        np.savez(os.path.join(out_dir, f"tc_probe_{M}_{N}_{K}.npz"), a=a.numpy(), w=w.numpy(), out=o.numpy(), ref=ref.numpy())
    if rc != 0:
        print("  error:", L.cnb_last_error().decode())
print("TC_PROBE", "PASS" if ok_all else "FAIL", flush=True)
