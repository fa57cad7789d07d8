"""Diagnostic (GPU box; not collected by pytest, not part of the product or of bench.py): times the oracle's ATen
restatement of one MNIST ControlNet denoising timestep on the GPU through stock PyTorch (cuDNN / cuBLAS), i.e. what
the unmodified reference modules would reach on the same B200, so DESIGN.md can state where libcnb200 stands against
the library path.   python tests/torch_cuda_ref.py [batch] [steps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cn_oracle as O  # noqa: E402

_te = O.time_embedding
O.time_embedding = lambda t, dim: _te(torch.as_tensor(t).cpu(), dim).cuda()   # the oracle builds its table on the CPU

syn = importlib.import_module("controlnet-pytorch_b200.utils.synthetic")
cn = importlib.import_module("controlnet-pytorch_b200.models.controlnet")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = syn.MNIST_PARAMS
sd = {k: v.cuda() for k, v in syn.det_state_dict(cn.ControlNet(cfg).state_dict()).items()}
x = torch.randn(B, 1, 28, 28, device="cuda")
hint = (torch.rand(B, 1, 28, 28, device="cuda") < 0.1).float().repeat(1, 3, 1, 1)
t = torch.tensor([500], device="cuda")
for name, tf32, amp in (("fp32", False, None), ("tf32", True, None), ("bf16-autocast", True, torch.bfloat16),
                        ("fp16-autocast", True, torch.float16)):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=amp, enabled=amp is not None):
            for _ in range(2):
                O.controlnet_ddpm_forward(sd, cfg, x, t, hint)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                O.controlnet_ddpm_forward(sd, cfg, x, t, hint)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"torch-cuda[{name}] B={B}: {ms:.2f} ms/timestep  {B / ms:.1f} sample-steps/ms  "
              f"{B * 3.878e9 / ms / 1e9:.1f} TFLOP/s", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"torch-cuda[{name}] failed: {type(e).__name__}: {e}", flush=True)
