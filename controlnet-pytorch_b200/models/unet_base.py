"""Drop-in for the reference's models/unet_base.py (DDPM U-Net used by the MNIST/CIFAR ControlNet and both
students): same constructors, forward signatures and state_dict keys; forward runs on libcnb200 kernels.

Reference: get_time_embedding unet_base.py:5-28, DownBlock :31-112, MidBlock :115-199, UpBlock :202-289,
Unet :292-374.  GroupNorm(8, .) and 4 heads are hard-coded there (:47, :40) and the last UpBlock emits 16 channels.
"""
import torch
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E
from ._entry import host_entry

_GROUPS = 8


def get_time_embedding(time_steps, temb_dim):
    """[sin | cos](t / 10000^(i/(D/2))) for a CUDA tensor of timesteps -> (B, temb_dim) fp32 (unet_base.py:5-28)."""
    assert temb_dim % 2 == 0, "time embedding dimension must be divisible by 2"
    rt.require_cuda(time_steps)
    return E.sinusoid(time_steps, temb_dim, time_steps.device)


class _Block(E.ResAttnStack):
    attn = True

    def _io(self, fn, x, *extra):
        x = ops.nchw_to_nhwc(E._check_x(x))
        return ops.nhwc_to_nchw(fn(x, *extra))


class DownBlock(_Block):
    def __init__(self, in_channels, out_channels, t_emb_dim, down_sample=True, num_heads=4, num_layers=1):
        super().__init__()
        self.num_layers, self.down_sample = num_layers, down_sample
        self._build(in_channels, out_channels, t_emb_dim, num_layers, num_layers, _GROUPS, num_heads)
        self.down_sample_conv = nn.Conv2d(out_channels, out_channels, 4, 2, 1) if down_sample else nn.Identity()

    def forward(self, x, t_emb):
        rt.require_cuda(x, t_emb)
        mode = rt.get_mode()
        return self._io(lambda h: E.run_down(self, h, E.Temb(raw_temb=t_emb.contiguous()), mode), x)


class MidBlock(_Block):
    def __init__(self, in_channels, out_channels, t_emb_dim, num_heads=4, num_layers=1):
        super().__init__()
        self.num_layers = num_layers
        self._build(in_channels, out_channels, t_emb_dim, num_layers + 1, num_layers, _GROUPS, num_heads)

    def forward(self, x, t_emb):
        rt.require_cuda(x, t_emb)
        mode = rt.get_mode()
        return self._io(lambda h: E.run_mid(self, h, E.Temb(raw_temb=t_emb.contiguous()), mode), x)


class UpBlock(_Block):
    attn = True

    def __init__(self, in_channels, out_channels, t_emb_dim, up_sample=True, num_heads=4, num_layers=1):
        super().__init__()
        self.num_layers, self.up_sample = num_layers, up_sample
        self._build(in_channels, out_channels, t_emb_dim, num_layers, num_layers, _GROUPS, num_heads)
        half = in_channels // 2
        self.up_sample_conv = nn.ConvTranspose2d(half, half, 4, 2, 1) if up_sample else nn.Identity()

    def forward(self, x, out_down, t_emb):
        rt.require_cuda(x, out_down, t_emb)
        mode = rt.get_mode()
        skip = ops.nchw_to_nhwc(E._check_x(out_down))
        return self._io(lambda h: E.run_up(self, h, skip, E.Temb(raw_temb=t_emb.contiguous()), mode), x)


class Unet(nn.Module):
    """model_config keys: im_channels, down_channels, mid_channels, time_emb_dim, down_sample, num_down_layers,
    num_mid_layers, num_up_layers (unet_base.py:299-306).  use_up=False drops ups / norm_out / conv_out entirely."""

    def __init__(self, model_config, use_up=True):
        super().__init__()
        c = model_config
        self.down_channels, self.mid_channels = c['down_channels'], c['mid_channels']
        self.t_emb_dim, self.down_sample = c['time_emb_dim'], c['down_sample']
        self.num_down_layers, self.num_mid_layers, self.num_up_layers = (
            c['num_down_layers'], c['num_mid_layers'], c['num_up_layers'])
        dc, mc = self.down_channels, self.mid_channels
        assert mc[0] == dc[-1]
        assert mc[-1] == dc[-2]
        assert len(self.down_sample) == len(dc) - 1

        D = self.t_emb_dim
        self.t_proj = nn.Sequential(nn.Linear(D, D), nn.SiLU(), nn.Linear(D, D))
        self.up_sample = list(reversed(self.down_sample))
        self.conv_in = nn.Conv2d(c['im_channels'], dc[0], kernel_size=3, padding=(1, 1))
        n = len(dc) - 1
        self.downs = nn.ModuleList(
            [DownBlock(dc[i], dc[i + 1], D, down_sample=self.down_sample[i], num_layers=self.num_down_layers)
             for i in range(n)])
        self.mids = nn.ModuleList(
            [MidBlock(mc[i], mc[i + 1], D, num_layers=self.num_mid_layers) for i in range(len(mc) - 1)])
        if use_up:
            self.ups = nn.ModuleList(
                [UpBlock(dc[i] * 2, dc[i - 1] if i != 0 else 16, D, up_sample=self.down_sample[i],
                         num_layers=self.num_up_layers) for i in reversed(range(n))])
            self.norm_out = nn.GroupNorm(_GROUPS, 16)
            self.conv_out = nn.Conv2d(16, c['im_channels'], kernel_size=3, padding=1)

    @host_entry
    def forward(self, x, t):
        """x (B, C, H, W) fp32 CUDA, t int (1,) or (B,) -> eps with x's shape (unet_base.py:341-374)."""
        x = E._check_x(x)
        mode = rt.get_mode()
        temb = E.unet_time(self, E.check_t(t, x.shape[0], x.device), x.device)
        plan = E.temb_plan(self, temb)
        h = E.conv_in(self, ops.nchw_to_nhwc(x), mode)
        return ops.nhwc_to_nchw(E.run_unet_body(self, h, plan, mode))
