"""Drop-in for the reference's models/unet_cond_base.py (LDM U-Net assembled from blocks.py), unconditional path.
Reference: Unet.__init__ unet_cond_base.py:15-123, forward :125-184.  `condition_config` (class / text / image
conditioning) is absent from every shipped config (celebhq.yaml has none) and is rejected here.
"""
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E
from ._entry import host_entry
from .blocks import DownBlock, MidBlock, UpBlockUnet


class Unet(nn.Module):
    def __init__(self, im_channels, model_config, use_up=True):
        super().__init__()
        c = model_config
        self.down_channels, self.mid_channels = c['down_channels'], c['mid_channels']
        self.t_emb_dim, self.down_sample = c['time_emb_dim'], c['down_sample']
        self.num_down_layers, self.num_mid_layers, self.num_up_layers = (
            c['num_down_layers'], c['num_mid_layers'], c['num_up_layers'])
        self.attns, self.norm_channels = c['attn_down'], c['norm_channels']
        self.num_heads, self.conv_out_channels = c['num_heads'], c['conv_out_channels']
        dc, mc = self.down_channels, self.mid_channels
        assert mc[0] == dc[-1]
        assert mc[-1] == dc[-2]
        assert len(self.down_sample) == len(dc) - 1
        assert len(self.attns) == len(dc) - 1
        self.condition_config = c.get('condition_config', None)
        if self.condition_config is not None:
            raise NotImplementedError("class/text/image conditioning is outside the ControlNet hot path "
                                      "(no shipped config sets condition_config)")
        self.class_cond = self.text_cond = self.image_cond = self.cond = False
        self.text_embed_dim = None

        D, G, NH = self.t_emb_dim, self.norm_channels, self.num_heads
        self.conv_in = nn.Conv2d(im_channels, dc[0], kernel_size=3, padding=1)
        self.t_proj = nn.Sequential(nn.Linear(D, D), nn.SiLU(), nn.Linear(D, D))
        self.up_sample = list(reversed(self.down_sample))
        n = len(dc) - 1
        self.downs = nn.ModuleList(
            [DownBlock(dc[i], dc[i + 1], D, down_sample=self.down_sample[i], num_heads=NH,
                       num_layers=self.num_down_layers, attn=self.attns[i], norm_channels=G) for i in range(n)])
        self.mids = nn.ModuleList(
            [MidBlock(mc[i], mc[i + 1], D, num_heads=NH, num_layers=self.num_mid_layers, norm_channels=G)
             for i in range(len(mc) - 1)])
        self.ups = nn.ModuleList([])
        if use_up:
            for i in reversed(range(n)):
                self.ups.append(UpBlockUnet(dc[i] * 2, dc[i - 1] if i != 0 else self.conv_out_channels, D,
                                            up_sample=self.down_sample[i], num_heads=NH,
                                            num_layers=self.num_up_layers, norm_channels=G))
            self.norm_out = nn.GroupNorm(G, self.conv_out_channels)
            self.conv_out = nn.Conv2d(self.conv_out_channels, im_channels, kernel_size=3, padding=1)

    @host_entry
    def forward(self, x, t, cond_input=None):
        x = E._check_x(x)
        mode = rt.get_mode()
        temb = E.unet_time(self, t, x.device)
        plan = E.temb_plan(self, temb)
        h = E.conv_in(self, ops.nchw_to_nhwc(x), mode)
        return ops.nhwc_to_nchw(E.run_unet_body(self, h, plan, mode))
