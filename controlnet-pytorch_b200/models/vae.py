"""Drop-in for the reference's models/vae.py: the autoencoder whose `decode` turns the sampled latents into images
right after the LDM sampling loop (tools/sample_ldm_controlnet.py:54-56).  Same constructor
(`VAE(im_channels, model_config)`, vae.py:6-84), attribute names and state_dict keys; `encode` (:86-100), `decode`
(:102-114) and `forward` (:116-119) run on the libcnb200 kernels (channels-last, fp16 activation stream in the
tensor-core modes), reusing the Down / Mid / Up block schedules of models/_engine.py with the time embedding off
(`t_emb_dim=None`).
"""
import torch
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E
from .blocks import DownBlock, MidBlock, UpBlock


class VAE(nn.Module):
    def __init__(self, im_channels, model_config):
        super().__init__()
        c = model_config
        self.down_channels, self.mid_channels = c['down_channels'], c['mid_channels']
        self.down_sample = c['down_sample']
        self.num_down_layers, self.num_mid_layers, self.num_up_layers = (c['num_down_layers'], c['num_mid_layers'],
                                                                         c['num_up_layers'])
        self.attns = c['attn_down']
        self.z_channels, self.norm_channels, self.num_heads = c['z_channels'], c['norm_channels'], c['num_heads']
        assert self.mid_channels[0] == self.down_channels[-1]
        assert self.mid_channels[-1] == self.down_channels[-1]
        assert len(self.down_sample) == len(self.down_channels) - 1
        assert len(self.attns) == len(self.down_channels) - 1
        self.up_sample = list(reversed(self.down_sample))
        dc, mc = self.down_channels, self.mid_channels

        # ---- encoder (vae.py:35-60)
        self.encoder_conv_in = nn.Conv2d(im_channels, dc[0], kernel_size=3, padding=(1, 1))
        self.encoder_layers = nn.ModuleList([
            DownBlock(dc[i], dc[i + 1], t_emb_dim=None, down_sample=self.down_sample[i], num_heads=self.num_heads,
                      num_layers=self.num_down_layers, attn=self.attns[i], norm_channels=self.norm_channels)
            for i in range(len(dc) - 1)])
        self.encoder_mids = nn.ModuleList([
            MidBlock(mc[i], mc[i + 1], t_emb_dim=None, num_heads=self.num_heads, num_layers=self.num_mid_layers,
                     norm_channels=self.norm_channels) for i in range(len(mc) - 1)])
        self.encoder_norm_out = nn.GroupNorm(self.norm_channels, dc[-1])
        self.encoder_conv_out = nn.Conv2d(dc[-1], 2 * self.z_channels, kernel_size=3, padding=1)
        self.pre_quant_conv = nn.Conv2d(2 * self.z_channels, 2 * self.z_channels, kernel_size=1)

        # ---- decoder (vae.py:63-84)
        self.post_quant_conv = nn.Conv2d(self.z_channels, self.z_channels, kernel_size=1)
        self.decoder_conv_in = nn.Conv2d(self.z_channels, mc[-1], kernel_size=3, padding=(1, 1))
        self.decoder_mids = nn.ModuleList([
            MidBlock(mc[i], mc[i - 1], t_emb_dim=None, num_heads=self.num_heads, num_layers=self.num_mid_layers,
                     norm_channels=self.norm_channels) for i in reversed(range(1, len(mc)))])
        self.decoder_layers = nn.ModuleList([
            UpBlock(dc[i], dc[i - 1], t_emb_dim=None, up_sample=self.down_sample[i - 1], num_heads=self.num_heads,
                    num_layers=self.num_up_layers, attn=self.attns[i - 1], norm_channels=self.norm_channels)
            for i in reversed(range(1, len(dc)))])
        self.decoder_norm_out = nn.GroupNorm(self.norm_channels, dc[0])
        self.decoder_conv_out = nn.Conv2d(dc[0], im_channels, kernel_size=3, padding=1)

    # ---- kernels ---------------------------------------------------------------------------------------
    @staticmethod
    def _conv(conv, h, kind, mode):
        return E.conv16(h, conv.weight, kind, conv.out_channels, mode, bias=E.raw(conv.bias))

    def _decode_nhwc(self, z_nhwc, mode):
        h = self._conv(self.post_quant_conv, z_nhwc, "1x1", mode)
        h = E.conv3x3_narrow_in(self.decoder_conv_in, h, mode)
        for mid in self.decoder_mids:
            h = E.run_mid(mid, h, None, mode)
        for up in self.decoder_layers:
            h = E.run_up(up, h, None, None, mode)
        return E.gn_silu_conv_tail(self.decoder_norm_out, self.decoder_conv_out, h, mode)

    def decode(self, z):
        mode = rt.get_mode()
        return ops.nhwc_to_nchw(self._decode_nhwc(ops.nchw_to_nhwc(E._check_x(z)), mode))

    def _encode_out(self, x, mode):
        h = E.conv3x3_narrow_in(self.encoder_conv_in, ops.nchw_to_nhwc(E._check_x(x)), mode)
        for down in self.encoder_layers:
            h = E.run_down(down, h, None, mode)
        for mid in self.encoder_mids:
            h = E.run_mid(mid, h, None, mode)
        h = E.gn_silu_conv_tail(self.encoder_norm_out, self.encoder_conv_out, h, mode)
        h = ops.conv(h, E.packed_conv(self.pre_quant_conv.weight, mode), "1x1", self.pre_quant_conv.out_channels,
                     bias=E.raw(self.pre_quant_conv.bias), mode=mode, out_f16=False)
        return ops.nhwc_to_nchw(h)

    def encode(self, x):
        out = self._encode_out(x, rt.get_mode())
        mean, logvar = torch.chunk(out, 2, dim=1)
        std = torch.exp(0.5 * logvar)
        # the reference draws the noise on the CPU default generator and moves it (vae.py:99)
        sample = mean + std * torch.randn(mean.shape).to(device=x.device)
        return sample, out

    def forward(self, x):
        z, encoder_output = self.encode(x)
        return self.decode(z), encoder_output
