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

    @staticmethod
    def _tail(norm, conv, h, mode):
        """GroupNorm -> SiLU -> 3x3 conv to a handful of channels; returns (nhwc tensor, valid channels).
        Wide inputs (decoder_conv_out: 128 -> 3 at 128x128 is 3.5 k MACs per pixel) run on the tensor core with Cout
        zero-padded to 16 -- 5x wasted MMA work is still 10x faster than the direct FFMA kernel, which is meant for
        the 16-channel U-Net tails; narrow ones keep the exact direct kernel."""
        cout, cin = conv.out_channels, conv.in_channels
        tc = mode != rt.MODE_F32 and E._F16_ENABLED and cout <= 4 and cin % 16 == 0 and cin >= 64
        h16 = mode != rt.MODE_F32 and E._F16_ENABLED and cout <= 4 and cin % 4 == 0
        h = ops.groupnorm(h, E.raw(norm.weight), E.raw(norm.bias), norm.num_groups, silu=True, out_f16=h16)
        if tc:
            def build():
                w = torch.zeros((16,) + tuple(conv.weight.shape[1:]), device=conv.weight.device, dtype=torch.float32)
                w[:cout] = conv.weight.detach()
                b = torch.zeros(16, device=conv.weight.device, dtype=torch.float32)
                b[:cout] = conv.bias.detach()
                packed = ops.pack_conv_weight(w, False)
                return packed, ops.cast_f16(packed), b
            stamp_holder = conv.weight
            store = stamp_holder.__dict__.setdefault("_cnb_pack", {})
            stamp = (conv.weight._version, conv.weight.data_ptr(), conv.bias._version, conv.bias.data_ptr())
            ent = store.get(("pad16",))
            if ent is None or ent[0] != stamp:
                with torch.no_grad():
                    ent = (stamp, build())
                store[("pad16",)] = ent
            packed, packed16, bias = ent[1]
            return ops.conv(h, packed, "3x3", 16, bias=bias, mode=mode, weight_lp=packed16, out_f16=True), cout
        return ops.conv(h, E.packed_conv(conv.weight, mode), "3x3", cout, bias=E.raw(conv.bias), mode=mode,
                        out_f16=False), cout

    def _decode_nhwc(self, z_nhwc, mode):
        h = self._conv(self.post_quant_conv, z_nhwc, "1x1", mode)
        h = self._conv(self.decoder_conv_in, h, "3x3", mode)
        for mid in self.decoder_mids:
            h = E.run_mid(mid, h, None, mode)
        for up in self.decoder_layers:
            h = E.run_up(up, h, None, None, mode)
        return self._tail(self.decoder_norm_out, self.decoder_conv_out, h, mode)

    def decode(self, z):
        mode = rt.get_mode()
        h, c = self._decode_nhwc(ops.nchw_to_nhwc(E._check_x(z)), mode)
        return ops.nhwc_to_nchw(h, c=c)

    def _encode_out(self, x, mode):
        h = self._conv(self.encoder_conv_in, ops.nchw_to_nhwc(E._check_x(x)), "3x3", mode)
        for down in self.encoder_layers:
            h = E.run_down(down, h, None, mode)
        for mid in self.encoder_mids:
            h = E.run_mid(mid, h, None, mode)
        h, _ = self._tail(self.encoder_norm_out, self.encoder_conv_out, h, mode)
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
