"""Drop-in for the reference's models/blocks.py: the configurable (norm groups, heads, per-block attention flag)
Down / Mid / Up blocks used by the LDM U-Net.  Reference: DownBlock blocks.py:31-150, MidBlock :153-271,
UpBlockUnet :377-503.  Cross-attention branches (:91-104,:136-146,:251-261,:488-501) are dead for every shipped
config and are not built (constructing with cross_attn=True raises NotImplementedError).  `UpBlock` (:274-374) is the
VAE decoder's block (no skip, ConvTranspose over in_channels, optional attention) used by models/vae.py.
"""
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E
from .unet_base import get_time_embedding  # noqa: F401  (blocks.py:5-28 is a duplicate of unet_base.py:5-28)


class _Block(E.ResAttnStack):
    def _io(self, fn, x):
        return ops.nhwc_to_nchw(fn(ops.nchw_to_nhwc(E._check_x(x))))

    @staticmethod
    def _temb(t_emb):
        if t_emb is None:
            return None
        rt.require_cuda(t_emb)
        return E.Temb(raw_temb=t_emb.contiguous())


class DownBlock(_Block):
    def __init__(self, in_channels, out_channels, t_emb_dim, down_sample, num_heads, num_layers, attn,
                 norm_channels, cross_attn=False, context_dim=None):
        super().__init__()
        self.num_layers, self.down_sample, self.attn = num_layers, down_sample, attn
        self.context_dim, self.cross_attn = context_dim, cross_attn
        self._build(in_channels, out_channels, t_emb_dim, num_layers, num_layers if attn else 0, norm_channels,
                    num_heads, cross_attn)
        self.down_sample_conv = nn.Conv2d(out_channels, out_channels, 4, 2, 1) if down_sample else nn.Identity()

    def forward(self, x, t_emb=None, context=None):
        mode = rt.get_mode()
        return self._io(lambda h: E.run_down(self, h, self._temb(t_emb), mode), x)


class MidBlock(_Block):
    def __init__(self, in_channels, out_channels, t_emb_dim, num_heads, num_layers, norm_channels,
                 cross_attn=None, context_dim=None):
        super().__init__()
        self.num_layers, self.context_dim, self.cross_attn = num_layers, context_dim, cross_attn
        self._build(in_channels, out_channels, t_emb_dim, num_layers + 1, num_layers, norm_channels, num_heads,
                    bool(cross_attn))

    def forward(self, x, t_emb=None, context=None):
        mode = rt.get_mode()
        return self._io(lambda h: E.run_mid(self, h, self._temb(t_emb), mode), x)


class UpBlockUnet(_Block):
    attn = True

    def __init__(self, in_channels, out_channels, t_emb_dim, up_sample, num_heads, num_layers, norm_channels,
                 cross_attn=False, context_dim=None):
        super().__init__()
        self.num_layers, self.up_sample = num_layers, up_sample
        self.cross_attn, self.context_dim = cross_attn, context_dim
        self._build(in_channels, out_channels, t_emb_dim, num_layers, num_layers, norm_channels, num_heads, cross_attn)
        half = in_channels // 2
        self.up_sample_conv = nn.ConvTranspose2d(half, half, 4, 2, 1) if up_sample else nn.Identity()

    def forward(self, x, out_down=None, t_emb=None, context=None):
        mode = rt.get_mode()
        skip = None if out_down is None else ops.nchw_to_nhwc(E._check_x(out_down))
        return self._io(lambda h: E.run_up(self, h, skip, self._temb(t_emb), mode), x)


class UpBlock(_Block):
    """blocks.py:284-374: ConvTranspose2d(in, in, 4, 2, 1) or identity, then num_layers x [resnet, (self-attn)]."""

    def __init__(self, in_channels, out_channels, t_emb_dim, up_sample, num_heads, num_layers, attn, norm_channels):
        super().__init__()
        self.num_layers, self.up_sample, self.attn = num_layers, up_sample, attn
        self._build(in_channels, out_channels, t_emb_dim, num_layers, num_layers if attn else 0, norm_channels,
                    num_heads)
        self.up_sample_conv = nn.ConvTranspose2d(in_channels, in_channels, 4, 2, 1) if up_sample else nn.Identity()

    def forward(self, x, out_down=None, t_emb=None):
        mode = rt.get_mode()
        skip = None if out_down is None else ops.nchw_to_nhwc(E._check_x(out_down))
        return self._io(lambda h: E.run_up(self, h, skip, self._temb(t_emb), mode), x)
