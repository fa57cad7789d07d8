"""Single-U-Net student body shared by the Consistency and Distribution-Matching ControlNets:
conv_in(x) + hint_block(hint) -> downs -> mids -> ups -> norm_out/SiLU/conv_out with the students' own
t_proj = SiLU -> Linear (consistency_controlnet_distilled.py:35-38,103-125;
distribution_matching_controlnet.py:114-157).  unet.t_proj stays in the state_dict but is unused, as in the reference.
"""
import torch
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E


def build_hint_block(hint_channels, c0, zero_tail):
    # layers are created in the reference's order so that seeded default initialisation draws the same numbers
    seq = nn.Sequential(nn.Conv2d(hint_channels, 64, kernel_size=3, padding=1), nn.SiLU(),
                        nn.Conv2d(64, 128, kernel_size=3, padding=1), nn.SiLU(),
                        nn.Conv2d(128, c0, kernel_size=3, padding=1), nn.SiLU(),
                        nn.Conv2d(c0, c0, kernel_size=1, padding=0))
    if zero_tail:
        E.zero_(seq[6])
    return seq


def student_body(model, x_nhwc, t_index, hint, mode):
    """x_nhwc: channels-last input already scaled; t_index: int64 CUDA (B,) or (1,).  Returns channels-last output."""
    unet = model.unet
    t_index = E.check_t(t_index, x_nhwc.shape[0], x_nhwc.device)
    emb = E.sinusoid(t_index, model.t_emb_dim, x_nhwc.device)
    lin = model.t_proj[1]
    temb = ops.linear_small(emb, E.raw(lin.weight), E.raw(lin.bias), silu_in=True)
    plan = E.temb_plan(unet, temb)
    seq = model.hint_block
    hf = model._hint_cache.get(hint, list(seq.parameters()), mode,
                               lambda: E.hint_stack_ddpm(seq, ops.nchw_to_nhwc(hint), mode))
    h = E.conv_in(unet, x_nhwc, mode, residual=hf)
    return E.run_unet_body(unet, h, plan, mode)

def teacher_x0(teacher, scheduler, x_t, t, hint):
    """x0 prediction of a DDPM ControlNet teacher: eps = teacher(x_t, t, hint), then
    clamp((x_t - sqrt(1 - abar_t) eps) / sqrt(abar_t), -1, 1) with per-sample t
    (distribution_matching_controlnet.py:191-216; consistency_controlnet_distilled.py:201-228)."""
    with torch.no_grad():
        x_t = E._check_x(x_t)
        t = E.as_t(t, x_t.device)
        eps = teacher(x_t, t, hint)
        tabs = scheduler.__dict__.setdefault("_cnb_dev_tables", {})
        key = str(x_t.device)
        if key not in tabs:
            tabs[key] = (scheduler.sqrt_one_minus_alpha_cum_prod.to(x_t.device).contiguous(),
                         scheduler.sqrt_alpha_cum_prod.to(x_t.device).contiguous())
        s1, s2 = tabs[key]
        if t.numel() not in (1, x_t.shape[0]):
            raise RuntimeError("shape mismatch: t has %d entries for a batch of %d" % (t.numel(), x_t.shape[0]))
        return ops.x0_from_eps(x_t, eps.contiguous(), s1, s2, t)
