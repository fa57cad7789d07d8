"""ControlNet wiring shared by the DDPM (models/controlnet.py:158-225) and LDM (models/controlnet_ldm.py:117-179)
variants, scheduled channels-last on libcnb200 kernels.

Dataflow (names as in the reference):
    a  = trained.conv_in(x)                     ; T_skips = inputs of trained.downs
    c  = control.conv_in(x) + hint_block(hint)  (hint feature cached: it depends on neither t nor x)
    C_skips[i] = zero_conv_i(c) written straight into the second half of the i-th up-block concat buffer,
                 with the trained skip added in the conv epilogue (controlnet.py:190-192, 216-218)
    mids: c = control.mid(c); a = trained.mid(a) + mid_zero_conv(c)   (injection fused in the epilogue, :207)
    ups on the trained U-Net, then norm_out / SiLU / conv_out.
"""
from .. import ops
from .. import runtime as rt
from . import _engine as E


def split_prefix(state, prefix, exclude=()):
    """{k[len(prefix):]: v} for the keys under `prefix` that are not under any of `exclude`."""
    return {k[len(prefix):]: v for k, v in state.items()
            if k.startswith(prefix) and not any(k.startswith(e) for e in exclude)}


def controlnet_forward(trained, control, down_zero, mid_zero, hint_feat_fn, x, t, hint):
    x = E._check_x(x)
    hint = E._check_x(hint)
    mode = rt.get_mode()
    dev = x.device
    xn = ops.nchw_to_nhwc(x)

    plan_t = E.temb_plan(trained, E.unet_time(trained, t, dev))
    plan_c = E.temb_plan(control, E.unet_time(control, t, dev), with_ups=False)

    # frozen encoder
    a = E.conv_in(trained, xn, mode)
    t_skips = []
    for d in trained.downs:
        t_skips.append(a)
        a = E.run_down(d, a, plan_t[d], mode)

    # control branch; skip sums land directly in the concat buffers of the decoder
    c = E.conv_in(control, xn, mode, residual=hint_feat_fn(hint, mode))
    cats = []
    for i, d in enumerate(control.downs):
        B, H, W, C = c.shape
        zc = down_zero[i]
        cat = ops.empty(B, H, W, 2 * C, device=dev, dtype=c.dtype)
        E.conv16(c, zc.weight, "1x1", C, mode, bias=E.raw(zc.bias), residual=t_skips[i], out=cat, out_coff=C)
        cats.append(cat)
        c = E.run_down(d, c, plan_c[d], mode)

    for i in range(len(control.mids)):
        c = E.run_mid(control.mids[i], c, plan_c[control.mids[i]], mode)
        a = E.run_mid(trained.mids[i], a, plan_t[trained.mids[i]], mode)
        zc = mid_zero[i]
        a = E.conv16(c, zc.weight, "1x1", zc.out_channels, mode, bias=E.raw(zc.bias), residual=a)

    for u in trained.ups:
        a = E.run_up(u, a, None, plan_t[u], mode, cat=cats.pop())
    return ops.nhwc_to_nchw(E.conv_out(trained, a, mode))
