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


_SIDE = {}
_FORCE = [None]          # sampler override: True / False / None (= environment default)


def set_branch_parallel(value):
    """Override CNB_BRANCH_PARALLEL for the calls that follow (None restores the default); used by the sampler when it
    already runs batch halves on parallel streams."""
    _FORCE[0] = value


def _side_stream(dev):
    import torch
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def controlnet_forward(trained, control, down_zero, mid_zero, hint_feat_fn, x, t, hint):
    import os
    import torch
    x = E._check_x(x)
    hint = E._check_x(hint)
    mode = rt.get_mode()
    dev = x.device
    xn = ops.nchw_to_nhwc(x)
    t = E.check_t(t, x.shape[0], dev)

    plan_t = E.temb_plan(trained, E.unet_time(trained, t, dev))
    plan_c = E.temb_plan(control, E.unet_time(control, t, dev), with_ups=False)
    hf = hint_feat_fn(hint, mode)

    def control_encoder():
        """control.conv_in(x) + hint feature -> downs -> mids; returns the inputs of the zero convs."""
        c = E.conv_in(control, xn, mode, residual=hf)
        cs, cms = [], []
        for d in control.downs:
            cs.append(c)
            c = E.run_down(d, c, plan_c[d], mode)
        for mblk in control.mids:
            c = E.run_mid(mblk, c, plan_c[mblk], mode)
            cms.append(c)
        return cs, cms

    # The frozen encoder and the control encoder only meet at the zero convs, so they run on two streams (forked and
    # joined with events; under CUDA-graph capture these become parallel graph branches): the kernels of one branch fill
    # the pipes the other leaves idle -- MUFU-bound attention next to tensor- / HBM-bound convolutions and GroupNorms.
    # Measured (MNIST, graph replay): B = 64 2.25 -> 1.95 ms, B = 256 4.15 -> 3.90 ms, B = 1024 13.00 -> 12.61 ms per step;
    # results are bit-identical (same kernels, same order per tensor).  CNB_BRANCH_PARALLEL=0 turns it off.
    # Only under graph capture: in eager mode the extra stream bookkeeping (event waits, record_stream on every
    # hand-over tensor) is host time, which is what a small-batch eager loop is bound by (CelebHQ B = 16: 6.9 -> 19.7 ms).
    parallel = os.environ.get("CNB_BRANCH_PARALLEL", "1") == "1" if _FORCE[0] is None else bool(_FORCE[0])
    parallel = parallel and torch.cuda.is_current_stream_capturing()
    if parallel:
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            cs, cms = control_encoder()
    # frozen encoder
    a = E.conv_in(trained, xn, mode)
    t_skips = []
    for d in trained.downs:
        t_skips.append(a)
        a = E.run_down(d, a, plan_t[d], mode)
    if parallel:
        cur.wait_stream(side)
        for tns in cs + cms:
            tns.record_stream(cur)
    else:
        cs, cms = control_encoder()

    # skip sums land directly in the concat buffers of the decoder (controlnet.py:190-192, 216-218)
    cats = []
    for i, c in enumerate(cs):
        B, H, W, C = c.shape
        zc = down_zero[i]
        cat = ops.empty(B, H, W, 2 * C, device=dev, dtype=c.dtype)
        E.conv16(c, zc.weight, "1x1", C, mode, bias=E.raw(zc.bias), residual=t_skips[i], out=cat, out_coff=C)
        cats.append(cat)

    # mids: a = trained.mid(a) + mid_zero_conv(control.mid output)   (injection fused in the epilogue, :207)
    for i in range(len(control.mids)):
        a = E.run_mid(trained.mids[i], a, plan_t[trained.mids[i]], mode)
        zc = mid_zero[i]
        a = E.conv16(cms[i], zc.weight, "1x1", zc.out_channels, mode, bias=E.raw(zc.bias), residual=a)

    for u in trained.ups:
        a = E.run_up(u, a, None, plan_t[u], mode, cat=cats.pop())
    return ops.nhwc_to_nchw(E.conv_out(trained, a, mode))
