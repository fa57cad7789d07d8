"""Host-side staging at the PUBLIC forward of the drop-in models.

The reference's modules run wherever their tensors live; its only acceptance test (`test_distribution_matching.py:36-56`)
builds the model and its inputs on the CPU.  The B200 path has no CPU arithmetic, so a public `forward` that receives
host tensors (or whose module still holds host parameters) STAGES them: a device replica of the module is kept (rebuilt
when a host parameter changes), the inputs are copied host -> device, the forward runs on libcnb200, and the result is
copied back to the device the caller's `x` lives on.  This is a copy at the boundary, not a fallback: without a CUDA
device it raises, and nothing is ever computed on the host.
"""
import copy
import functools

import torch

from .. import runtime as rt


def _exec_device():
    if not torch.cuda.is_available():
        raise rt.CnbError("controlnet-pytorch_b200 has no CPU path: a CUDA device is required")
    return torch.device("cuda", torch.cuda.current_device())


def _replica(module, dev):
    """Device copy of a host-resident module, cached on the module and rebuilt when any host parameter / buffer changes."""
    stamp = tuple((p._version, p.data_ptr()) for p in list(module.parameters()) + list(module.buffers())) + (str(dev),)
    ent = module.__dict__.get("_cnb_replica")
    if ent is None or ent[0] != stamp:
        memo = {}
        if ent is not None:            # never deep-copy an old replica along with the module
            module.__dict__.pop("_cnb_replica", None)
        rep = copy.deepcopy(module, memo).to(dev)
        rep.train(module.training)
        ent = (stamp, rep)
        module.__dict__["_cnb_replica"] = ent
    return ent[1]


def host_entry(fn):
    """Decorator for `forward(self, x, *rest)`: device tensors with a device module go straight through."""
    @functools.wraps(fn)
    def wrapped(self, x, *rest, **kw):
        p = next(self.parameters(), None)
        params_on_host = p is not None and not p.is_cuda
        x_on_host = isinstance(x, torch.Tensor) and not x.is_cuda
        if not params_on_host and not x_on_host and all(
                (not isinstance(a, torch.Tensor)) or a.is_cuda or a.dim() == 0 or a.numel() <= 1 or
                not torch.is_floating_point(a) for a in rest):
            return fn(self, x, *rest, **kw)
        dev = p.device if (p is not None and p.is_cuda) else _exec_device()
        target = _replica(self, dev) if params_on_host else self
        home = x.device if isinstance(x, torch.Tensor) else torch.device("cpu")

        def up(a):
            return a.to(dev) if isinstance(a, torch.Tensor) else a
        with torch.no_grad():
            out = fn(target, up(x), *[up(a) for a in rest], **{k: up(v) for k, v in kw.items()})
        return out.to(home) if isinstance(out, torch.Tensor) else out
    return wrapped
