"""Drop-in for the inference surface of the reference's models/distribution_matching_controlnet.py:
DistributionMatchingControlNet.forward (:120-159) and the Distilled wrapper's constructor / forward (:166-189) with
the reference's attribute names (`student`, `teacher`, `feature_extractor`, `teacher_scheduler`).  The moment /
Wasserstein / Gram losses (:218-358) and the BatchNorm FeatureExtractor forward are training-only (SURVEY.md section
2, row 7): FeatureExtractor is kept as a parameter container so state_dicts round-trip, and refuses to run.
"""
import torch
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E
from ._entry import host_entry
from ._student_common import build_hint_block, student_body
from .controlnet import ControlNet
from .unet_base import Unet, get_time_embedding  # noqa: F401


def make_zero_module(module):
    return E.zero_(module)


class FeatureExtractor(nn.Module):
    """Parameter container with the reference layout (`features.{0..3}.{0,1,3,4}.*`, :16-64)."""

    def __init__(self, in_channels=1, trainable=False):
        super().__init__()
        base = 32 if in_channels == 1 else 64

        def pair(cin, cout, stride):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1, stride=stride), nn.BatchNorm2d(cout),
                                 nn.ReLU(), nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout),
                                 nn.ReLU())
        widths = [base, base * 2, base * 4, base * 8]
        self.features = nn.ModuleList(
            [pair(in_channels if i == 0 else widths[i - 1], w, 1 if i == 0 else 2) for i, w in enumerate(widths)])
        if not trainable:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                    if m.bias is not None:
                        nn.init.constant_(m.bias, 0)
            for prm in self.parameters():
                prm.requires_grad = False

    def forward(self, x):
        raise NotImplementedError("FeatureExtractor is used only by the distillation losses (training); it is "
                                  "outside the B200 inference hot path")


class DistributionMatchingControlNet(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        self.unet = Unet(model_config)
        # the DM hint tail IS zero-initialised (:108-110), unlike the consistency student's
        self.hint_block = build_hint_block(model_config['hint_channels'], self.unet.down_channels[0], zero_tail=True)
        self.t_emb_dim = model_config['time_emb_dim']
        self.t_proj = E.act_linear(self.t_emb_dim, self.t_emb_dim)
        self._hint_cache = E.HintCache()

    @host_entry
    def forward(self, x_t, t, hint):
        """x0 = student(x_t, t, hint): the output IS the sample (:155-157).  t: int scalar / (1,) / (B,)."""
        x_t = E._check_x(x_t)
        hint = E._check_x(hint)
        mode = rt.get_mode()
        out = student_body(self, ops.nchw_to_nhwc(x_t), E.as_t(t, x_t.device), hint, mode)
        return ops.nhwc_to_nchw(out)


class DistributionMatchingControlNetDistilled(nn.Module):
    def __init__(self, model_config, teacher_ckpt_path, device=None):
        super().__init__()
        self.student = DistributionMatchingControlNet(model_config)
        self.teacher = ControlNet(model_config, model_locked=True, model_ckpt=teacher_ckpt_path, device=device)
        self.teacher.eval()
        self.feature_extractor = FeatureExtractor(in_channels=model_config.get('im_channels', 1))
        from ..scheduler.linear_noise_scheduler import LinearNoiseScheduler
        self.teacher_scheduler = LinearNoiseScheduler(num_timesteps=1000, beta_start=0.0001, beta_end=0.02)

    def forward(self, x_t, t, hint):
        return self.student(x_t, t, hint)

    def get_teacher_prediction(self, x_t, t, hint):
        """Teacher's noise prediction converted to x_0 (:191-216), forward-only on the libcnb200 kernels."""
        from ._student_common import teacher_x0
        return teacher_x0(self.teacher, self.teacher_scheduler, x_t, t, hint)
