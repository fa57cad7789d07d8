"""Drop-in for the inference surface of the reference's models/consistency_controlnet_distilled.py:
ConsistencyControlNet (forward + c_skip/c_out/c_in/c_noise, :45-134) and the ConsistencyControlNetDistilled
wrapper's forward / single- and multi-step generate (:198-199, :375-409).  The training losses, EMA update and
DDPM-teacher prediction (:171-177, :201-373) are training-only and outside the hot path.
"""
from copy import deepcopy

import torch
import torch.nn as nn

from .. import ops
from .. import runtime as rt
from . import _engine as E
from ._entry import host_entry
from ._student_common import build_hint_block, student_body
from .controlnet import ControlNet
from .unet_base import Unet


class ConsistencyControlNet(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        self.unet = Unet(model_config)
        # the last 1x1 of the consistency hint block is NOT zero-initialised (:28-30)
        self.hint_block = build_hint_block(model_config['hint_channels'], self.unet.down_channels[0], zero_tail=False)
        self.t_emb_dim = model_config['time_emb_dim']
        self.t_proj = E.act_linear(self.t_emb_dim, self.t_emb_dim)
        self.sigma_min = model_config.get('sigma_min', 0.002)
        self.sigma_max = model_config.get('sigma_max', 80.0)
        self.sigma_data = model_config.get('sigma_data', 0.5)
        self._hint_cache = E.HintCache()

    # ---- EDM scalings: tiny host-visible helpers kept for API parity (:45-74) ---------------------------
    def _sd(self, sigma):
        if not isinstance(sigma, torch.Tensor):
            sigma = torch.tensor(sigma, dtype=torch.float32)
        return sigma, torch.tensor(self.sigma_data, dtype=sigma.dtype, device=sigma.device)

    def c_skip(self, sigma):
        sigma, sd = self._sd(sigma)
        return sd ** 2 / (sigma ** 2 + sd ** 2)

    def c_out(self, sigma):
        sigma, sd = self._sd(sigma)
        return sigma * sd / torch.sqrt(sigma ** 2 + sd ** 2)

    def c_in(self, sigma):
        sigma, sd = self._sd(sigma)
        return 1.0 / torch.sqrt(sigma ** 2 + sd ** 2)

    def c_noise(self, sigma):
        if not isinstance(sigma, torch.Tensor):
            sigma = torch.tensor(sigma, dtype=torch.float32)
        return 0.25 * torch.log(sigma.clamp(min=1e-8))

    @host_entry
    def forward(self, x_t, sigma, hint):
        """x0 = c_skip(sigma) x_t + c_out(sigma) F(c_in(sigma) x_t, t(sigma), hint); returns x_t itself when
        all(sigma <= sigma_min) (:81-82).  sigma: float tensor, 0-d / (B,) / (B,1,1,1)."""
        x_t = E._check_x(x_t)
        hint = E._check_x(hint)
        if not torch.is_floating_point(sigma):
            raise rt.CnbError("sigma must be a floating tensor (an integer sigma truncates sigma_data to 0 in the "
                              "reference, SURVEY.md Appendix A.11)")
        B = x_t.shape[0]
        dev = x_t.device
        sig = sigma.to(device=dev, dtype=torch.float32).reshape(-1)
        if sig.numel() == 1 and B > 1:
            sig = sig.expand(B)
        sig = sig.contiguous()
        coef = torch.empty((4, B), device=dev, dtype=torch.float32)
        t_index = torch.empty((B,), device=dev, dtype=torch.int64)
        flag = torch.empty((1,), device=dev, dtype=torch.int32)
        rt.check(rt.lib().cnb_edm_coeffs(sig.data_ptr(), B, float(self.sigma_data), float(self.sigma_min),
                                         coef.data_ptr(), t_index.data_ptr(), flag.data_ptr(), rt.stream()))
        if not getattr(self, "_skip_boundary_sync", False) and int(flag.item()):
            return x_t                # host sync, exactly like the reference's `if torch.all(...)` (:81-82)
        mode = rt.get_mode()
        # c_in * x_t on the way into the channels-last workspace, c_skip * x_t + c_out * F on the way out: one launch
        # each (round 1 ran scale_rows -> nchw_to_nhwc -> body -> nhwc_to_nchw -> scale_rows), bit-identical
        f_theta = student_body(self, ops.scale_nchw_to_nhwc(coef[0], x_t), t_index, hint, mode)
        return ops.edm_combine_to_nchw(coef[1], x_t, coef[2], f_theta)


class ConsistencyControlNetDistilled(nn.Module):
    """student + EMA teacher (+ optional DDPM teacher) container with the reference's attribute / state_dict names
    (`student.*`, `ema_teacher.*`, `ddpm_teacher.*`; :141-169).  Only the inference entry points are implemented."""

    def __init__(self, model_config, teacher_ckpt_path=None, device=None):
        super().__init__()
        self.student = ConsistencyControlNet(model_config)
        self.ema_teacher = deepcopy(self.student)
        self.ema_teacher.eval()
        self.ddpm_teacher = None
        if teacher_ckpt_path:
            self.ddpm_teacher = ControlNet(model_config, model_locked=True, model_ckpt=teacher_ckpt_path, device=device)
            self.ddpm_teacher.eval()
            from ..scheduler.linear_noise_scheduler import LinearNoiseScheduler
            self.teacher_scheduler = LinearNoiseScheduler(num_timesteps=1000, beta_start=0.0001, beta_end=0.02)
        self.sigma_min = model_config.get('sigma_min', 0.002)
        self.sigma_max = model_config.get('sigma_max', 80.0)
        self.num_timesteps = 1000
        self.ema_decay = 0.995

    def get_noise_schedule(self, num_steps, device=None):
        """Karras et al. schedule, rho = 7 (:179-196): host-side scalar table, not part of the per-sample path."""
        if device is None:
            device = next(self.parameters()).device
        rho = 7.0
        steps = torch.arange(num_steps, dtype=torch.float32, device=device)
        lo = torch.tensor(self.sigma_min, device=device) ** (1 / rho)
        hi = torch.tensor(self.sigma_max, device=device) ** (1 / rho)
        return (lo + steps / (num_steps - 1) * (hi - lo)) ** rho

    def forward(self, x_t, sigma, hint):
        return self.student(x_t, sigma, hint)

    def sigma_to_timestep(self, sigma):
        """Closest entry of the teacher's sigma schedule sqrt((1 - abar) / abar) (:230-257): a 1000-entry host-side
        table lookup (torch ops on the model's device, not per-sample work)."""
        if not isinstance(sigma, torch.Tensor):
            sigma = torch.tensor(sigma, dtype=torch.float32)
        device = next(self.parameters()).device
        sigma = sigma.to(device)
        if hasattr(self, 'teacher_scheduler'):
            acp = self.teacher_scheduler.alpha_cum_prod.to(device)
        else:
            betas = torch.linspace(0.0001, 0.02, self.num_timesteps, device=device)
            acp = torch.cumprod(1.0 - betas, dim=0)
        schedule = torch.sqrt((1 - acp) / acp)
        if sigma.dim() == 0:
            sigma = sigma.unsqueeze(0)
        t = torch.argmin(torch.abs(schedule.unsqueeze(0) - sigma.unsqueeze(-1)), dim=-1)
        return t.long().clamp(0, self.num_timesteps - 1)

    def get_ddpm_teacher_prediction(self, x_t, sigma, hint):
        """DDPM teacher's x_0 prediction at the timestep closest to sigma (:201-228)."""
        if self.ddpm_teacher is None:
            raise ValueError("DDPM teacher not initialized")
        from ._student_common import teacher_x0
        return teacher_x0(self.ddpm_teacher, self.teacher_scheduler, x_t, self.sigma_to_timestep(sigma.squeeze()), hint)

    def generate(self, hint, shape, num_steps=1, guidance_scale=1.0):
        """Single-step (:381-389) and multi-step (:390-409) sampling.  x_T / re-noising draws use torch's CUDA
        generator exactly like the reference (`torch.randn(shape, device=device)`), so seeded runs see the same
        noise stream; the arithmetic (student forward, x_0 + sigma_next * noise) runs on libcnb200 kernels."""
        device = next(self.parameters()).device
        with torch.no_grad():
            if num_steps == 1:
                x_T = torch.randn(shape, device=device)
                sigma = torch.full((shape[0],), self.sigma_max, device=device)
                return self.student(x_T, sigma, hint)
            sigmas = self.get_noise_schedule(num_steps + 1, device)
            x = torch.randn(shape, device=device)
            ones = torch.ones((shape[0],), device=device)
            for i in range(num_steps):
                x_0 = self.student(x, sigmas[i], hint)
                if i < num_steps - 1:
                    noise = torch.randn_like(x)
                    x = ops.scale_rows(ones, x_0, sigmas[i + 1].expand(shape[0]).contiguous(), noise)
                else:
                    x = x_0
            return x
