"""CPU oracle for the ControlNet denoising hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (CPU, fp32) of the reference's algorithm for the path named by
BASELINE.json.  It is imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, and only as the checker or as the timed CPU baseline - never by the product
package `controlnet-pytorch_b200/`, which has no CPU path at all.

Where the arithmetic lives: the reference is pure Python on top of PyTorch/ATen (requirements.txt:6
pins torch==2.3.1; this image has torch 2.11.0).  ATen is a third-party dependency that is not under
/root/reference, so the restatement below re-expresses each reference module as a function over a
flat ``state_dict`` (the reference's own key names) using the same ATen
primitives the reference's nn.Modules dispatch to (conv2d, conv_transpose2d, group_norm, silu, linear,
softmax) - it does NOT instantiate or import any reference class.  nn.MultiheadAttention is restated
explicitly (packed in_proj rows [q;k;v], per-head softmax(QK^T/sqrt(d))V, out_proj).

Pinning: parity is NOT pinned by the reference's own tests (test_distribution_matching.py:49-56,79-83
assert shapes only; there are no golden vectors).  The oracle is therefore pinned against outputs of
the reference itself, run in the build container by oracle/make_golden.py (imports /root/reference,
fills both sides with utils/synthetic.det_state_dict weights) and committed as tests/golden/*.npz;
tests/test_oracle.py re-checks the oracle against those vectors everywhere, and against the live
reference when /root/reference is present.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# architecture descriptors
# --------------------------------------------------------------------------------------------
def base_arch(cfg):
    """models/unet_base.py:297-339 - GroupNorm(8,.), 4 heads, attention in every block, last up = 16."""
    nd = len(cfg["down_channels"]) - 1
    return dict(down_channels=list(cfg["down_channels"]), mid_channels=list(cfg["mid_channels"]),
                down_sample=list(cfg["down_sample"]), t_emb_dim=cfg["time_emb_dim"],
                n_down=cfg["num_down_layers"], n_mid=cfg["num_mid_layers"], n_up=cfg["num_up_layers"],
                groups=8, heads=4, attn_down=[True] * nd, last_up=16)


def ldm_arch(cfg):
    """models/unet_cond_base.py:15-123 (unconditional path) on models/blocks.py."""
    return dict(down_channels=list(cfg["down_channels"]), mid_channels=list(cfg["mid_channels"]),
                down_sample=list(cfg["down_sample"]), t_emb_dim=cfg["time_emb_dim"],
                n_down=cfg["num_down_layers"], n_mid=cfg["num_mid_layers"], n_up=cfg["num_up_layers"],
                groups=cfg["norm_channels"], heads=cfg["num_heads"], attn_down=list(cfg["attn_down"]),
                last_up=cfg["conv_out_channels"])


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def time_embedding(t, dim):
    """models/unet_base.py:5-28 (duplicate models/blocks.py:5-28): [sin | cos] of t / 10000^(i/(D/2))."""
    t = torch.as_tensor(t).long()
    if t.dim() == 0:
        t = t.unsqueeze(0)
    half = dim // 2
    factor = 10000 ** (torch.arange(0, half, dtype=torch.float32) / half)
    arg = t[:, None].repeat(1, half) / factor
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)


def _gn(sd, key, x, groups):
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)


def _conv(sd, key, x, stride=1, padding=0):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)


def t_proj(sd, p, emb):
    """models/unet_base.py:313-317: Linear -> SiLU -> Linear."""
    h = F.linear(emb, sd[p + "t_proj.0.weight"], sd[p + "t_proj.0.bias"])
    return F.linear(F.silu(h), sd[p + "t_proj.2.weight"], sd[p + "t_proj.2.bias"])


def resnet(sd, p, j, x, temb, groups):
    """models/unet_base.py:96-100 (and :175-179, :274-278; blocks.py:119-125)."""
    h = _conv(sd, p + f"resnet_conv_first.{j}.2", F.silu(_gn(sd, p + f"resnet_conv_first.{j}.0", x, groups)),
              padding=1)
    if temb is not None:
        tb = F.linear(F.silu(temb), sd[p + f"t_emb_layers.{j}.1.weight"], sd[p + f"t_emb_layers.{j}.1.bias"])
        h = h + tb[:, :, None, None]
    h = _conv(sd, p + f"resnet_conv_second.{j}.2", F.silu(_gn(sd, p + f"resnet_conv_second.{j}.0", h, groups)),
              padding=1)
    return h + _conv(sd, p + f"residual_input_conv.{j}", x)


def mha(sd, p, x, heads):
    """nn.MultiheadAttention(E, heads, batch_first=True)(x, x, x)[0]  (SURVEY.md Appendix F)."""
    B, L, E = x.shape
    d = E // heads
    qkv = F.linear(x, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(B, L, heads, d).transpose(1, 2)
    k = k.reshape(B, L, heads, d).transpose(1, 2)
    v = v.reshape(B, L, heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = o.transpose(1, 2).reshape(B, L, E)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def self_attention(sd, p, j, x, groups, heads):
    """models/unet_base.py:103-109."""
    B, C, H, W = x.shape
    a = _gn(sd, p + f"attention_norms.{j}", x.reshape(B, C, H * W), groups).transpose(1, 2)
    y = mha(sd, p + f"attentions.{j}", a, heads)
    return x + y.transpose(1, 2).reshape(B, C, H, W)


def down_block(sd, p, x, temb, n_layers, groups, heads, attn, down_sample):
    """models/unet_base.py:91-112 ; models/blocks.py:115-150 (self-attention path)."""
    for j in range(n_layers):
        x = resnet(sd, p, j, x, temb, groups)
        if attn:
            x = self_attention(sd, p, j, x, groups, heads)
    if down_sample:
        x = _conv(sd, p + "down_sample_conv", x, stride=2, padding=1)
    return x


def mid_block(sd, p, x, temb, n_layers, groups, heads):
    """models/unet_base.py:171-199 ; models/blocks.py:230-271."""
    x = resnet(sd, p, 0, x, temb, groups)
    for j in range(n_layers):
        x = self_attention(sd, p, j, x, groups, heads)
        x = resnet(sd, p, j + 1, x, temb, groups)
    return x


def up_block(sd, p, x, skip, temb, n_layers, groups, heads, up_sample):
    """models/unet_base.py:267-289 ; models/blocks.py:465-503 (UpBlockUnet)."""
    if up_sample:
        x = F.conv_transpose2d(x, sd[p + "up_sample_conv.weight"], sd[p + "up_sample_conv.bias"],
                               stride=2, padding=1)
    if skip is not None:
        x = torch.cat([x, skip], dim=1)
    for j in range(n_layers):
        x = resnet(sd, p, j, x, temb, groups)
        x = self_attention(sd, p, j, x, groups, heads)
    return x


def vae_up_block(sd, p, x, n_layers, groups, heads, attn, up_sample):
    """models/blocks.py:347-374 (VAE UpBlock: ConvTranspose over in_channels, no skip, optional self-attention)."""
    if up_sample:
        x = F.conv_transpose2d(x, sd[p + "up_sample_conv.weight"], sd[p + "up_sample_conv.bias"],
                               stride=2, padding=1)
    for j in range(n_layers):
        x = resnet(sd, p, j, x, None, groups)
        if attn:
            x = self_attention(sd, p, j, x, groups, heads)
    return x


# --------------------------------------------------------------------------------------------
# VAE (models/vae.py) - the step after the sampling loop in tools/sample_ldm_controlnet.py:54-56
# --------------------------------------------------------------------------------------------
def vae_decode(sd, cfg, z):
    """models/vae.py:102-114: post_quant_conv -> decoder_conv_in -> MidBlocks (reversed) -> UpBlocks (reversed)
    -> GroupNorm -> SiLU -> decoder_conv_out."""
    dc, mc, g, heads = cfg["down_channels"], cfg["mid_channels"], cfg["norm_channels"], cfg["num_heads"]
    up_sample = list(reversed(cfg["down_sample"]))          # vae.py:33 builds it but the blocks index down_sample[i-1]
    del up_sample
    h = _conv(sd, "post_quant_conv", z)
    h = _conv(sd, "decoder_conv_in", h, padding=1)
    for k, _ in enumerate(reversed(range(1, len(mc)))):
        h = mid_block(sd, f"decoder_mids.{k}.", h, None, cfg["num_mid_layers"], g, heads)
    for k, i in enumerate(reversed(range(1, len(dc)))):
        h = vae_up_block(sd, f"decoder_layers.{k}.", h, cfg["num_up_layers"], g, heads, cfg["attn_down"][i - 1],
                         cfg["down_sample"][i - 1])
    h = F.silu(_gn(sd, "decoder_norm_out", h, g))
    return _conv(sd, "decoder_conv_out", h, padding=1)


def vae_encode(sd, cfg, x, noise=None):
    """models/vae.py:86-100.  Returns (sample, encoder_output); `noise` replaces the reference's
    torch.randn(mean.shape) (CPU default generator), sample is None when it is not given."""
    dc, mc, g, heads = cfg["down_channels"], cfg["mid_channels"], cfg["norm_channels"], cfg["num_heads"]
    h = _conv(sd, "encoder_conv_in", x, padding=1)
    for i in range(len(dc) - 1):
        h = down_block(sd, f"encoder_layers.{i}.", h, None, cfg["num_down_layers"], g, heads, cfg["attn_down"][i],
                       cfg["down_sample"][i])
    for i in range(len(mc) - 1):
        h = mid_block(sd, f"encoder_mids.{i}.", h, None, cfg["num_mid_layers"], g, heads)
    h = F.silu(_gn(sd, "encoder_norm_out", h, g))
    out = _conv(sd, "pre_quant_conv", _conv(sd, "encoder_conv_out", h, padding=1))
    if noise is None:
        return None, out
    mean, logvar = torch.chunk(out, 2, dim=1)
    return mean + torch.exp(0.5 * logvar) * noise, out


# --------------------------------------------------------------------------------------------
# U-Net pieces
# --------------------------------------------------------------------------------------------
def unet_temb(sd, p, arch, t):
    return t_proj(sd, p, time_embedding(t, arch["t_emb_dim"]))


def unet_encode(sd, p, arch, h, temb):
    """conv_in output -> (skips, bottleneck input): the down loop of Unet.forward (unet_base.py:355-359)."""
    skips = []
    for i in range(len(arch["down_channels"]) - 1):
        skips.append(h)
        h = down_block(sd, p + f"downs.{i}.", h, temb, arch["n_down"], arch["groups"], arch["heads"],
                       arch["attn_down"][i], arch["down_sample"][i])
    return skips, h


def unet_decode(sd, p, arch, h, skips, temb):
    """ups + norm_out + SiLU + conv_out (unet_base.py:366-374)."""
    skips = list(skips)
    nd = len(arch["down_channels"]) - 1
    for u, i in enumerate(reversed(range(nd))):
        h = up_block(sd, p + f"ups.{u}.", h, skips.pop(), temb, arch["n_up"], arch["groups"], arch["heads"],
                     arch["down_sample"][i])
    h = F.silu(_gn(sd, p + "norm_out", h, arch["groups"]))
    return _conv(sd, p + "conv_out", h, padding=1)


def unet_forward(sd, p, arch, x, t):
    """models/unet_base.py:341-374 / models/unet_cond_base.py:125-184 without conditioning."""
    temb = unet_temb(sd, p, arch, t)
    h = _conv(sd, p + "conv_in", x, padding=1)
    skips, h = unet_encode(sd, p, arch, h, temb)
    for i in range(len(arch["mid_channels"]) - 1):
        h = mid_block(sd, p + f"mids.{i}.", h, temb, arch["n_mid"], arch["groups"], arch["heads"])
    return unet_decode(sd, p, arch, h, skips, temb)


# --------------------------------------------------------------------------------------------
# hint encoders
# --------------------------------------------------------------------------------------------
def hint_block_ddpm(sd, p, hint):
    """models/controlnet.py:69-89 (same stack in the students: consistency_controlnet_distilled.py:21-31,
    distribution_matching_controlnet.py:101-111): 3x3 -> SiLU -> 3x3 -> SiLU -> 3x3 -> SiLU -> 1x1."""
    h = F.silu(_conv(sd, p + "0", hint, padding=1))
    h = F.silu(_conv(sd, p + "2", h, padding=1))
    h = F.silu(_conv(sd, p + "4", h, padding=1))
    return _conv(sd, p + "6", h)


def hint_block_ldm(sd, p, hint):
    """models/controlnet_ldm.py:45-79: 3x3+SiLU, then per halving [3x3 s2, SiLU, 3x3], then 3x3+SiLU+1x1."""
    n_stage = 0
    while (p + f"{n_stage}.0.weight") in sd:
        n_stage += 1
    h = F.silu(_conv(sd, p + "0.0", hint, padding=1))
    for s in range(1, n_stage - 1):
        h = F.silu(_conv(sd, p + f"{s}.0", h, stride=2, padding=1))
        h = _conv(sd, p + f"{s}.2", h, padding=1)
    last = n_stage - 1
    h = F.silu(_conv(sd, p + f"{last}.0", h, padding=1))
    return _conv(sd, p + f"{last}.2", h)


# --------------------------------------------------------------------------------------------
# ControlNet forwards
# --------------------------------------------------------------------------------------------
def _controlnet(sd, arch, x, t, hint_feat, tp, cp, zp_down, zp_mid):
    """Shared wiring of models/controlnet.py:158-225 and models/controlnet_ldm.py:117-179."""
    temb_t = unet_temb(sd, tp, arch, t)
    a = _conv(sd, tp + "conv_in", x, padding=1)
    t_skips, a = unet_encode(sd, tp, arch, a, temb_t)

    temb_c = unet_temb(sd, cp, arch, t)
    c = _conv(sd, cp + "conv_in", x, padding=1) + hint_feat
    c_skips = []
    for i in range(len(arch["down_channels"]) - 1):
        c_skips.append(_conv(sd, zp_down + f"{i}", c))
        c = down_block(sd, cp + f"downs.{i}.", c, temb_c, arch["n_down"], arch["groups"], arch["heads"],
                       arch["attn_down"][i], arch["down_sample"][i])
    for i in range(len(arch["mid_channels"]) - 1):
        c = mid_block(sd, cp + f"mids.{i}.", c, temb_c, arch["n_mid"], arch["groups"], arch["heads"])
        a = mid_block(sd, tp + f"mids.{i}.", a, temb_t, arch["n_mid"], arch["groups"], arch["heads"])
        a = a + _conv(sd, zp_mid + f"{i}", c)
    skips = [cs + ts for cs, ts in zip(c_skips, t_skips)]
    return unet_decode(sd, tp, arch, a, skips, temb_t)


def controlnet_ddpm_forward(sd, cfg, x, t, hint):
    """models/controlnet.py:158-225."""
    arch = base_arch(cfg)
    hf = hint_block_ddpm(sd, "control_copy_unet_hint_block.", hint)
    return _controlnet(sd, arch, x, t, hf, "trained_unet.", "control_copy_unet.",
                       "control_copy_unet_down_zero_convs.", "control_copy_unet_mid_zero_convs.")


def controlnet_ldm_forward(sd, cfg, x, t, hint):
    """models/controlnet_ldm.py:117-179."""
    arch = ldm_arch(cfg)
    hf = hint_block_ldm(sd, "control_unet_hint_block.", hint)
    return _controlnet(sd, arch, x, t, hf, "trained_unet.", "control_unet.",
                       "control_unet_down_zero_convs.", "control_unet_mid_zero_convs.")


# --------------------------------------------------------------------------------------------
# single-step students
# --------------------------------------------------------------------------------------------
def _student_unet(sd, cfg, x_in, temb, hint):
    arch = base_arch(cfg)
    h = _conv(sd, "unet.conv_in", x_in, padding=1) + hint_block_ddpm(sd, "hint_block.", hint)
    skips, h = unet_encode(sd, "unet.", arch, h, temb)
    for i in range(len(arch["mid_channels"]) - 1):
        h = mid_block(sd, "unet." + f"mids.{i}.", h, temb, arch["n_mid"], arch["groups"], arch["heads"])
    return unet_decode(sd, "unet.", arch, h, skips, temb)


def _student_temb(sd, idx, dim):
    """own t_proj = SiLU -> Linear (consistency_controlnet_distilled.py:35-38)."""
    return F.linear(F.silu(time_embedding(idx, dim)), sd["t_proj.1.weight"], sd["t_proj.1.bias"])


def dm_forward(sd, cfg, x_t, t, hint):
    """models/distribution_matching_controlnet.py:120-159."""
    temb = _student_temb(sd, torch.as_tensor(t).long(), cfg["time_emb_dim"])
    return _student_unet(sd, cfg, x_t, temb, hint)


def consistency_forward(sd, cfg, x_t, sigma, hint):
    """models/consistency_controlnet_distilled.py:76-134 (+ c_skip/c_out/c_in/c_noise :45-74)."""
    sigma_min = cfg.get("sigma_min", 0.002)
    sigma_data = cfg.get("sigma_data", 0.5)
    sigma = torch.as_tensor(sigma, dtype=torch.float32)
    if torch.all(sigma <= sigma_min):
        return x_t
    if sigma.dim() == 0:
        sigma = sigma.unsqueeze(0)
    if sigma.dim() == 1:
        sigma = sigma.view(-1, 1, 1, 1)
    sd_t = torch.tensor(sigma_data, dtype=sigma.dtype)
    c_in = 1.0 / torch.sqrt(sigma ** 2 + sd_t ** 2)
    c_noise = 0.25 * torch.log(sigma.squeeze().clamp(min=1e-8))
    idx = (c_noise * 1000).long().clamp(0, 999)
    temb = _student_temb(sd, idx, cfg["time_emb_dim"])
    f_theta = _student_unet(sd, cfg, c_in * x_t, temb, hint)
    c_skip = sd_t ** 2 / (sigma ** 2 + sd_t ** 2)
    c_out = sigma * sd_t / torch.sqrt(sigma ** 2 + sd_t ** 2)
    return c_skip * x_t + c_out * f_theta


# --------------------------------------------------------------------------------------------
# scheduler
# --------------------------------------------------------------------------------------------
class SchedulerOracle:
    """scheduler/linear_noise_scheduler.py:8-23 (tables) and :49-77 (sample_prev_timestep).
    ``z`` is passed in instead of drawn from the CPU generator (:71) so both sides share the noise."""

    def __init__(self, num_timesteps, beta_start, beta_end, ldm_scheduler=False):
        if ldm_scheduler:
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_timesteps) ** 2
        else:
            self.betas = torch.linspace(beta_start, beta_end, num_timesteps)
        self.alphas = 1. - self.betas
        self.alpha_cum_prod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alpha_cum_prod = torch.sqrt(self.alpha_cum_prod)
        self.sqrt_one_minus_alpha_cum_prod = torch.sqrt(1 - self.alpha_cum_prod)

    def sample_prev_timestep(self, xt, eps, t, z=None):
        t = int(t)
        x0 = (xt - self.sqrt_one_minus_alpha_cum_prod[t] * eps) / torch.sqrt(self.alpha_cum_prod[t])
        x0 = torch.clamp(x0, -1., 1.)
        mean = xt - (self.betas[t] * eps) / self.sqrt_one_minus_alpha_cum_prod[t]
        mean = mean / torch.sqrt(self.alphas[t])
        if t == 0:
            return mean, x0
        variance = (1 - self.alpha_cum_prod[t - 1]) / (1.0 - self.alpha_cum_prod[t])
        variance = variance * self.betas[t]
        sigma = variance ** 0.5
        return mean + sigma * z, x0


def teacher_x0(sd, cfg, sched, x_t, t, hint, prefix="teacher."):
    """distribution_matching_controlnet.py:191-216 / consistency_controlnet_distilled.py:201-228: the DDPM ControlNet
    teacher's noise prediction converted to a clamped x_0 with per-sample t."""
    tsd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    t = torch.as_tensor(t).long()
    if t.dim() == 0:
        t = t.unsqueeze(0)
    eps = controlnet_ddpm_forward(tsd, cfg, x_t, t, hint)
    B = x_t.shape[0]
    s1 = sched.sqrt_one_minus_alpha_cum_prod[t].reshape(B, 1, 1, 1)
    s2 = sched.sqrt_alpha_cum_prod[t].reshape(B, 1, 1, 1)
    return torch.clamp((x_t - s1 * eps) / s2, -1., 1.)


def sigma_to_timestep(sched, sigma):
    """consistency_controlnet_distilled.py:230-257 with the teacher's schedule."""
    sigma = torch.as_tensor(sigma, dtype=torch.float32)
    if sigma.dim() == 0:
        sigma = sigma.unsqueeze(0)
    schedule = torch.sqrt((1 - sched.alpha_cum_prod) / sched.alpha_cum_prod)
    return torch.argmin(torch.abs(schedule.unsqueeze(0) - sigma.unsqueeze(-1)), dim=-1).long().clamp(0, 999)


def ddpm_sample(forward_fn, sched, x_T, hint, steps, zs):
    """The loop of tools/sample_ddpm_controlnet.py:43-51 for t = steps-1 .. 0, with the
    per-step z injected: zs[k] is used at the k-th iteration (none at t == 0)."""
    xt = x_T
    x0 = None
    for k, t in enumerate(reversed(range(steps))):
        eps = forward_fn(xt, torch.as_tensor(t).unsqueeze(0), hint)
        xt, x0 = sched.sample_prev_timestep(xt, eps, t, zs[k] if t > 0 else None)
    return xt, x0
