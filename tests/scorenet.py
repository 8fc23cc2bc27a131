"""Deterministic stand-in for the score network used by parity tests and golden generation.

The samplers treat the score model as a black box ``score(x, t) -> eps_hat``.  This one
has no learned weights: its Tweedie estimate is a box-blurred, clamped version of the
current iterate, which makes the DDS chain behave like a plug-and-play reconstruction
(so PSNR against the ground truth is meaningful) while staying bit-reproducible across
devices (only elementwise ops and avg_pool2d in fp32 -- no cuDNN/TF32 convolution).
"""
import numpy as np
import torch
import torch.nn.functional as F


class BlurScore(torch.nn.Module):
    def __init__(self, beta_min=1e-4, beta_max=0.02, num_steps=1000):
        super().__init__()
        betas = torch.from_numpy(np.linspace(beta_min, beta_max, num_steps, dtype=np.float64))
        betas = torch.cat([torch.zeros(1, dtype=torch.float64), betas])
        self.register_buffer('abar', (1 - betas).cumprod(0).to(torch.float32))

    def forward(self, x, t):
        ab = self.abar.index_select(0, t.long() + 1)[:, None, None, None]
        m, sd = ab.sqrt(), (1 - ab).sqrt()
        z = x / m
        den = F.avg_pool2d(z, 3, stride=1, padding=1, count_include_pad=True).clamp(0., 1.)
        den = 0.5 * den + 0.5 * z.clamp(-0.25, 1.25)
        return (x - m * den) / sd


class LoraInjectedConv2d(torch.nn.Module):
    """3x3 convolution (1 -> 1 channel) with a rank-``r`` side branch and a runtime ``scale`` switch.

    The class NAME is what the reference's ``_has_lora_active`` / ``_tune_lora_scale`` look for
    (reference src/samplers/utils.py:262-278), so its SCD code runs on this stand-in unmodified.
    The convolution is written as unfold + matmul in fp32 (no cuDNN / TF32), initialised from an
    explicit generator so that every process builds the same weights."""

    def __init__(self, r=2, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.weight = torch.nn.Parameter(0.05 * torch.randn(1, 9, generator=g))
        self.lora_down = torch.nn.Parameter(0.2 * torch.randn(r, 9, generator=g))
        self.lora_up = torch.nn.Parameter(0.1 * torch.randn(1, r, generator=g))
        self.bias = torch.nn.Parameter(torch.zeros(1))
        self.scale = 1.0

    def forward(self, z):
        b, _, h, w = z.shape
        cols = F.unfold(z, 3, padding=1)                          # [B, 9, H*W]
        out = self.weight @ cols + self.scale * (self.lora_up @ (self.lora_down @ cols))
        return out.reshape(b, 1, h, w) + self.bias


class AdaptableScore(BlurScore):
    """BlurScore plus a small trainable correction of the denoised estimate: the score model of the
    SCD (adapted sampling) parity fixture."""

    def __init__(self, r=2, seed=0):
        super().__init__()
        self.adapter = LoraInjectedConv2d(r=r, seed=seed)

    def forward(self, x, t):
        ab = self.abar.index_select(0, t.long() + 1)[:, None, None, None]
        m, sd = ab.sqrt(), (1 - ab).sqrt()
        z = x / m
        den = F.avg_pool2d(z, 3, stride=1, padding=1, count_include_pad=True).clamp(0., 1.)
        den = 0.5 * den + 0.5 * z.clamp(-0.25, 1.25) + 0.1 * torch.tanh(self.adapter(z))
        return (x - m * den) / sd
