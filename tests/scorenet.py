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
