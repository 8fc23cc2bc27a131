"""Noise schedules used by the samplers.

Same classes and methods as reference src/utils/sde.py (``SDE``, ``VESDE``,
``VPSDE``, ``DDPM`` with ``marginal_prob_mean/std``, ``prior_sampling`` ...).
``DDPM`` additionally exposes :meth:`DDPM.alpha_bar_table`: the fp32 alpha-bar
table the fused kernels index with ``t + 1`` on the device -- it is produced by
the very expression of ``DDPM._compute_alpha_cumprod`` (reference sde.py:172-174),
once, instead of five times per reverse step.
"""
import abc

import numpy as np
import torch


class SDE(abc.ABC):
    """Interface of a forward noising process; all methods take a batch of times."""

    def diffusion_coeff(self, t):
        raise NotImplementedError

    def sde(self, x, t):
        raise NotImplementedError

    def marginal_prob(self, x, t):
        raise NotImplementedError

    def marginal_prob_std(self, t):
        raise NotImplementedError

    def marginal_prob_mean(self, t):
        raise NotImplementedError

    def prior_sampling(self, shape):
        raise NotImplementedError


class VESDE(SDE):
    """Variance-exploding SDE, sigma(t) = sigma_min (sigma_max/sigma_min)^t (reference sde.py:55-104)."""

    def __init__(self, sigma_min: float = 0.01, sigma_max: float = 50):
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max

    def marginal_prob_std(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def marginal_prob_mean(self, t):
        return torch.ones_like(t)

    def diffusion_coeff(self, t):
        log_ratio = np.log(self.sigma_max) - np.log(self.sigma_min)
        return self.marginal_prob_std(t) * torch.sqrt(torch.tensor(2 * log_ratio, device=t.device))

    def sde(self, x, t):
        return torch.zeros_like(x), self.diffusion_coeff(t)

    def marginal_prob(self, x, t):
        return x, self.marginal_prob_std(t)

    def prior_sampling(self, shape):
        return torch.randn(*shape) * self.sigma_max


class VPSDE(SDE):
    """Variance-preserving SDE with linear beta(t) (reference sde.py:106-157)."""

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20):
        self.beta_min = beta_min
        self.beta_max = beta_max

    def _beta(self, t):
        return self.beta_min + t * (self.beta_max - self.beta_min)

    def _log_mean_coeff(self, t):
        return -0.25 * t ** 2 * (self.beta_max - self.beta_min) - 0.5 * t * self.beta_min

    def diffusion_coeff(self, t):
        return torch.sqrt(self._beta(t))

    def sde(self, x, t):
        return -0.5 * self._beta(t)[:, None, None, None] * x, self.diffusion_coeff(t)

    def marginal_prob_mean(self, t):
        return torch.exp(self._log_mean_coeff(t))

    def marginal_prob_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self._log_mean_coeff(t)))

    def marginal_prob(self, x, t):
        return torch.exp(self._log_mean_coeff(t)[:, None, None, None]) * x, self.marginal_prob_std(t)

    def prior_sampling(self, shape):
        return torch.randn(*shape)


class DDPM(SDE):
    """Discrete DDPM schedule: linear betas in fp64, alpha-bar by cumulative product;
    time index ``t`` selects entry ``t + 1`` so that ``t = -1`` gives alpha-bar = 1
    (reference sde.py:159-194)."""

    def __init__(self, beta_min: float = 0.0001, beta_max: float = 0.02, num_steps: int = 1000):
        self.beta_min = beta_min
        self.beta_max = beta_max
        self.num_steps = num_steps
        self.betas = torch.from_numpy(
            np.linspace(self.beta_min, self.beta_max, self.num_steps, dtype=np.float64))
        assert self.betas.dim() == 1, 'betas must be 1-D'
        assert (self.betas > 0).all() and (self.betas <= 1).all()
        self.alphas = 1.0 - self.betas
        self._tables = {}

    def alpha_bar_table(self, device=None) -> torch.Tensor:
        """fp32 tensor ``[num_steps + 1]``: entry ``k`` is alpha-bar at time ``k - 1``.

        The fp64 cumulative product is formed on the host and the rounded fp32 table moved to ``device`` once.
        The reference forms it on ``t.device`` at every call (src/utils/sde.py:172-174): bit-identical to this
        table when the reference runs on CPU (the golden vectors); a reference running on a GPU may differ in the
        last bit (the fp64 ``cumprod`` of the two devices need not associate the same way)."""
        key = str(device)
        tab = self._tables.get(key)
        if tab is None:
            betas = torch.cat([torch.zeros(1), self.betas], dim=0)          # promotes to fp64
            tab = (1 - betas).cumprod(dim=0).to(torch.float32)
            if device is not None:
                tab = tab.to(device)
            self._tables[key] = tab.contiguous()
            tab = self._tables[key]
        return tab

    def _compute_alpha_cumprod(self, t):
        return self.alpha_bar_table(t.device).index_select(0, t.long() + 1)

    def diffusion_coeff(self):
        raise NotImplementedError

    def sde(self):
        raise NotImplementedError

    def marginal_prob_std(self, t):
        return (1. - self._compute_alpha_cumprod(t)).pow(.5)

    def marginal_prob_mean(self, t):
        return self._compute_alpha_cumprod(t).pow(.5)

    def marginal_prob(self, x, t):
        return x * self.marginal_prob_mean(t)[:, None, None, None], self.marginal_prob_std(t)

    def prior_sampling(self, shape):
        # drawn on the CPU generator, like the reference (sde.py:193-194), so that a
        # fixed seed gives the same chain start on both implementations
        return torch.randn(*shape)


_EPSILON_PRED_CLASSES = [DDPM]
_SCORE_PRED_CLASSES = [VPSDE, VESDE]
