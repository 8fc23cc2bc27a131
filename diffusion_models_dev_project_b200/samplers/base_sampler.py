"""Reverse-time sampling loop.

Contract of the reference's ``BaseSampler`` (src/samplers/base_sampler.py:19-123), kept so that its
drivers and ``sample_kwargs`` dictionaries work unchanged:

* constructor ``BaseSampler(score, sde, predictor, sample_kwargs, device)``;
* ``sample(logg_kwargs={}, logging=True) -> x_mean`` (the second return value of the last
  predictor call, i.e. the Tweedie estimate for the DDS / SCD predictors);
* time grid: ``linspace(1, eps, num_steps)`` for VE/VP schedules; for DDPM the (optionally
  time-travelling) index list of ``_schedule_jump`` turned into ``(t, t_prev)`` pairs on the
  ``num_steps``-strided grid, truncated by ``early_stopping_pct`` when that key is present;
* the chain starts from ``sde.prior_sampling`` (drawn on the CPU generator, then moved), and every
  step calls ``predictor(score=, sde=, x=, time_step=, step_size=, datafitscale=, **sample_kwargs['predictor'])``
  with per-sample time tensors; ``adapt_freq`` toggles ``use_adapt`` in the predictor kwargs.

The loop itself is plain Python; everything heavy happens inside the predictor.
"""
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from .utils import _schedule_jump
from ..utils.metrics import PSNR
from ..utils.sde import SDE, _EPSILON_PRED_CLASSES, _SCORE_PRED_CLASSES


class _BoardLog:
    """TensorBoard side channel of ``sample(logging=True)``: start images, the running reconstruction
    every ``num_img_in_log`` steps with its PSNR, the final reconstruction."""

    def __init__(self, logg_kwargs: Dict):
        import torchvision
        from torch.utils.tensorboard import SummaryWriter
        self.kw = logg_kwargs
        self.writer = SummaryWriter(log_dir=os.path.join(logg_kwargs['log_dir'], str(logg_kwargs['sample_num'])))
        self._grid = lambda img: torchvision.utils.make_grid(img, normalize=True, scale_each=True)
        self.psnr = None

    def image(self, tag: str, img: Tensor, step: int = 0):
        self.writer.add_image(tag, self._grid(img), global_step=step)

    def start(self, x0: Tensor):
        self.image('init_x', x0)
        for tag in ('ground_truth', 'filtbackproj'):
            if self.kw[tag] is not None:
                self.image(tag, self.kw[tag].squeeze())

    def step(self, i: int, offset: int, x_mean: Tensor):
        if (i - offset) % self.kw['num_img_in_log'] == 0:
            self.image('reco', x_mean.squeeze(), i)
            self.psnr = PSNR(x_mean[0, 0].cpu().numpy(), self.kw['ground_truth'][0, 0].cpu().numpy())
        if self.psnr is not None:
            self.writer.add_scalar('PSNR', self.psnr, i)


class BaseSampler:
    def __init__(self, score, sde: SDE, predictor: callable, sample_kwargs: Dict,
                 device: Optional[Any] = None) -> None:
        self.score, self.sde, self.predictor = score, sde, predictor
        self.sample_kwargs = sample_kwargs
        self.device = device

    # ---------------------------------------------------------------- time grid ----
    def _schedule(self) -> Tuple[Any, List]:
        """``(time_steps, steps)``: the grid that defines the step size, and what the loop iterates."""
        kw = self.sample_kwargs
        n = kw['num_steps']
        if isinstance(self.sde, tuple(_SCORE_PRED_CLASSES)):
            grid = np.linspace(1., kw['eps'], n)
            return grid, [float(t) for t in grid]
        if isinstance(self.sde, tuple(_EPSILON_PRED_CLASSES)):
            assert self.sde.num_steps >= n
            stride = self.sde.num_steps // n
            idx = _schedule_jump(n, kw['travel_length'], kw['travel_repeat'])
            pairs = [(cur * stride, nxt * stride if nxt > 0 else -1) for cur, nxt in zip(idx[:-1], idx[1:])]
            if 'early_stopping_pct' in kw:
                pairs = pairs[:int(kw['early_stopping_pct'] * len(pairs))]
            return idx, pairs
        raise NotImplementedError(self.sde.__class__)

    # --------------------------------------------------------------------- loop ----
    def sample(self, logg_kwargs: Dict = {}, logging: bool = True) -> Tensor:
        kw = self.sample_kwargs
        board = _BoardLog(logg_kwargs) if logging else None
        time_steps, steps = self._schedule()
        step_size = time_steps[0] - time_steps[1]
        x = self.sde.prior_sampling([kw['batch_size'], *kw['im_shape']]).to(self.device)
        if board:
            board.start(x)
        ones = torch.ones(kw['batch_size'], device=self.device)
        x_mean = x
        adapt_every = kw.get('adapt_freq', None)
        for i, step in enumerate(steps):
            if isinstance(step, tuple):                       # DDPM: (t, t_prev)
                time_step, datafitscale = (ones * step[0], ones * step[1]), 1.
            else:                                             # VE / VP: continuous t
                time_step, datafitscale = ones * step, step / kw['num_steps']
            if adapt_every is not None:
                kw['predictor']['use_adapt'] = (i % adapt_every == 0)
            x, x_mean = self.predictor(score=self.score, sde=self.sde, x=x, time_step=time_step,
                                       step_size=step_size, datafitscale=datafitscale, **kw['predictor'])
            if board:
                board.step(i, kw['start_time_step'], x_mean)
        if board:
            board.image('final_reco', x_mean.squeeze())
        return x_mean
