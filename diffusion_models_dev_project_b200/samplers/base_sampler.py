"""Reverse-time sampling loop.

Same constructor, ``sample_kwargs`` keys and ``sample(logg_kwargs, logging)``
contract as the reference's ``BaseSampler`` (src/samplers/base_sampler.py:19-123):
the schedule, the CPU-generator prior draw, the per-step predictor call with
``time_step=(t, t_prev)`` tensors and the returned ``x_mean`` are identical.
The loop stays Python; everything heavy happens inside the predictor.
"""
import os
from typing import Any, Dict, Optional

import numpy as np
import torch
from torch import Tensor

from .utils import _schedule_jump
from ..utils.metrics import PSNR
from ..utils.sde import SDE, _EPSILON_PRED_CLASSES, _SCORE_PRED_CLASSES


class BaseSampler:
    def __init__(self, score, sde: SDE, predictor: callable, sample_kwargs: Dict,
                 device: Optional[Any] = None) -> None:
        self.score = score
        self.sde = sde
        self.predictor = predictor
        self.sample_kwargs = sample_kwargs
        self.device = device

    def _schedule(self):
        kw = self.sample_kwargs
        num_steps = kw['num_steps']
        if any(isinstance(self.sde, c) for c in _SCORE_PRED_CLASSES):
            time_steps = np.linspace(1., kw['eps'], num_steps)
            return time_steps, list(time_steps)
        if any(isinstance(self.sde, c) for c in _EPSILON_PRED_CLASSES):
            assert self.sde.num_steps >= num_steps
            skip = self.sde.num_steps // num_steps
            time_steps = _schedule_jump(num_steps, kw['travel_length'], kw['travel_repeat'])
            pairs = [(i * skip, j * skip if j > 0 else -1) for i, j in zip(time_steps[:-1], time_steps[1:])]
            if 'early_stopping_pct' in kw:
                pairs = pairs[:int(kw['early_stopping_pct'] * len(pairs))]
            return time_steps, pairs
        raise NotImplementedError(self.sde.__class__)

    def sample(self, logg_kwargs: Dict = {}, logging: bool = True) -> Tensor:
        kw = self.sample_kwargs
        writer = None
        if logging:
            import torchvision
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(log_dir=os.path.join(logg_kwargs['log_dir'], str(logg_kwargs['sample_num'])))

            def grid(img):
                return torchvision.utils.make_grid(img, normalize=True, scale_each=True)

        time_steps, steps = self._schedule()
        step_size = time_steps[0] - time_steps[1]
        x = self.sde.prior_sampling([kw['batch_size'], *kw['im_shape']]).to(self.device)

        if logging:
            writer.add_image('init_x', grid(x), global_step=0)
            if logg_kwargs['ground_truth'] is not None:
                writer.add_image('ground_truth', grid(logg_kwargs['ground_truth'].squeeze()), global_step=0)
            if logg_kwargs['filtbackproj'] is not None:
                writer.add_image('filtbackproj', grid(logg_kwargs['filtbackproj'].squeeze()), global_step=0)

        ones_vec = torch.ones(kw['batch_size'], device=self.device)
        x_mean = x
        psnr = None
        for i, step in enumerate(steps):
            if isinstance(step, tuple):
                time_step = (ones_vec * step[0], ones_vec * step[1])      # (t, t_prev)
                datafitscale = 1.
            else:
                time_step = ones_vec * float(step)
                datafitscale = float(step) / kw['num_steps']

            if kw.get('adapt_freq', None) is not None:
                kw['predictor'].update({'use_adapt': i % kw['adapt_freq'] == 0})

            x, x_mean = self.predictor(
                score=self.score, sde=self.sde, x=x, time_step=time_step, step_size=step_size,
                datafitscale=datafitscale, **kw['predictor'])

            if logging:
                if (i - kw['start_time_step']) % logg_kwargs['num_img_in_log'] == 0:
                    writer.add_image('reco', grid(x_mean.squeeze()), i)
                    psnr = PSNR(x_mean[0, 0].cpu().numpy(), logg_kwargs['ground_truth'][0, 0].cpu().numpy())
                if psnr is not None:
                    writer.add_scalar('PSNR', psnr, i)

        if logging:
            writer.add_image('final_reco', grid(x_mean.squeeze()), global_step=0)
        return x_mean
