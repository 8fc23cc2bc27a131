"""BASELINE config 5: SCD adapted sampling (LoRA fine-tune inside the reverse chain) at 256x256, batch 1.

    python tools/adapted_bench.py [--steps 3] [--num_optim_step 10] [--cg_iter 1]

Defaults are the reference's run_adapted_sampling.py defaults (:17-37): 50 reverse steps per sample, adaptation
at every step with 10 Adam steps (lr 1e-3) on rank-4 LoRA branches, loss mean((A x - y)^2) + 1e-6 TV evaluated
on the CG(1) data-consistency result, eta 0.85, gamma 0.01, `--add_cg --dc_type cg`.  Each adapted reverse step
differentiates 10 times through Tweedie -> CG -> A / A* (CUDA kernels behind autograd Functions with ODL's
gradient pairing) and the fused adaptation loss.  Prints one JSON line: device time per reverse step (CUDA
events) with the full-size UNet caller (random init, LoRA stand-in of bench_support/lora.py) and with the
weight-free score of tests/scorenet.py (the cost of this repository's path alone), and samples/s.
"""
import argparse
import copy
import json
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import diffusion_models_dev_project_b200 as pkg                      # noqa: E402
from diffusion_models_dev_project_b200 import _lib                  # noqa: E402
from bench_support.adm_unet import aapm_unet                        # noqa: E402
from bench_support.lora import inject_trainable_lora                # noqa: E402
from bench_support.phantoms import disk_ellipses                    # noqa: E402


def time_steps(sampler_factory, n_steps, warm):
    """ms per reverse step of the sampler's predictor (adaptation at every step), CUDA events."""
    sampler = sampler_factory()
    kw = sampler.sample_kwargs
    dev = torch.device('cuda')
    sde = sampler.sde
    skip = sde.num_steps // kw['num_steps']
    x = sde.prior_sampling([kw['batch_size'], *kw['im_shape']]).to(dev)
    ones = torch.ones(kw['batch_size'], device=dev)
    pred_kw = dict(kw['predictor'], use_adapt=True)

    def step(i, x):
        t = (kw['num_steps'] - 1 - i) * skip
        out, _ = sampler.predictor(score=sampler.score, sde=sde, x=x, time_step=(ones * t, ones * max(t - skip, -1)),
                                   step_size=1, datafitscale=1., **pred_kw)
        return out
    for i in range(warm):
        x = step(i, x)
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.scd_launch_count_reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n_steps):
        x = step(warm + i, x)
    b.record()
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    return a.elapsed_time(b) / n_steps, int(lib.scd_launch_count()) // n_steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--num_steps', type=int, default=50)
    ap.add_argument('--num_optim_step', type=int, default=10)
    ap.add_argument('--cg_iter', type=int, default=1)
    ap.add_argument('--lora_rank', type=int, default=4)
    a = ap.parse_args()
    dev = torch.device('cuda')
    args = NS(method='dds', num_steps=a.num_steps, adapt_freq=1, eta=0.85, gamma=0.01, adaptation='lora',
              lora_include_blocks=None, lora_rank=a.lora_rank, tv_penalty=1e-6, num_optim_step=a.num_optim_step,
              lr=1e-3, add_cg=True, dc_type='cg', cg_iter=a.cg_iter, early_stopping_pct=1.0)
    config = NS(device=dev, sampling=NS(batch_size=1, eps=1e-3, travel_length=1, travel_repeat=1), model=NS(in_channels=1))
    rt = pkg.B200RayTrafo((256, 256), 60)
    sde = pkg.DDPM()
    gt = torch.from_numpy(disk_ellipses(1, 256, seed=1)).to(dev)
    y = pkg.simulate(gt, rt, 0.01, rng=np.random.default_rng(1))

    def with_unet():
        torch.manual_seed(0)
        score = aapm_unet().to(dev).eval()
        return pkg.get_standard_adapted_sampler(args=args, config=config, score=score, sde=sde, ray_trafo=rt,
                                                observation=y, device=dev, lora_inject_fn=inject_trainable_lora)

    def path_only():
        from scorenet import AdaptableScore
        a2 = copy.copy(args)
        a2.adaptation = 'full'
        return pkg.get_standard_adapted_sampler(args=a2, config=config, score=AdaptableScore().to(dev), sde=sde,
                                                ray_trafo=rt, observation=y, device=dev)
    ms_path, launches = time_steps(path_only, a.steps, a.warmup)
    ms_unet, _ = time_steps(with_unet, a.steps, a.warmup)
    print(json.dumps({
        'workload': 'SCD adapted sampling 256x256, 60 angles, batch 1, %d reverse steps/sample, %d Adam steps per reverse '
                    'step, LoRA rank %d, CG(%d), tv 1e-6, eta 0.85' % (a.num_steps, a.num_optim_step, a.lora_rank, a.cg_iter),
        'ms_per_reverse_step_full_unet': ms_unet,
        'samples_per_s_full_unet': 1e3 / (ms_unet * a.num_steps),
        'ms_per_reverse_step_path_only': ms_path,
        'path_share_of_step': ms_path / ms_unet,
        'library_launches_per_reverse_step': launches,
        'score_model': 'aapm ADM UNet, random init, LoRA stand-in (bench_support); path-only: tests/scorenet.AdaptableScore',
    }))


if __name__ == '__main__':
    main()
