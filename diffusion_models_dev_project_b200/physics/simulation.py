"""Measurement simulation ``y = A x + white noise`` (reference src/physics/simulation.py:12-23)."""
import numpy as np
import torch
from torch import Tensor


def simulate(x: Tensor, ray_trafo, white_noise_rel_stddev: float, rng=None,
             return_noise_level: bool = False):
    """Noise level = ``rel_stddev * mean|A x|``; the noise comes from a numpy
    generator (pass a seeded one for reproducibility), exactly as in the reference."""
    observation = ray_trafo(x)
    if rng is None:
        rng = np.random.default_rng()
    noise_level = white_noise_rel_stddev * torch.mean(torch.abs(observation)).item()
    noise = torch.from_numpy(rng.normal(scale=noise_level, size=observation.shape)).to(
        dtype=observation.dtype, device=observation.device)
    noisy_observation = observation + noise
    return (noisy_observation, noise_level) if return_noise_level else noisy_observation
