"""Measurement simulation ``y = A x + white noise`` (reference src/physics/simulation.py:12-74)."""
from typing import Any, Iterator, Optional, Tuple

import numpy as np
import torch
from torch import Tensor


def simulate(x: Tensor, ray_trafo, white_noise_rel_stddev: float, rng=None,
             return_noise_level: bool = False):
    """Noise level = ``rel_stddev * mean|A x|``; the noise comes from a numpy generator (pass a
    seeded one for reproducibility), as in the reference (:12-23).

    The reference reads the noise level back to the host (``.item()``) before it draws
    ``rng.normal(scale=noise_level)``.  numpy forms that draw as ``scale * standard_normal`` in
    float64, so here the standard-normal field is drawn first and scaled on the device by the
    float64 noise level: same random stream, same values, and no host synchronisation between the
    projection and the noise (the level is only read back when ``return_noise_level`` asks for it)."""
    observation = ray_trafo(x)
    if rng is None:
        rng = np.random.default_rng()
    noise_level = white_noise_rel_stddev * torch.mean(torch.abs(observation)).to(torch.float64)
    z = torch.from_numpy(rng.standard_normal(size=tuple(observation.shape)))
    noise = (z.to(observation.device) * noise_level).to(dtype=observation.dtype)
    noisy_observation = observation + noise
    return (noisy_observation, noise_level.item()) if return_noise_level else noisy_observation


class SimulatedDataset(torch.utils.data.Dataset):
    """``(noisy_observation, x, filtbackproj)`` triples of an image dataset (reference :25-74):
    per-item seeds ``use_fixed_seeds_starting_from + idx`` unless a generator is passed."""

    def __init__(self, image_dataset, ray_trafo, white_noise_rel_stddev: float,
                 use_fixed_seeds_starting_from: Optional[int] = 1,
                 rng: Optional[np.random.Generator] = None, device: Optional[Any] = None):
        super().__init__()
        if rng is not None and use_fixed_seeds_starting_from is not None:
            raise AssertionError('must not use fixed seeds when passing a custom rng')
        self.image_dataset = image_dataset
        self.ray_trafo = ray_trafo
        self.white_noise_rel_stddev = white_noise_rel_stddev
        self.rng = rng
        self.use_fixed_seeds_starting_from = use_fixed_seeds_starting_from
        self.device = device

    def __len__(self):
        return len(self.image_dataset)

    def _generate_item(self, idx: int, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        rng = self.rng
        if rng is None:
            start = self.use_fixed_seeds_starting_from
            rng = np.random.default_rng(None if start is None else start + idx)
        x = x.to(device=self.device)
        noisy = simulate(x[None], ray_trafo=self.ray_trafo, white_noise_rel_stddev=self.white_noise_rel_stddev,
                         rng=rng)[0].to(device=self.device)
        filtbackproj = self.ray_trafo.fbp(noisy[None])[0].to(device=self.device)
        return noisy, x, filtbackproj

    def __iter__(self) -> Iterator[Tuple[Tensor, Tensor, Tensor]]:
        for idx, x in enumerate(self.image_dataset):
            yield self._generate_item(idx, x)

    def __getitem__(self, idx: int) -> Tuple[Tensor, Tensor, Tensor]:
        return self._generate_item(idx, self.image_dataset[idx])
