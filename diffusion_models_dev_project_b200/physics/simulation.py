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
    """Lazily simulated measurements of an image dataset: item ``k`` is the triple
    ``(y_k, x_k, fbp(y_k))`` with ``y_k = simulate(x_k)`` (contract of the reference class,
    src/physics/simulation.py:25-74, which the drivers iterate and index).

    Noise is reproducible per item: unless a generator is shared across items (``rng``), item ``k`` draws from
    ``numpy.random.default_rng(use_fixed_seeds_starting_from + k)`` (``None`` = fresh entropy); passing both a
    generator and a seed base is refused like in the reference."""

    def __init__(self, image_dataset, ray_trafo, white_noise_rel_stddev: float,
                 use_fixed_seeds_starting_from: Optional[int] = 1,
                 rng: Optional[np.random.Generator] = None, device: Optional[Any] = None):
        super().__init__()
        assert rng is None or use_fixed_seeds_starting_from is None, \
            'must not use fixed seeds when passing a custom rng'
        self.image_dataset, self.ray_trafo, self.device = image_dataset, ray_trafo, device
        self.white_noise_rel_stddev = white_noise_rel_stddev
        self.use_fixed_seeds_starting_from, self.rng = use_fixed_seeds_starting_from, rng

    def _rng_for(self, k: int) -> np.random.Generator:
        if self.rng is not None:
            return self.rng
        base = self.use_fixed_seeds_starting_from
        return np.random.default_rng(base + k if base is not None else None)

    def _triple(self, k: int, image: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        image = image.to(device=self.device)
        y = simulate(image.unsqueeze(0), self.ray_trafo, self.white_noise_rel_stddev, rng=self._rng_for(k))
        y = y.to(device=self.device)
        return y[0], image, self.ray_trafo.fbp(y)[0].to(device=self.device)

    def __len__(self):
        return len(self.image_dataset)

    def __getitem__(self, k: int) -> Tuple[Tensor, Tensor, Tensor]:
        return self._triple(k, self.image_dataset[k])

    def __iter__(self) -> Iterator[Tuple[Tensor, Tensor, Tensor]]:
        return (self._triple(k, image) for k, image in enumerate(self.image_dataset))
