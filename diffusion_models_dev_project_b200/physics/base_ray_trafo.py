"""Ray-transform object contract (boundary type of the hot path).

Mirrors the interface of the reference's ``BaseRayTrafo``
(reference src/physics/base_ray_trafo.py:13-201): shapes, method names and the
flat <-> 4-D adapter semantics (flat tensors are ``(numel, batch)``, i.e. one
*column* per sample) are the same, so samplers and scripts written against the
reference work unchanged with any subclass defined here.
"""
from abc import ABC, abstractmethod
from math import prod

from torch import Tensor, nn


class BaseRayTrafo(nn.Module, ABC):
    """Abstract ray transform ``A`` with adjoint ``A*`` and filtered back-projection.

    Attributes
    ----------
    im_shape : tuple of int
        ``(im_0, im_1)`` for 2-D geometries.
    obs_shape : tuple of int
        ``(angles, det_cols)`` for 2-D geometries.

    A subclass implements either the 4-D pair (:meth:`trafo`,
    :meth:`trafo_adjoint`) or the flat pair (:meth:`trafo_flat`,
    :meth:`trafo_adjoint_flat`) and obtains the other pair from the
    ``_*_via_*`` adapters.
    """

    def __init__(self, im_shape, obs_shape):
        super().__init__()
        self.im_shape = im_shape
        self.obs_shape = obs_shape

    # The reference probes ``hasattr(self, 'resize')`` to find out whether the
    # operator works on a resized image (reference base_ray_trafo.py:77,105,145,176).
    def _working_im_shape(self):
        return self.resize.shape if hasattr(self, 'resize') else self.im_shape

    @property
    def angles(self):
        """Projection angles in radians (``numpy.ndarray``)."""
        raise NotImplementedError

    # ------------------------------------------------------------- forward ---
    @abstractmethod
    def trafo(self, x: Tensor) -> Tensor:
        """``(batch, channels, im_0, im_1) -> (batch, channels, angles, det_cols)``."""
        raise NotImplementedError

    @abstractmethod
    def trafo_flat(self, x: Tensor) -> Tensor:
        """``(im_numel,)`` or ``(im_numel, batch)`` -> ``(obs_numel,)`` or ``(obs_numel, batch)``."""
        raise NotImplementedError

    def _trafo_via_trafo_flat(self, x: Tensor) -> Tensor:
        nb, nc = x.shape[:2]
        cols = x.reshape(nb * nc, prod(self._working_im_shape())).T
        return self.trafo_flat(cols).T.reshape(nb, nc, *self.obs_shape)

    def _trafo_flat_via_trafo(self, x: Tensor) -> Tensor:
        nb = x.shape[1]
        stack = x.T.reshape(1, nb, *self._working_im_shape())
        return self.trafo(stack).reshape(nb, prod(self.obs_shape)).T

    # ------------------------------------------------------------- adjoint ---
    @abstractmethod
    def trafo_adjoint(self, observation: Tensor) -> Tensor:
        """``(batch, channels, angles, det_cols) -> (batch, channels, im_0, im_1)``."""
        raise NotImplementedError

    @abstractmethod
    def trafo_adjoint_flat(self, observation: Tensor) -> Tensor:
        """``(obs_numel,)`` or ``(obs_numel, batch)`` -> ``(im_numel,)`` or ``(im_numel, batch)``."""
        raise NotImplementedError

    def _trafo_adjoint_via_trafo_adjoint_flat(self, observation: Tensor) -> Tensor:
        nb, nc = observation.shape[:2]
        cols = observation.reshape(nb * nc, prod(self.obs_shape)).T
        return self.trafo_adjoint_flat(cols).T.reshape(nb, nc, *self._working_im_shape())

    def _trafo_adjoint_flat_via_trafo_adjoint(self, observation: Tensor) -> Tensor:
        nb = observation.shape[1]
        stack = observation.T.reshape(1, nb, *self.obs_shape)
        return self.trafo_adjoint(stack).reshape(nb, prod(self._working_im_shape())).T

    # ----------------------------------------------------------------- fbp ---
    def fbp(self, observation: Tensor) -> Tensor:
        """Filtered back-projection, same shapes as :meth:`trafo_adjoint`."""
        raise NotImplementedError

    def forward(self, x: Tensor) -> Tensor:
        """``forward = trafo`` (reference base_ray_trafo.py:199-201)."""
        return self.trafo(x)
