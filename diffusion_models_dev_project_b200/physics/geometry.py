"""Host-side 2-D parallel-beam geometry (no ODL).

Restates what the reference obtains from ``odl.uniform_discr`` and
``odl.tomo.parallel_beam_geometry`` in ``SimpleTrafo.__init__``
(reference src/physics/trafo.py:18-27; rule written out in SURVEY.md §3.5 and
Appendix A):

* image domain ``[(-n)//2, n//2]^2`` with ``n x n`` unit-ish cells -- note the
  operator precedence ``-n//2 == (-n)//2``: 256 -> [-128,128], 501 -> [-251,250];
* angles: midpoints of a uniform partition of ``[0, pi)``;
* detector: ``rho`` = largest corner distance, ``n_det = 2*ceil(rho/cell)+1``
  cells on ``[-rho, rho]``.
"""
from dataclasses import dataclass
from math import ceil, pi, sqrt

import numpy as np


@dataclass(frozen=True)
class ParallelBeamGeometry2D:
    n0: int
    n1: int
    x_min: float
    y_min: float
    dx: float
    angles: np.ndarray      # [n_angles] float64, radians
    n_det: int
    s_min: float
    ds: float

    @property
    def n_angles(self):
        return int(self.angles.shape[0])

    @property
    def im_shape(self):
        return (self.n0, self.n1)

    @property
    def obs_shape(self):
        return (self.n_angles, self.n_det)

    @property
    def dphi(self):
        """Angle cell size (uniform partition of [0, pi))."""
        return pi / self.n_angles

    @property
    def range_weight(self):
        """ODL cell-volume ratio c_w = dphi*ds/(dx*dx)."""
        return self.dphi * self.ds / (self.dx * self.dx)

    @staticmethod
    def from_im_shape(im_shape, num_angles):
        """Geometry of ``SimpleTrafo(im_shape, num_angles)``."""
        n0, n1 = int(im_shape[0]), int(im_shape[1])
        lo = np.array([(-n0) // 2, (-n1) // 2], dtype=np.float64)
        hi = np.array([n0 // 2, n1 // 2], dtype=np.float64)
        cell = (hi - lo) / np.array([n0, n1], dtype=np.float64)
        if abs(cell[0] - cell[1]) > 1e-12:
            raise ValueError('only square pixels are supported (got cell sides %r)' % (cell,))
        corners = [(x, y) for x in (lo[0], hi[0]) for y in (lo[1], hi[1])]
        rho = max(sqrt(x * x + y * y) for x, y in corners)
        n_det = 2 * int(ceil(rho / float(cell.min()))) + 1
        angles = (np.arange(num_angles, dtype=np.float64) + 0.5) * (pi / num_angles)
        return ParallelBeamGeometry2D(
            n0=n0, n1=n1, x_min=float(lo[0]), y_min=float(lo[1]), dx=float(cell[0]),
            angles=angles, n_det=n_det, s_min=-rho, ds=2.0 * rho / n_det)
