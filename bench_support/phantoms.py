"""Synthetic disk-distributed ellipse phantoms (no ODL).

Follows the recipe of the reference's DiskDistributedEllipsesDataset
(src/dataset/ellipses.py:108-136 for the random parameters, :72-79 for the foreground
normalisation): up to `max_n_ellipse` ellipses with centres drawn in a disk of relative
radius `diameter`, values in [-0.4, 1], summed, non-zero pixels shifted to start at 0 and
the image scaled to [0, 1].  The rasterisation is a plain inside-the-ellipse test on pixel
centres of the domain [-1, 1]^2.
"""
import numpy as np


def disk_ellipses(n_images, im_size=256, seed=1, diameter=0.4745, max_n_ellipse=140):
    rng = np.random.RandomState(seed)
    c = (np.arange(im_size) + 0.5) / im_size * 2.0 - 1.0
    X, Y = np.meshgrid(c, c, indexing='ij')
    out = np.zeros((n_images, 1, im_size, im_size), dtype=np.float32)
    for i in range(n_images):
        v = rng.uniform(-0.4, 1.0, (max_n_ellipse,))
        a1 = .2 * diameter * rng.exponential(1., (max_n_ellipse,))
        a2 = .2 * diameter * rng.exponential(1., (max_n_ellipse,))
        c_r = rng.triangular(0., diameter, diameter, size=(max_n_ellipse,))
        c_a = rng.uniform(0., 2 * np.pi, (max_n_ellipse,))
        x0, y0 = np.cos(c_a) * c_r, np.sin(c_a) * c_r
        rot = rng.uniform(0., 2 * np.pi, (max_n_ellipse,))
        n_ell = min(rng.poisson(max_n_ellipse), max_n_ellipse)
        img = np.zeros((im_size, im_size), dtype=np.float64)
        for k in range(n_ell):
            cr, sr = np.cos(rot[k]), np.sin(rot[k])
            xr = (X - x0[k]) * cr + (Y - y0[k]) * sr
            yr = -(X - x0[k]) * sr + (Y - y0[k]) * cr
            img[(xr / max(a1[k], 1e-6)) ** 2 + (yr / max(a2[k], 1e-6)) ** 2 <= 1.0] += v[k]
        nz = img != 0.
        if nz.any():
            img[nz] -= img.min()
            if img.max() > 0:
                img /= img.max()
        out[i, 0] = img.astype(np.float32)
    return out
