"""Image quality figures the reference's drivers print (src/utils/metrics.py): PSNR and SSIM.

PSNR is the parity gate of the full-reconstruction tests.  The reference takes SSIM from
scikit-image, which is not available offline; ``SSIM`` below evaluates the same published definition
(Wang et al. 2004 as parameterised by ``skimage.metrics.structural_similarity`` defaults: 7x7 uniform
window, K1 = 0.01, K2 = 0.03, sample covariance, mean over the window-valid interior).
"""
import numpy as np


def _range_of(img, data_range):
    return float(np.max(img) - np.min(img)) if data_range is None else data_range


def PSNR(reconstruction, ground_truth, data_range=None):
    """``20 log10(range) - 10 log10(mse)``; the range defaults to the ground truth's value span."""
    ref = np.asarray(ground_truth)
    err = np.asarray(reconstruction) - ref
    mse = np.mean(err * err)
    if mse == 0.:
        return float('inf')
    return 20 * np.log10(_range_of(ref, data_range)) - 10 * np.log10(mse)


def SSIM(reconstruction, ground_truth, data_range=None, win_size=7, k1=0.01, k2=0.03):
    from scipy.ndimage import uniform_filter
    ref = np.asarray(ground_truth, dtype=np.float64)
    img = np.asarray(reconstruction, dtype=np.float64)
    if img.shape != ref.shape or ref.ndim != 2 or min(ref.shape) < win_size:
        raise ValueError('SSIM expects two equally shaped 2-D images of at least %d pixels a side' % win_size)
    rng = _range_of(ref, data_range)
    n = float(win_size ** ref.ndim)
    box = lambda a: uniform_filter(a, size=win_size)          # noqa: E731
    mu_x, mu_y = box(img), box(ref)
    unbiased = n / (n - 1.0)
    var_x = unbiased * (box(img * img) - mu_x * mu_x)
    var_y = unbiased * (box(ref * ref) - mu_y * mu_y)
    cov = unbiased * (box(img * ref) - mu_x * mu_y)
    c1, c2 = (k1 * rng) ** 2, (k2 * rng) ** 2
    s = ((2 * mu_x * mu_y + c1) * (2 * cov + c2)) / ((mu_x ** 2 + mu_y ** 2 + c1) * (var_x + var_y + c2))
    pad = (win_size - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())
