"""PSNR as used for the parity gate (reference src/utils/metrics.py:4-11).
SSIM of the reference needs scikit-image, which is out of scope here."""
import numpy as np


def PSNR(reconstruction, ground_truth, data_range=None):
    gt = np.asarray(ground_truth)
    mse = np.mean((np.asarray(reconstruction) - gt) ** 2)
    if mse == 0.:
        return float('inf')
    if data_range is None:
        data_range = np.max(gt) - np.min(gt)
    return 20 * np.log10(data_range) - 10 * np.log10(mse)
