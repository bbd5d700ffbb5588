"""CPU restatement of the reference's train-time quality metrics -- TEST INFRASTRUCTURE ONLY (imported by tests/ and nowhere else).

The reference calls scikit-image (utils.py:548-578):
    psnr_skimage:  mean_i peak_signal_noise_ratio(targets[i,0], predictions[i,0], data_range=targets.max() - targets.min())
    ssim_skimage:  mean_i structural_similarity(targets[i,0], predictions[i,0], data_range=targets.max() - targets.min())
scikit-image (environment.yml pins scikit-image 0.22) is NOT installed in this container and is not vendored by the reference, so
this file restates its published algorithm with numpy / scipy:
  * PSNR = 10 log10(data_range^2 / mean((a - b)^2));
  * SSIM (Wang et al. 2004, skimage defaults): win_size 7, uniform filter, use_sample_covariance=True (cov_norm = 49/48),
    K1 = 0.01, K2 = 0.03, S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), mean of S over the map cropped by
    (win_size - 1) / 2 = 3 pixels on every side.
PARITY UNPINNED against scikit-image itself (absent); pinned only to a brute-force evaluation of the same formula and to the
algebraic properties checked in tests/test_quality_cpu.py.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def psnr_batch(predictions: np.ndarray, targets: np.ndarray) -> float:
    """utils.py:548-552."""
    r = float(targets.max() - targets.min())
    vals = []
    for i in range(targets.shape[0]):
        a, b = targets[i, 0].astype(np.float64), predictions[i, 0].astype(np.float64)
        vals.append(10.0 * np.log10(r * r / np.mean((a - b) ** 2)))
    return float(np.mean(vals))


def ssim_image(x: np.ndarray, y: np.ndarray, data_range: float, win: int = 7) -> float:
    x, y = x.astype(np.float64), y.astype(np.float64)
    npx = win * win
    cov_norm = npx / (npx - 1.0)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())


def ssim_batch(predictions: np.ndarray, targets: np.ndarray) -> float:
    """utils.py:554-578."""
    r = float(targets.max() - targets.min())
    return float(np.mean([ssim_image(targets[i, 0], predictions[i, 0], r) for i in range(targets.shape[0])]))


def ssim_image_bruteforce(x: np.ndarray, y: np.ndarray, data_range: float, win: int = 7) -> float:
    """Window-by-window evaluation of the same definition (small images only)."""
    x, y = x.astype(np.float64), y.astype(np.float64)
    h, w = x.shape
    pad = (win - 1) // 2
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    vals = []
    for i in range(pad, h - pad):
        for j in range(pad, w - pad):
            a, b = x[i - pad:i + pad + 1, j - pad:j + pad + 1], y[i - pad:i + pad + 1, j - pad:j + pad + 1]
            ux, uy = a.mean(), b.mean()
            vx, vy = a.var(ddof=1), b.var(ddof=1)
            vxy = ((a - ux) * (b - uy)).sum() / (win * win - 1)
            vals.append(((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2)))
    return float(np.mean(vals))
