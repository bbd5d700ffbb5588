"""The CPU restatement of the reference's PSNR / SSIM (oracle/quality_oracle.py; scikit-image is absent here): filter form against the
brute-force window-by-window form, and the algebraic properties of the definition."""
import numpy as np

import quality_oracle as Q


def test_ssim_filter_form_matches_bruteforce():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(20, 24)).astype(np.float32)
    y = (0.7 * x + 0.3 * rng.normal(size=x.shape)).astype(np.float32)
    r = float(x.max() - x.min())
    assert abs(Q.ssim_image(x, y, r) - Q.ssim_image_bruteforce(x, y, r)) < 1e-12


def test_metric_properties():
    rng = np.random.default_rng(1)
    t = rng.normal(size=(3, 1, 32, 32)).astype(np.float32)
    assert abs(Q.ssim_batch(t, t) - 1.0) < 1e-12                       # identical images
    p = t + 0.1 * rng.normal(size=t.shape).astype(np.float32)
    assert 0.0 < Q.ssim_batch(p, t) < 1.0
    r = float(t.max() - t.min())
    mse = np.mean((p.astype(np.float64) - t) ** 2, axis=(1, 2, 3))
    assert abs(Q.psnr_batch(p, t) - np.mean(10 * np.log10(r * r / mse))) < 1e-9
    assert Q.psnr_batch(t + 0.01, t) > Q.psnr_batch(t + 0.1, t)        # smaller error, higher PSNR
