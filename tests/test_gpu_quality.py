"""On-device PSNR / SSIM (csrc/quality.cu) against the CPU restatement of the reference's skimage calls (-m gpu)."""
import numpy as np
import pytest
import torch

from sifnn_b200 import ops
import quality_oracle as Q
import sifnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 256, 256), (3, 64, 64), (1, 40, 72), (4, 7, 9)])
def test_psnr_ssim_matches_oracle(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(5)
    t = torch.randn(B, 1, H, W, generator=g)
    p = 0.8 * t + 0.2 * torch.randn(B, 1, H, W, generator=g)
    got = ops.psnr_ssim(p.cuda(), t.cuda()).cpu().numpy()
    want = np.array([Q.psnr_batch(p.numpy(), t.numpy()), Q.ssim_batch(p.numpy(), t.numpy())])
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6), (got, want)


def test_psnr_ssim_on_smooth_images_and_identity():
    lst, up, ndvi = O.smooth_batch(2)
    t = up
    p = up + 0.05 * ndvi
    got = ops.psnr_ssim(p.cuda(), t.cuda()).cpu().numpy()
    want = np.array([Q.psnr_batch(p.numpy(), t.numpy()), Q.ssim_batch(p.numpy(), t.numpy())])
    assert np.allclose(got, want, rtol=1e-5), (got, want)
    same = ops.psnr_ssim(t.cuda() + 1e-3, t.cuda()).cpu().numpy()
    assert same[1] > 0.999
