"""Size-independent properties at BASELINE.json's full sizes (batch 32, 256x256; -m gpu): the oracle cannot run these in seconds, so the
checks are algebraic -- batch-permutation equivariance and chunk independence of the eval forward (bit-exact), linearity of the data
gradient and of the weight gradient, and the sum rule of the BatchNorm statistics the convolution epilogue produces."""
import pytest
import torch

import model as model_mod
import sifnn_b200
import sifnn_oracle as O
from sifnn_b200 import ops
from conftest import load_ckpt, rel_err

pytestmark = pytest.mark.gpu


def test_eval_forward_batch_permutation_and_chunking_bit_exact():
    m = model_mod.ModelB_2(2)
    m.load_state_dict(load_ckpt("1009"))
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(9)
    lst, ndvi = torch.randn(40, 1, 64, 64, generator=g).cuda(), torch.randn(40, 1, 256, 256, generator=g).cuda()   # 40 > one chunk of 32
    with torch.inference_mode():
        y = m.forward_from_lowres(lst, ndvi)
        perm = torch.randperm(40, generator=g).cuda()
        yp = m.forward_from_lowres(lst[perm].contiguous(), ndvi[perm].contiguous())
        y1 = m.forward_from_lowres(lst[7:8].contiguous(), ndvi[7:8].contiguous())
    assert torch.equal(yp, y[perm])          # patches are independent in eval mode: any order, any batch split, same bits
    assert torch.equal(y1[0], y[7])


@pytest.mark.parametrize("shape", [(32, 16, 16, 256), (32, 64, 32, 128), (32, 128, 64, 64)])
def test_dgrad_and_wgrad_linearity_full_size(shape):
    B, Cin, Cout, HW = shape
    g = torch.Generator().manual_seed(10)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1).cuda()
    x = torch.randn(B, Cin, HW, HW, generator=g).cuda()
    d1, d2 = torch.randn(B, Cout, HW, HW, generator=g).cuda(), torch.randn(B, Cout, HW, HW, generator=g).cuda()
    a = 0.37
    lhs = ops.conv3x3_dgrad_tc(a * d1 + d2, w)
    rhs = a * ops.conv3x3_dgrad_tc(d1, w) + ops.conv3x3_dgrad_tc(d2, w)
    assert rel_err(lhs, rhs) < 1e-5
    wl = ops.conv3x3_wgrad_tc(x, a * d1 + d2)
    wr = a * ops.conv3x3_wgrad_tc(x, d1) + ops.conv3x3_wgrad_tc(x, d2)
    assert rel_err(wl, wr) < 2e-5
    # <dy, conv(x)> == <dgrad(dy), x> == <wgrad(x, dy), w>: the three kernels are adjoints of one bilinear form
    y = ops.conv3x3_fwd_tc(x, w)
    f = float((y.double() * d1.double()).sum())
    assert abs(float((ops.conv3x3_dgrad_tc(d1, w).double() * x.double()).sum()) - f) < 2e-5 * abs(f) + 1e-3
    assert abs(float((ops.conv3x3_wgrad_tc(x, d1).double() * w.double()).sum()) - f) < 2e-5 * abs(f) + 1e-3


def test_conv_statistics_epilogue_full_size():
    g = torch.Generator().manual_seed(11)
    x, w = torch.randn(32, 16, 256, 256, generator=g).cuda(), (torch.randn(16, 16, 3, 3, generator=g) * 0.2).cuda()
    stats = torch.zeros(32, dtype=torch.float64, device="cuda")
    y = ops.conv3x3_fwd_tc(x, w, None, None, None, stats)
    # per-thread fp32 partial sums over ~100 values, then fp64 atomics: error far below the fp32 rounding of the 2 M summands themselves
    yd = y.double()
    assert float((stats[:16] - yd.sum((0, 2, 3))).abs().max()) < 1e-6 * float(yd.abs().sum((0, 2, 3)).max())
    assert rel_err(stats[16:], (yd ** 2).sum((0, 2, 3))) < 1e-6
