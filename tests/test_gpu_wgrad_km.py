"""Weight gradient with 16-bit K-major operands (csrc/wgrad_km.cu; autograd of model.py:135 w.r.t. the convolution weight) against torch fp64 on
the CPU (-m gpu).  Both operands are split in BF16 (hi + lo, 16 significant bits; kind::f16 wants one format for both operands): measured error ~4e-6 of max|dW| -> 2e-5."""
import pytest
import torch
import torch.nn.functional as F

import sifnn_b200
from sifnn_b200 import ops, _lib
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def ref_wgrad(x, dy):
    w = torch.zeros(dy.shape[1], x.shape[1], 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="replicate"), w) * dy.double()).sum().backward()
    return w.grad


# (B, Cin, Cout, H, W): every tile shape (16/16: 8 rows, 32/16 and 16/32: 4 rows, 32/32: 2 rows), channel blocks over blockIdx.y, several column tiles,
# fewer tiles than accumulator sets, more tiles than CTAs
SHAPES = [(2, 16, 16, 8, 32), (1, 16, 16, 16, 64), (2, 32, 16, 8, 64), (2, 16, 32, 8, 32), (1, 32, 32, 12, 128), (2, 64, 32, 6, 64),
          (1, 32, 64, 6, 32), (2, 128, 64, 4, 64), (3, 64, 64, 32, 32), (2, 32, 16, 256, 256), (40, 16, 16, 64, 64), (1, 16, 16, 8, 32)]


@pytest.mark.parametrize("shape", SHAPES)
def test_km_wgrad(shape):
    B, Cin, Cout, H, W = shape
    x, dy = rnd(B, Cin, H, W, seed=21), rnd(B, Cout, H, W, seed=22)
    dw = ops.conv3x3_wgrad_km(x.cuda(), dy.cuda())
    assert rel_err(dw, ref_wgrad(x, dy)) < TOL
    assert torch.equal(dw, ops.conv3x3_wgrad_km(x.cuda(), dy.cuda()))  # deterministic reduction


@pytest.mark.parametrize("shape", [(2, 16, 16, 8, 32), (2, 32, 16, 8, 64), (1, 64, 64, 8, 32), (2, 128, 64, 4, 64)])
def test_km_wgrad_affine(shape):
    B, Cin, Cout, H, W = shape
    x, dy = rnd(B, Cin, H, W, seed=23), rnd(B, Cout, H, W, seed=24)
    sc, sh = 1 + 0.3 * rnd(Cin, seed=25), 0.2 * rnd(Cin, seed=26)
    a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    dw = ops.conv3x3_wgrad_km(x.cuda(), dy.cuda(), sc.cuda(), sh.cuda())
    assert rel_err(dw, ref_wgrad(a, dy)) < TOL


def test_km_wgrad_tiny_gradients_keep_their_precision():
    """dy of magnitude 1e-6 (what a mean-reduced loss over millions of pixels produces): the BF16 split keeps the fp32 exponent range, so the
    relative error is the same as for O(1) gradients."""
    B, Cin, Cout, H, W = 2, 16, 16, 16, 64
    x, dy = rnd(B, Cin, H, W, seed=31), rnd(B, Cout, H, W, seed=32) * 1e-6
    dw = ops.conv3x3_wgrad_km(x.cuda(), dy.cuda())
    assert rel_err(dw, ref_wgrad(x, dy)) < TOL


def test_km_wgrad_same_sign_sum_is_unbiased():
    """All-positive operands at the benchmark's size: every product has the same sign, the worst case for the truncating tensor-core accumulator;
    the rotating accumulator sets keep the drift inside the parity tolerance."""
    B, Cin, Cout, H, W = 32, 16, 16, 256, 256
    x, dy = rnd(B, Cin, H, W, seed=27).abs(), rnd(B, Cout, H, W, seed=28).abs()
    dw = ops.conv3x3_wgrad_km(x.cuda(), dy.cuda())
    ref = ops.conv3x3_wgrad(x.cuda(), dy.cuda())
    assert rel_err(dw, ref) < 1e-4 / 2


def test_km_wgrad_rejects_unsupported():
    with pytest.raises(sifnn_b200.SifnnError):
        ops.conv3x3_wgrad_km(rnd(1, 8, 8, 32).cuda(), rnd(1, 16, 8, 32).cuda())
    with pytest.raises(sifnn_b200.SifnnError):
        ops.conv3x3_wgrad_km(rnd(1, 16, 8, 48).cuda(), rnd(1, 16, 8, 48).cuda())


def test_km_two_step_form_equals_the_one_call_form():
    """sifnn_conv3x3_wgrad_km_partials + sifnn_wgrad_reduce (what the network plan uses, reduces batched per backward phase) == sifnn_conv3x3_wgrad_km."""
    import ctypes
    lib = _lib.load()
    for (B, Cin, Cout, H, W) in [(2, 16, 16, 8, 64), (2, 32, 16, 8, 64), (1, 64, 32, 4, 128)]:
        x, dy = rnd(B, Cin, H, W, seed=41).cuda(), rnd(B, Cout, H, W, seed=42).cuda()
        ref = ops.conv3x3_wgrad_km(x, dy)
        ws = torch.empty(lib.sifnn_conv3x3_wgrad_km_workspace(B, Cin, Cout, H, W), dtype=torch.uint8, device="cuda")
        dw = torch.full((Cout, Cin, 3, 3), float("nan"), device="cuda")
        slots = ctypes.c_int(0)
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("sifnn_conv3x3_wgrad_km_partials", x.data_ptr(), None, None, dy.data_ptr(), ws.data_ptr(), B, Cin, Cout, H, W, st, ctypes.addressof(slots))
        assert slots.value > 0
        _lib.call("sifnn_wgrad_reduce", ws.data_ptr(), dw.data_ptr(), Cout * Cin * 9, slots.value, st)
        assert torch.equal(dw, ref)
