"""tcgen05 / TMEM implicit-GEMM convolution (3-term TF32 split) against torch fp64 (-m gpu)."""
import pytest
import torch
import torch.nn.functional as F

import sifnn_b200
from sifnn_b200 import ops
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def ref_conv(x, w, b=None):
    return F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="replicate"), w.double(), None if b is None else b.double())


SHAPES = [(2, 16, 16, 8, 128), (1, 32, 16, 6, 256), (2, 64, 32, 5, 128), (1, 8, 64, 4, 128), (3, 16, 32, 16, 128), (1, 128, 64, 3, 128),
          (2, 32, 32, 64, 64), (1, 128, 64, 10, 64), (3, 64, 32, 7, 64), (2, 16, 16, 9, 64), (40, 32, 64, 64, 64),
          (4, 64, 64, 32, 32), (2, 16, 32, 5, 96), (3, 32, 16, 8, 16), (2, 8, 16, 6, 20)]


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_fwd_plain(shape):
    B, Cin, Cout, H, W = shape
    x, w = rnd(B, Cin, H, W, seed=1), rnd(Cout, Cin, 3, 3, seed=2, scale=0.2)
    y = ops.conv3x3_fwd_tc(x.cuda(), w.cuda())
    assert rel_err(y, ref_conv(x, w)) < TOL


@pytest.mark.parametrize("shape", [(2, 16, 16, 8, 128), (1, 32, 64, 6, 256), (2, 32, 32, 64, 64), (2, 64, 64, 6, 64), (3, 64, 64, 32, 32)])
def test_tc_fwd_affine_stats_bias(shape):
    B, Cin, Cout, H, W = shape
    x, w = rnd(B, Cin, H, W, seed=4), rnd(Cout, Cin, 3, 3, seed=5, scale=0.2)
    sc, sh, bias = 1 + 0.3 * rnd(Cin, seed=6), 0.2 * rnd(Cin, seed=7), rnd(Cout, seed=8)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = ops.conv3x3_fwd_tc(x.cuda(), w.cuda(), bias.cuda(), sc.cuda(), sh.cuda(), stats)
    a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    ref = ref_conv(a, w, bias)
    assert rel_err(y, ref) < TOL
    assert rel_err(stats[:Cout], ref.sum((0, 2, 3))) < 1e-4
    assert rel_err(stats[Cout:], (ref * ref).sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("shape", [(2, 16, 16, 8, 128), (1, 16, 32, 6, 256), (2, 64, 32, 5, 128), (1, 32, 64, 4, 128),
                                   (2, 128, 64, 64, 64), (2, 32, 32, 9, 64), (1, 128, 64, 5, 64), (2, 64, 32, 64, 64),
                                   (4, 64, 64, 32, 32), (2, 32, 16, 7, 96), (2, 16, 16, 8, 16)])
def test_tc_dgrad(shape):
    B, Cin, Cout, H, W = shape
    w, dy = rnd(Cout, Cin, 3, 3, seed=8, scale=0.2), rnd(B, Cout, H, W, seed=9)
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    (ref_conv(x, w) * dy.double()).sum().backward()
    dx = ops.conv3x3_dgrad_tc(dy.cuda(), w.cuda())
    assert rel_err(dx, x.grad) < TOL
    base = rnd(B, Cin, H, W, seed=10)
    acc = ops.conv3x3_dgrad_tc(dy.cuda(), w.cuda(), base.cuda().clone(), accumulate=True)
    assert rel_err(acc, x.grad + base.double()) < TOL


def test_tc_matches_simt_closely():
    """Same inputs through the SIMT fp32 kernel and the 3xTF32 tensor-core kernel."""
    x, w = rnd(2, 32, 32, 256, seed=11).cuda(), rnd(16, 32, 3, 3, seed=12, scale=0.2).cuda()
    a, b = ops.conv3x3_fwd(x, w), ops.conv3x3_fwd_tc(x, w)
    assert rel_err(b, a) < 5e-6


def ref_wgrad(x, dy):
    w = torch.zeros(dy.shape[1], x.shape[1], 3, 3, dtype=torch.float64, requires_grad=True)
    (ref_conv(x, w) * dy.double()).sum().backward()
    return w.grad


WG_SHAPES = [(2, 16, 16, 8, 16), (1, 16, 16, 5, 32), (2, 32, 16, 8, 64), (2, 16, 32, 7, 48), (1, 32, 32, 12, 128), (2, 64, 32, 9, 64),
             (1, 32, 64, 6, 32), (2, 128, 64, 5, 64), (3, 64, 64, 32, 32), (1, 16, 64, 3, 16), (2, 32, 16, 256, 256), (40, 16, 16, 64, 64)]


@pytest.mark.parametrize("shape", WG_SHAPES)
def test_tc_wgrad(shape):
    B, Cin, Cout, H, W = shape
    x, dy = rnd(B, Cin, H, W, seed=21), rnd(B, Cout, H, W, seed=22)
    dw = ops.conv3x3_wgrad_tc(x.cuda(), dy.cuda())
    assert rel_err(dw, ref_wgrad(x, dy)) < TOL
    assert torch.equal(dw, ops.conv3x3_wgrad_tc(x.cuda(), dy.cuda()))  # deterministic reduction


@pytest.mark.parametrize("shape", [(2, 16, 16, 8, 32), (2, 32, 16, 6, 64), (1, 64, 64, 8, 32), (2, 128, 64, 4, 64)])
def test_tc_wgrad_affine(shape):
    B, Cin, Cout, H, W = shape
    x, dy = rnd(B, Cin, H, W, seed=23), rnd(B, Cout, H, W, seed=24)
    sc, sh = 1 + 0.3 * rnd(Cin, seed=25), 0.2 * rnd(Cin, seed=26)
    a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    dw = ops.conv3x3_wgrad_tc(x.cuda(), dy.cuda(), sc.cuda(), sh.cuda())
    assert rel_err(dw, ref_wgrad(a, dy)) < TOL


def test_tc_wgrad_same_sign_sum_is_unbiased():
    """All-positive operands: every product has the same sign, the worst case for the truncating tensor-core accumulator
    (a long chain would drift by ~6e-8 per step); the rotating accumulator sets keep it inside the parity tolerance."""
    B, Cin, Cout, H, W = 32, 16, 16, 256, 256
    x, dy = rnd(B, Cin, H, W, seed=27).abs(), rnd(B, Cout, H, W, seed=28).abs()
    dw = ops.conv3x3_wgrad_tc(x.cuda(), dy.cuda())
    ref = ops.conv3x3_wgrad(x.cuda(), dy.cuda())
    assert rel_err(dw, ref) < 1e-4 / 2


def test_tc_bf16x3_optin_mode_accuracy():
    """Opt-in 3-term BF16 split of the kx-folded kernel (SIFNN_TC_BF16=1, separate process): K = 16 per MMA, stated tolerance 2e-5 per layer
    (measured 5e-6) against 2e-5 / measured 4e-7 for the default TF32 split."""
    import os, subprocess, sys
    code = (
        "import sys, torch, torch.nn.functional as F\n"
        "sys.path.insert(0, %r)\n"
        "import sifnn_b200\n"
        "from sifnn_b200 import ops\n"
        "g = torch.Generator().manual_seed(3)\n"
        "for ci, co, hw in [(16, 16, 128), (32, 16, 256)]:\n"
        "    x = torch.randn(2, ci, 12, hw, generator=g); w = torch.randn(co, ci, 3, 3, generator=g) * 0.2; dy = torch.randn(2, co, 12, hw, generator=g)\n"
        "    ref = F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode='replicate'), w.double())\n"
        "    y = ops.conv3x3_fwd_tc(x.cuda(), w.cuda()).cpu().double()\n"
        "    e = float((y - ref).abs().max() / ref.abs().max())\n"
        "    xx = torch.zeros_like(x, dtype=torch.float64, requires_grad=True)\n"
        "    (F.conv2d(F.pad(xx, (1, 1, 1, 1), mode='replicate'), w.double()) * dy.double()).sum().backward()\n"
        "    dx = ops.conv3x3_dgrad_tc(dy.cuda(), w.cuda()).cpu().double()\n"
        "    e2 = float((dx - xx.grad).abs().max() / xx.grad.abs().max())\n"
        "    assert 1e-7 < e < 2e-5 and e2 < 2e-5, (ci, co, hw, e, e2)\n"
        "print('ok')\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, SIFNN_TC_BF16="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr
