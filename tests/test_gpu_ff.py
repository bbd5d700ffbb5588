"""Full-fold tcgen05 convolution (csrc/conv3x3_ff.cu): forward with replicate padding (model.py:135, nn.Conv2d(padding_mode='replicate'))
and its complete autograd data gradient, against torch fp64 on the CPU (-m gpu).

Tolerances: max|a - b| / max|b|.  TF32 and FP16 splits (22 significant bits): ~2e-6 measured -> 1e-5.  BF16 split (16 bits): ~5e-6 measured -> 3e-5.
All far inside the 1e-4 bar."""
import pytest
import torch
import torch.nn.functional as F

import sifnn_b200
from sifnn_b200 import ops, _lib
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = {0: 3e-5, 1: 1e-5, 2: 1e-5}      # forward, keyed by the split kind: 0 BF16, 1 TF32, 2 FP16 (forward only)
TOL_DG = {0: 3e-5, 1: 1e-5, 2: 3e-5}   # data gradient: kind 2 runs the BF16 split there


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def ref_conv(x, w):
    return F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="replicate"), w.double())


@pytest.fixture(params=[0, 1, 2], ids=["bf16x3", "tf32x3", "fp16x3"])
def prec(request):
    lib = _lib.load()
    lib.sifnn_conv3x3_ff_config(request.param, 0)
    yield request.param
    lib.sifnn_conv3x3_ff_config(2, 0)   # library default: FP16 forward, BF16 data gradient


# (B, Cin, Cout, H, W): every width class (4 / 2 / 1 segments per tile, two tiles per row), 1 and 2 output groups per CTA, blockIdx.y > 1,
# ragged row partitions, single-row images
SHAPES = [(2, 16, 16, 8, 128), (1, 32, 16, 6, 256), (2, 64, 32, 5, 128), (3, 16, 32, 16, 128), (2, 32, 32, 64, 64), (3, 64, 32, 7, 64),
          (2, 16, 16, 9, 64), (4, 64, 64, 32, 32), (5, 32, 64, 3, 32), (1, 16, 16, 1, 128), (1, 16, 16, 2, 256), (2, 16, 16, 1, 32),
          (2, 64, 64, 64, 64), (7, 16, 16, 5, 256), (1, 64, 128, 11, 64), (3, 16, 16, 256, 256), (1, 32, 32, 33, 256)]


@pytest.mark.parametrize("shape", SHAPES)
def test_ff_fwd_plain(shape, prec):
    B, Cin, Cout, H, W = shape
    x, w = rnd(B, Cin, H, W, seed=1), rnd(Cout, Cin, 3, 3, seed=2, scale=0.2)
    y = ops.conv3x3_fwd_ff(x.cuda(), w.cuda())
    assert rel_err(y, ref_conv(x, w)) < TOL[prec]


@pytest.mark.parametrize("shape", [(2, 16, 16, 8, 128), (1, 32, 16, 6, 256), (2, 32, 32, 64, 64), (2, 64, 64, 6, 64), (3, 64, 64, 32, 32), (2, 16, 32, 40, 128)])
def test_ff_fwd_affine_stats(shape, prec):
    B, Cin, Cout, H, W = shape
    x, w = rnd(B, Cin, H, W, seed=4), rnd(Cout, Cin, 3, 3, seed=5, scale=0.2)
    sc, sh = 1 + 0.3 * rnd(Cin, seed=6), 0.2 * rnd(Cin, seed=7)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = ops.conv3x3_fwd_ff(x.cuda(), w.cuda(), sc.cuda(), sh.cuda(), stats)
    a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    ref = ref_conv(a, w)
    assert rel_err(y, ref) < TOL[prec]
    assert rel_err(stats[:Cout], ref.sum((0, 2, 3))) < 1e-4
    assert rel_err(stats[Cout:], (ref * ref).sum((0, 2, 3))) < 3e-5


DG_SHAPES = [(2, 16, 16, 8, 128), (1, 16, 32, 6, 256), (2, 64, 32, 5, 128), (1, 32, 64, 4, 128), (2, 128, 64, 64, 64), (2, 32, 32, 9, 64),
             (1, 128, 64, 5, 64), (4, 64, 64, 32, 32), (2, 32, 16, 7, 256), (1, 16, 16, 1, 64), (3, 16, 16, 2, 32), (2, 32, 16, 150, 256)]


@pytest.mark.parametrize("shape", DG_SHAPES)
def test_ff_dgrad_complete(shape, prec):
    """dy -> dx including the adjoint of the replicate padding (rows, columns, corners), plain and accumulating."""
    B, Cin, Cout, H, W = shape
    w, dy = rnd(Cout, Cin, 3, 3, seed=8, scale=0.2), rnd(B, Cout, H, W, seed=9)
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    (ref_conv(x, w) * dy.double()).sum().backward()
    dx = ops.conv3x3_dgrad_ff(dy.cuda(), w.cuda())
    assert rel_err(dx, x.grad) < TOL_DG[prec]
    base = rnd(B, Cin, H, W, seed=10)
    acc = ops.conv3x3_dgrad_ff(dy.cuda(), w.cuda(), base.cuda().clone(), accumulate=True)
    assert rel_err(acc, x.grad + base.double()) < TOL_DG[prec]


@pytest.mark.parametrize("max_ctas", [1, 2, 3, 7])
def test_ff_long_strips_and_image_crossings(max_ctas):
    """Few CTAs -> each walks many rows: strips longer than the 64-row carry window, ranges that cross image boundaries."""
    lib = _lib.load()
    try:
        for tf32 in (0, 1, 2):
            lib.sifnn_conv3x3_ff_config(tf32, max_ctas)
            for (B, Cin, Cout, H, W) in [(3, 16, 16, 100, 256), (5, 16, 32, 20, 64), (3, 32, 16, 37, 128), (9, 16, 16, 8, 32)]:
                x, w = rnd(B, Cin, H, W, seed=21), rnd(Cout, Cin, 3, 3, seed=22, scale=0.2)
                sc, sh = 1 + 0.3 * rnd(Cin, seed=23), 0.2 * rnd(Cin, seed=24)
                stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
                y = ops.conv3x3_fwd_ff(x.cuda(), w.cuda(), sc.cuda(), sh.cuda(), stats)
                a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
                ref = ref_conv(a, w)
                assert rel_err(y, ref) < TOL[tf32], (tf32, B, Cin, Cout, H, W)
                assert rel_err(stats[:Cout], ref.sum((0, 2, 3))) < 1e-4
                dy = rnd(B, Cout, H, W, seed=25)
                xx = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
                (ref_conv(xx, w) * dy.double()).sum().backward()
                dx = ops.conv3x3_dgrad_ff(dy.cuda(), w.cuda())
                assert rel_err(dx, xx.grad) < TOL_DG[tf32], (tf32, B, Cin, Cout, H, W)
    finally:
        lib.sifnn_conv3x3_ff_config(2, 0)   # library default: FP16 forward, BF16 data gradient


def test_ff_matches_simt_kernel():
    """Same inputs through the strict-fp32 SIMT kernel and the full-fold kernel (FP16 split, the training default)."""
    lib = _lib.load()
    lib.sifnn_conv3x3_ff_config(2, 0)
    try:
        x, w = rnd(2, 32, 32, 256, seed=11).cuda(), rnd(16, 32, 3, 3, seed=12, scale=0.2).cuda()
        a, b = ops.conv3x3_fwd(x, w), ops.conv3x3_fwd_ff(x, w)
        assert rel_err(b, a) < 5e-6
    finally:
        lib.sifnn_conv3x3_ff_config(2, 0)   # library default: FP16 forward, BF16 data gradient


def test_ff_rejects_unsupported():
    x, w = rnd(1, 8, 8, 128).cuda(), rnd(16, 8, 3, 3).cuda()
    with pytest.raises(sifnn_b200.SifnnError):
        ops.conv3x3_fwd_ff(x, w)
    x, w = rnd(1, 16, 8, 96).cuda(), rnd(16, 16, 3, 3).cuda()
    with pytest.raises(sifnn_b200.SifnnError):
        ops.conv3x3_fwd_ff(x, w)


@pytest.mark.parametrize("kind", [0, 2], ids=["bf16x3", "fp16x3"])
@pytest.mark.parametrize("shape", [(1, 128, 64, 11, 64), (2, 128, 64, 64, 64), (2, 128, 32, 8, 32), (1, 128, 16, 5, 128), (3, 96, 64, 6, 64)])
def test_ff_128_input_channels(shape, kind):
    """65..128 input channels (the decoder's first convolution, 128 -> 64 @ 64^2): one output group per CTA, 16-bit splits only; forward with the
    BatchNorm prologue and statistics, and the data gradient of a layer with that many OUTPUT channels (its K)."""
    lib = _lib.load()
    lib.sifnn_conv3x3_ff_config(kind, 0)
    try:
        B, Cin, Cout, H, W = shape
        assert lib.sifnn_conv3x3_ff_supported(Cin, Cout, H, W)
        x, w = rnd(B, Cin, H, W, seed=31), rnd(Cout, Cin, 3, 3, seed=32, scale=0.1)
        sc, sh = 1 + 0.3 * rnd(Cin, seed=33), 0.2 * rnd(Cin, seed=34)
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
        y = ops.conv3x3_fwd_ff(x.cuda(), w.cuda(), sc.cuda(), sh.cuda(), stats)
        a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
        ref = ref_conv(a, w)
        assert rel_err(y, ref) < TOL[kind]
        assert rel_err(stats[:Cout], ref.sum((0, 2, 3))) < 1e-4
        assert rel_err(ops.conv3x3_fwd_ff(x.cuda(), w.cuda()), ref_conv(x, w)) < TOL[kind]
        # data gradient with K = 128: dy has Cin channels here, dx Cout
        wt = rnd(Cin, Cout, 3, 3, seed=35, scale=0.1)           # a layer Cout -> Cin
        dy = rnd(B, Cin, H, W, seed=36)
        xx = torch.zeros(B, Cout, H, W, dtype=torch.float64, requires_grad=True)
        (ref_conv(xx, wt) * dy.double()).sum().backward()
        assert rel_err(ops.conv3x3_dgrad_ff(dy.cuda(), wt.cuda()), xx.grad) < TOL_DG[kind]
    finally:
        lib.sifnn_conv3x3_ff_config(2, 0)


def test_ff_tf32_keeps_64_input_channels():
    lib = _lib.load()
    lib.sifnn_conv3x3_ff_config(1, 0)
    try:
        assert not lib.sifnn_conv3x3_ff_supported(128, 64, 64, 64) and lib.sifnn_conv3x3_ff_supported(64, 64, 64, 64)
    finally:
        lib.sifnn_conv3x3_ff_config(2, 0)
