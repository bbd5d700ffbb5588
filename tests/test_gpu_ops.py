"""Per-kernel parity (-m gpu): every C-ABI op against the same op written with plain torch
on the CPU (fp64, so the tolerance measures OUR rounding only).  Tolerance: 1e-4 relative to
the tensor's max-abs (north_star), in practice 1e-6."""
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


CONV_SHAPES = [  # (B, Cin, Cout, H, W)
    (2, 2, 16, 64, 64), (1, 16, 16, 256, 256), (2, 16, 32, 32, 32), (2, 32, 32, 24, 40), (1, 32, 64, 16, 16),
    (1, 64, 64, 32, 32), (1, 128, 64, 16, 16), (2, 64, 32, 8, 8), (2, 32, 16, 40, 72), (3, 16, 1, 64, 64), (1, 16, 1, 8, 8),
    (1, 5, 24, 19, 37),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_fwd_plain(shape):
    B, Cin, Cout, H, W = shape
    x, w = rnd(B, Cin, H, W, seed=1), rnd(Cout, Cin, 3, 3, seed=2, scale=0.2)
    b = rnd(Cout, seed=3) if Cout == 1 else None
    y = ops.conv3x3_fwd(x.cuda(), w.cuda(), None if b is None else b.cuda())
    assert rel_err(y, ref_conv(x, w, b)) < TOL


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (1, 32, 16, 32, 96), (2, 64, 32, 16, 16), (2, 16, 1, 32, 32)])
def test_conv_fwd_affine_and_stats(shape):
    B, Cin, Cout, H, W = shape
    x, w = rnd(B, Cin, H, W, seed=4), rnd(Cout, Cin, 3, 3, seed=5, scale=0.2)
    sc, sh = 1 + 0.3 * rnd(Cin, seed=6), 0.2 * rnd(Cin, seed=7)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = ops.conv3x3_fwd(x.cuda(), w.cuda(), None, sc.cuda(), sh.cuda(), stats)
    a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    ref = ref_conv(a, w)
    assert rel_err(y, ref) < TOL
    assert rel_err(stats[:Cout], ref.sum((0, 2, 3))) < 1e-5 or ref.sum((0, 2, 3)).abs().max() < 1e-3
    assert rel_err(stats[Cout:], (ref * ref).sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (1, 16, 32, 32, 32), (2, 128, 64, 16, 16), (2, 64, 64, 8, 8), (2, 16, 1, 32, 64), (1, 32, 16, 24, 40)])
def test_conv_dgrad(shape):
    B, Cin, Cout, H, W = shape
    w, dy = rnd(Cout, Cin, 3, 3, seed=8, scale=0.2), rnd(B, Cout, H, W, seed=9)
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    (ref_conv(x, w) * dy.double()).sum().backward()
    dx = ops.conv3x3_dgrad(dy.cuda(), w.cuda())
    assert rel_err(dx, x.grad) < TOL
    base = rnd(B, Cin, H, W, seed=10)
    acc = ops.conv3x3_dgrad(dy.cuda(), w.cuda(), base.cuda().clone(), accumulate=True)
    assert rel_err(acc, x.grad + base.double()) < TOL


@pytest.mark.parametrize("shape", [(2, 2, 16, 64, 64), (2, 16, 16, 64, 64), (3, 16, 32, 32, 32), (2, 128, 64, 16, 16), (2, 64, 64, 8, 8),
                                   (4, 16, 1, 32, 64), (1, 32, 16, 24, 40), (5, 32, 32, 16, 16)])
def test_conv_wgrad(shape):
    B, Cin, Cout, H, W = shape
    x, dy = rnd(B, Cin, H, W, seed=11), rnd(B, Cout, H, W, seed=12)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (ref_conv(x, w) * dy.double()).sum().backward()
    if Cout == 1:
        dw, db = ops.conv3x3_wgrad(x.cuda(), dy.cuda(), want_bias=True)
        assert rel_err(db, dy.double().sum((0, 2, 3))) < TOL
    else:
        dw = ops.conv3x3_wgrad(x.cuda(), dy.cuda())
    assert rel_err(dw, w.grad) < TOL
    # deterministic: bitwise identical on a second run
    dw2 = ops.conv3x3_wgrad(x.cuda(), dy.cuda(), want_bias=Cout == 1)
    dw2 = dw2[0] if Cout == 1 else dw2
    assert torch.equal(dw, dw2)


def test_conv_wgrad_affine():
    B, Cin, Cout, H, W = 2, 16, 32, 32, 32
    x, dy = rnd(B, Cin, H, W, seed=13), rnd(B, Cout, H, W, seed=14)
    sc, sh = 1 + 0.3 * rnd(Cin, seed=15), 0.2 * rnd(Cin, seed=16)
    a = F.relu(x.double() * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (ref_conv(a, w) * dy.double()).sum().backward()
    dw = ops.conv3x3_wgrad(x.cuda(), dy.cuda(), sc.cuda(), sh.cuda())
    assert rel_err(dw, w.grad) < TOL


def test_bn_train_finalize_and_eval_affine():
    C, n = 32, 2 * 64 * 64
    x = rnd(2, C, 64, 64, seed=17) * 3 + 1
    gamma, beta = 1 + 0.1 * rnd(C, seed=18), 0.1 * rnd(C, seed=19)
    rm, rv = 0.1 * rnd(C, seed=20), 1 + 0.1 * rnd(C, seed=21).abs()
    stats = torch.cat([x.double().sum((0, 2, 3)), (x.double() ** 2).sum((0, 2, 3))]).cuda()
    rm_d, rv_d = rm.cuda().clone(), rv.cuda().clone()
    sc, sh, mean, invstd = ops.bn_train_finalize(stats, gamma.cuda(), beta.cuda(), n, rm_d, rv_d)
    rm_r, rv_r = rm.clone(), rv.clone()
    y_ref = F.batch_norm(x, rm_r, rv_r, gamma, beta, True, 0.1, 1e-5)
    y = x.cuda() * sc[None, :, None, None] + sh[None, :, None, None]
    assert rel_err(y, y_ref) < TOL
    assert rel_err(rm_d, rm_r) < 1e-6 and rel_err(rv_d, rv_r) < 1e-6
    assert rel_err(mean, x.double().mean((0, 2, 3))) < 1e-6
    sc2, sh2 = ops.bn_eval_affine(gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda())
    y2_ref = F.batch_norm(x, rm.clone(), rv.clone(), gamma, beta, False, 0.1, 1e-5)
    assert rel_err(x.cuda() * sc2[None, :, None, None] + sh2[None, :, None, None], y2_ref) < TOL


@pytest.mark.parametrize("shape", [(2, 16, 64, 64), (3, 64, 8, 8), (1, 32, 32, 32)])
def test_bn_relu_bwd(shape):
    B, C, H, W = shape
    x = (rnd(B, C, H, W, seed=22) * 2 + 0.5).double().requires_grad_(True)
    gamma = (1 + 0.1 * rnd(C, seed=23)).double().requires_grad_(True)
    beta = (0.1 * rnd(C, seed=24)).double().requires_grad_(True)
    dY = rnd(B, C, H, W, seed=25)
    y = F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5))
    (y * dY.double()).sum().backward()
    xf = x.detach().float()
    stats = torch.cat([xf.double().sum((0, 2, 3)), (xf.double() ** 2).sum((0, 2, 3))]).cuda()
    sc, sh, mean, invstd = ops.bn_train_finalize(stats, gamma.detach().float().cuda(), beta.detach().float().cuda(), B * H * W)
    dx, dg, db = ops.bn_relu_bwd(dY.cuda(), xf.cuda(), sc, sh, mean, invstd, gamma.detach().float().cuda())
    assert rel_err(dx, x.grad) < 5e-5
    assert rel_err(dg, gamma.grad) < 5e-5 and rel_err(db, beta.grad) < 5e-5


def test_pool_residual_upcat():
    B, C, H, W = 2, 16, 32, 64
    raw = rnd(B, C, H, W, seed=26)
    sc, sh = 1 + 0.3 * rnd(C, seed=27), 0.2 * rnd(C, seed=28)
    act = F.relu(raw * sc[None, :, None, None] + sh[None, :, None, None])
    p = ops.act_avgpool2_fwd(raw.cuda(), sc.cuda(), sh.cuda())
    assert rel_err(p, F.avg_pool2d(act, 2, 2)) < 1e-6
    xr = rnd(B, C, H, W, seed=29)
    r = ops.act_residual_fwd(xr.cuda(), raw.cuda(), sc.cuda(), sh.cuda())
    assert rel_err(r, xr + act) < 1e-6
    # pool backward (+ accumulate)
    g = rnd(B, C, H // 2, W // 2, seed=30)
    xx = torch.zeros(B, C, H, W, dtype=torch.float64, requires_grad=True)
    (F.avg_pool2d(xx, 2, 2) * g.double()).sum().backward()
    assert rel_err(ops.avgpool2_bwd(g.cuda()), xx.grad) < 1e-6
    base = rnd(B, C, H, W, seed=31)
    assert rel_err(ops.avgpool2_bwd(g.cuda(), base.cuda().clone(), accumulate=True), xx.grad + base.double()) < 1e-6
    # upcat
    C2 = 8
    skip = rnd(B, C2, 2 * H, 2 * W, seed=32)
    sc2, sh2 = 1 + 0.3 * rnd(C2, seed=33), 0.2 * rnd(C2, seed=34)
    u = ops.act_upcat_fwd(raw.cuda(), sc.cuda(), sh.cuda(), skip.cuda(), sc2.cuda(), sh2.cuda())
    ref = torch.cat([F.interpolate(act, scale_factor=2, mode="bilinear", align_corners=True),
                     F.relu(skip * sc2[None, :, None, None] + sh2[None, :, None, None])], 1)
    assert rel_err(u, ref) < 2e-6
    # upcat backward
    go = rnd(B, C + C2, 2 * H, 2 * W, seed=35)
    lo = torch.zeros(B, C, H, W, dtype=torch.float64, requires_grad=True)
    (F.interpolate(lo, scale_factor=2, mode="bilinear", align_corners=True) * go[:, :C].double()).sum().backward()
    dlow, dskip = ops.upcat_bwd(go.cuda(), C)
    assert rel_err(dlow, lo.grad) < 1e-5
    assert torch.equal(dskip.cpu(), go[:, C:])


@pytest.mark.parametrize("shape", [(2, 16, 32, 32), (1, 8, 64, 64), (2, 4, 128, 128), (1, 3, 20, 128), (2, 5, 12, 24), (1, 2, 8, 8), (1, 2, 13, 32), (2, 3, 9, 64)])
def test_upcat_bwd_shapes(shape):
    """Adjoint of the bilinear x2 (align_corners=True) up-sampling: tiled kernel (power-of-two widths <= 128) and the generic one."""
    B, C, H, W = shape
    for C2 in (3, C):            # C2 == C: the tiled kernel also copies the skip half (no second launch)
        go = rnd(B, C + C2, 2 * H, 2 * W, seed=61)
        lo = torch.zeros(B, C, H, W, dtype=torch.float64, requires_grad=True)
        (F.interpolate(lo, scale_factor=2, mode="bilinear", align_corners=True) * go[:, :C].double()).sum().backward()
        dlow, dskip = ops.upcat_bwd(go.cuda(), C)
        assert rel_err(dlow, lo.grad) < 1e-5
        assert torch.equal(dskip.cpu(), go[:, C:])


def test_bicubic4_cat_matches_cv2_golden(golden):
    b = golden("bicubic.npz")
    lst = torch.from_numpy(b["lst"]).cuda()
    ndvi = rnd(2, 1, 256, 256, seed=36).cuda()
    x = sifnn_b200.bicubic4_cat(lst, ndvi)
    assert rel_err(x[:, :1], b["up_cv2"]) < 2e-6
    assert torch.equal(x[:, 1:], ndvi)
    edge = torch.arange(64 * 64, dtype=torch.float32).reshape(1, 1, 64, 64).cuda()  # exercises the clamped borders
    ref = F.interpolate(edge.cpu(), scale_factor=4, mode="bicubic", align_corners=False)
    assert rel_err(sifnn_b200.bicubic4_cat(edge, torch.zeros(1, 1, 256, 256, device="cuda"))[:, :1], ref) < 2e-6


def test_adam_matches_torch():
    n = 10007
    p0, g = rnd(n, seed=37), rnd(n, seed=38) * 0.01
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p, m, v = p0.cuda().clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    t = torch.zeros((), dtype=torch.int64, device="cuda")
    for i in range(5):
        gi = g * (i + 1)
        p_ref.grad = gi.clone()
        opt.step()
        ops.adam_step(p, (gi * 4).cuda(), m, v, t, 1e-3, grad_scale=0.25)
    assert int(t) == 5
    assert (p.cpu() - p_ref.detach()).abs().max() < 2e-7


def test_bad_arguments_raise():
    x = torch.zeros(1, 2, 8, 8)
    with pytest.raises(sifnn_b200.SifnnError):
        ops.conv3x3_fwd(x, torch.zeros(4, 2, 3, 3))  # CPU tensors: no fallback
    with pytest.raises(sifnn_b200.SifnnError):
        sifnn_b200._lib.call("sifnn_conv3x3_fwd", None, None, None, None, None, None, None, 1, 1, 1, 8, 8, None)


# The output layer's dedicated kernels (conv3x3_to1 / dgrad_from1 / wgrad_to1): partial row bands, the smallest image, several
# column groups per warp straddling rows, many channels, and widths the kernels do not take (generic path, same contract).
TO1_SHAPES = [(1, 16, 2, 4), (2, 16, 5, 8), (3, 7, 10, 12), (1, 64, 9, 36), (2, 3, 33, 260), (5, 16, 13, 20), (1, 16, 6, 10), (1, 70, 8, 8)]


@pytest.mark.parametrize("shape", TO1_SHAPES)
def test_single_output_channel_layer(shape):
    B, Cin, H, W = shape
    x, w, b = rnd(B, Cin, H, W, seed=71), rnd(1, Cin, 3, 3, seed=72, scale=0.2), rnd(1, seed=73)
    sc, sh = 1 + 0.3 * rnd(Cin, seed=74), 0.2 * rnd(Cin, seed=75)
    dy = rnd(B, 1, H, W, seed=76)
    # forward, plain and with the BatchNorm + ReLU prologue
    assert rel_err(ops.conv3x3_fwd(x.cuda(), w.cuda(), b.cuda()), ref_conv(x, w, b)) < TOL
    xa = torch.relu(x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    assert rel_err(ops.conv3x3_fwd(x.cuda(), w.cuda(), b.cuda(), sc.cuda(), sh.cuda()), ref_conv(xa, w, b)) < TOL
    # data gradient (padding adjoint folded in) and weight / bias gradient
    xr = xa.clone().requires_grad_(True)
    wr = w.double().clone().requires_grad_(True)
    br = b.double().clone().requires_grad_(True)
    (F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="replicate"), wr, br) * dy.double()).sum().backward()
    assert rel_err(ops.conv3x3_dgrad(dy.cuda(), w.cuda()), xr.grad) < TOL
    dw, db = ops.conv3x3_wgrad(x.cuda(), dy.cuda(), sc.cuda(), sh.cuda(), want_bias=True)
    assert rel_err(dw, wr.grad) < TOL and rel_err(db, br.grad) < TOL
    dw2 = ops.conv3x3_wgrad(x.cuda(), dy.cuda())
    wr2 = w.double().clone().requires_grad_(True)
    (ref_conv(x, wr2) * dy.double()).sum().backward()
    assert rel_err(dw2, wr2.grad) < TOL


@pytest.mark.parametrize("shape", [(2, 16, 16, 32, 32), (1, 4, 4, 5, 6), (3, 8, 8, 17, 34), (1, 32, 32, 8, 8), (2, 4, 8, 9, 10), (1, 6, 6, 7, 8), (2, 3, 5, 4, 4)])
def test_act_upcat_shapes(shape):
    """BatchNorm + ReLU + bilinear x2 (align_corners=True) + concat: the fused 4 + 4 channel kernel (C1 == C2, C1 % 4 == 0) and the
    generic one, odd heights, widths that leave a partial last CTA."""
    B, C1, C2, H, W = shape
    low, skip = rnd(B, C1, H, W, seed=81), rnd(B, C2, 2 * H, 2 * W, seed=82)
    s1, h1 = 1 + 0.3 * rnd(C1, seed=83), 0.2 * rnd(C1, seed=84)
    s2, h2 = 1 + 0.3 * rnd(C2, seed=85), 0.2 * rnd(C2, seed=86)
    u = ops.act_upcat_fwd(low.cuda(), s1.cuda(), h1.cuda(), skip.cuda(), s2.cuda(), h2.cuda())
    act = F.relu(low * s1[None, :, None, None] + h1[None, :, None, None])
    ref = torch.cat([F.interpolate(act, scale_factor=2, mode="bilinear", align_corners=True),
                     F.relu(skip * s2[None, :, None, None] + h2[None, :, None, None])], 1)
    assert u.shape == ref.shape and rel_err(u, ref) < 2e-6
