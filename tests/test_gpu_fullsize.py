"""BASELINE.json's full sizes (batch 32, 256x256; -m gpu).

* configs[1] itself against the pinned oracle: one SR1 and one SR2 training step at B = 32 (losses, the flat 282 705-element gradient,
  the weights after Adam) in fp32, with the oracle's fp64 run as the noise yardstick (a B = 32 oracle step takes about a second on the
  host cores).  Tolerance: rel 1e-4 on losses / outputs / the whole gradient vector; per weight, |d| <= 1e-4 max|w| for more than 99.9 %
  and never more than 2 lr (Adam turns a noise-level gradient into a full +-lr step, in the reference's own fp32 run as much as here).
* every one of the 83 real MODIS pairs the reference ships, eval forward with the 1009 weights, against the oracle run live.
* size-independent algebraic properties: batch-permutation equivariance and chunk independence of the eval forward (bit-exact), linearity
  and adjointness of data gradient / weight gradient, the sum rule of the BatchNorm statistics of the convolution epilogue."""
import os

import numpy as np
import pytest
import torch

import model as model_mod
import sifnn_b200
import sifnn_oracle as O
from sifnn_b200 import ops
from conftest import GOLDEN, load_ckpt, rel_err

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


@pytest.mark.parametrize("kind,alpha,gamma,lr", [("sr1", 0.99, -0.5, 1e-3), ("sr2", 0.5, -0.25, 1e-4)])
def test_train_step_batch32_vs_oracle(kind, alpha, gamma, lr):
    """The benchmark's own configuration (BASELINE.json configs[1] / [2]): B = 32, 256x256, one full step, from the shipped 1009 weights."""
    sd = load_ckpt("1009")
    lst, up, ndvi = O.synthetic_batch(32, seed=1234)
    ref32 = O.Trainer(sd, kind, alpha, gamma, lr)
    ref64 = O.Trainer(sd, kind, alpha, gamma, lr, dtype=torch.float64)
    _, l32, _ = ref32.loss_and_grads(lst, up, ndvi)
    _, l64, _ = ref64.loss_and_grads(lst, up, ndvi)
    g32, g64 = ref32.flat_grads().clone(), ref64.flat_grads().clone()
    ref32.opt.step()
    m = model_mod.ModelB_2(2)
    m.load_state_dict(sd)
    m = m.cuda().train()
    tr = sifnn_b200.Trainer(m, kind, alpha, gamma, lr)
    x = torch.cat((up, ndvi), 1).cuda()
    y, ws, key = m._run_forward(x, train=True, keep=True)
    losses, dsr = sifnn_b200.loss_fwd_bwd(kind, y, lst.cuda(), ndvi.cuda(), alpha, gamma, want_grad=True)
    grads = m._run_backward(x, dsr, ws, phase=0).clone()
    m._ws.give(key, ws)
    assert np.allclose(losses.cpu().numpy(), np.array(l32), rtol=1e-4), (losses.cpu().numpy(), l32)
    e_ours, e_ref = rel_err(grads, g64), rel_err(g32, g64)
    print("\nB=32 %s: flat gradient rel.err vs oracle fp64: ours %.2e, reference fp32 %.2e" % (kind, e_ours, e_ref))
    assert rel_err(grads, g32) < 1e-4 and e_ours < 1e-4
    # the whole fused step (device bicubic included) from the same start: weights after Adam
    m2 = model_mod.ModelB_2(2)
    m2.load_state_dict(sd)
    m2 = m2.cuda().train()
    tr2 = sifnn_b200.Trainer(m2, kind, alpha, gamma, lr)
    got = tr2.step(lst.cuda(), ndvi.cuda()).cpu().numpy()
    assert np.allclose(got, np.array(l32), rtol=1e-4)
    p = torch.cat([q.detach().reshape(-1) for q in m2.parameters()]).cpu().double()
    want = ref32.flat_params().double()
    d = (p - want).abs()
    scale = float(want.abs().max())
    frac_ok = float((d <= 1e-4 * scale).double().mean())
    print("B=32 %s: post-Adam weights within 1e-4 of max|w|: %.5f of %d, worst %.2e (2 lr = %.0e)" % (kind, frac_ok, p.numel(), float(d.max()), 2 * lr))
    assert frac_ok > 0.999 and float(d.max()) <= 2.02 * lr
    # BatchNorm running statistics after the step
    rm = torch.cat([b.running_mean for b in m2.modules() if isinstance(b, torch.nn.BatchNorm2d)]).cpu()
    want_rm = torch.cat([ref32.sd[k] for k in ref32.sd if k.endswith("running_mean")])
    assert rel_err(rm, want_rm) < 1e-4


def test_eval_forward_all_83_real_pairs_vs_oracle():
    """Every real pair of the reference's test set (test_data_formatted/data), the predict.py:86-103 body: clip NDVI, z-score,
    bicubic x4 on the device, eval forward with the shipped SR1 weights -- against the oracle (torch bicubic == cv2 to 5e-7) run live."""
    d = np.load(os.path.join(GOLDEN, "real_pairs_all.npz"))
    assert d["lst"].shape == (83, 64, 64) and d["ndvi"].shape == (83, 256, 256)
    lst = torch.from_numpy((d["lst"] - O.MEAN_LST) / O.STD_LST).float()[:, None]
    ndvi = torch.from_numpy((np.clip(d["ndvi"], -1, 1) - O.MEAN_NDVI) / O.STD_NDVI).float()[:, None]
    sd = load_ckpt("1009")
    m = model_mod.ModelB_2(2)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with torch.inference_mode():
        y = m.forward_from_lowres(lst.cuda(), ndvi.cuda()).cpu()
        ref = torch.cat([O.forward(sd, torch.cat((O.bicubic_up4(lst[i:i + 8]), ndvi[i:i + 8]), 1)) for i in range(0, 83, 8)])
    errs = [(float((y[i] - ref[i]).abs().max() / ref[i].abs().max()), int(d["ids"][i])) for i in range(83)]
    worst = max(errs)
    print("\n83 real pairs: worst rel.err %.2e (pair id %d), median %.2e" % (worst[0], worst[1], float(np.median([e for e, _ in errs]))))
    assert worst[0] < 1e-4


@pytest.mark.parametrize("shape", [(32, 16, 16, 256), (32, 64, 32, 128), (32, 64, 64, 64)])
def test_round2_convolutions_adjoint_and_linear_full_size(shape):
    """The round-2 kernels (fold + shift for 128-multiple widths, full fold below) at B = 32: forward and COMPLETE data gradient (padding adjoint
    included) are adjoints of one bilinear form, and the data gradient is linear."""
    B, Cin, Cout, HW = shape
    fwd, dgrad = (ops.conv3x3_fwd_fs, ops.conv3x3_dgrad_fs) if HW % 128 == 0 else (ops.conv3x3_fwd_ff, ops.conv3x3_dgrad_ff)
    g = torch.Generator().manual_seed(12)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1).cuda()
    x = torch.randn(B, Cin, HW, HW, generator=g).cuda()
    d1, d2 = torch.randn(B, Cout, HW, HW, generator=g).cuda(), torch.randn(B, Cout, HW, HW, generator=g).cuda()
    a = 0.37
    assert rel_err(dgrad(a * d1 + d2, w), a * dgrad(d1, w) + dgrad(d2, w)) < 3e-5
    f = float((fwd(x, w).double() * d1.double()).sum())
    dx = dgrad(d1, w).double()
    scale = float(dx.norm()) * float(x.double().norm())   # Cauchy-Schwarz scale of the inner product (f itself is a heavily cancelling sum)
    assert abs(float((dx * x.double()).sum()) - f) < 1e-6 * scale
    # against the strict-fp32 SIMT kernels on the same inputs (forward: FP16 split, 22 bits; data gradient: BF16 split, 16 bits)
    assert rel_err(fwd(x, w), ops.conv3x3_fwd(x, w)) < 5e-6
    assert rel_err(dgrad(d1, w), ops.conv3x3_dgrad(d1, w)) < 3e-5
