#!/usr/bin/env python3
"""Targets for `ncu --set full` on the HBM-bound kernels: BatchNorm+ReLU backward (reduce, apply) on a 16-channel 256x256 layer at batch 32, the
SR1 loss kernel, act_upcat.  Also prints CUDA-event times and the achieved GB/s against the algorithmic bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops
from sifnn_b200.losses import loss_fwd_bwd

B, C, HW = 32, 16, 256
dY, raw = torch.randn(B, C, HW, HW, device="cuda"), torch.randn(B, C, HW, HW, device="cuda")
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
stats = torch.stack([raw.double().sum((0, 2, 3)), (raw.double() ** 2).sum((0, 2, 3))]).reshape(-1).contiguous()
sc, sh, mean, invstd = ops.bn_train_finalize(stats, gamma, beta, B * HW * HW)
y, lst, ndvi = torch.randn(B, 1, 256, 256, device="cuda"), torch.randn(B, 1, 64, 64, device="cuda"), torch.randn(B, 1, 256, 256, device="cuda")


def timed(fn, reps=10):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t = timed(lambda: ops.bn_relu_bwd(dY, raw, sc, sh, mean, invstd, gamma))
nbytes = B * C * HW * HW * 4
print(f"bn_relu_bwd (reduce + apply) 16@256^2 B=32: {t * 1e6:.1f} us, algorithmic 5 tensor passes = {5 * nbytes / 1e6:.0f} MB -> {5 * nbytes / t / 1e9:.0f} GB/s")
t = timed(lambda: loss_fwd_bwd("sr1", y, lst, ndvi, 0.99, -0.5, want_grad=True))
lb = B * 802816
print(f"loss_kernel SR1 B=32: {t * 1e6:.1f} us, algorithmic {lb / 1e6:.1f} MB -> {lb / t / 1e9:.0f} GB/s")
go = torch.randn(B, 32, 256, 256, device="cuda")
t = timed(lambda: ops.upcat_bwd(go, 16))
ub = (B * 32 * 256 * 256 + B * 16 * 128 * 128 + B * 16 * 256 * 256) * 4
print(f"upcat_bwd (low adjoint + skip copy) 16@128^2 <- 32@256^2 B=32: {t * 1e6:.1f} us, algorithmic {ub / 1e6:.0f} MB -> {ub / t / 1e9:.0f} GB/s")
low, skip = torch.randn(B, 16, 128, 128, device="cuda"), torch.randn(B, 16, 256, 256, device="cuda")
s16, h16 = torch.rand(16, device="cuda") + 0.5, torch.randn(16, device="cuda") * 0.1
t = timed(lambda: ops.act_upcat_fwd(low, s16, h16, skip, s16, h16))
fb = (B * 16 * 128 * 128 + B * 16 * 256 * 256 + B * 32 * 256 * 256) * 4
print(f"act_upcat 16@128^2 + 16@256^2 -> 32@256^2 B=32: {t * 1e6:.1f} us, algorithmic {fb / 1e6:.0f} MB -> {fb / t / 1e9:.0f} GB/s")
