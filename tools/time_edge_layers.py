#!/usr/bin/env python3
"""CUDA-event timing of the network's two edge layers through the per-op C-ABI (batch 32, 256x256): 2 -> 16 and 16 -> 1
(with the BatchNorm + ReLU prologue), forward / data gradient / weight gradient, against the same op in torch (fp32, no TF32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import sifnn_b200
from sifnn_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return ts[len(ts) // 2]


B, H = 32, 256
g = torch.Generator(device="cuda").manual_seed(0)
x16 = torch.randn(B, 16, H, H, device="cuda", generator=g)
w1 = torch.randn(1, 16, 3, 3, device="cuda", generator=g) * 0.1
bias = torch.randn(1, device="cuda", generator=g)
sc = torch.rand(16, device="cuda", generator=g) + 0.5
sh = torch.randn(16, device="cuda", generator=g) * 0.1
y = ops.conv3x3_fwd(x16, w1, bias, sc, sh)
ref = F.conv2d(F.pad(torch.relu(x16 * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)), (1, 1, 1, 1), mode="replicate"), w1, bias)
print("16->1 fwd rel err", float((y - ref).abs().max() / ref.abs().max()))
print(f"16->1 fwd (BN+ReLU prologue, bias): {timeit(lambda: ops.conv3x3_fwd(x16, w1, bias, sc, sh)):7.1f} us   HBM floor {(x16.numel() + y.numel()) * 4 / 6.5e12 * 1e6:5.1f} us")
x2 = torch.randn(B, 2, H, H, device="cuda", generator=g)
w0 = torch.randn(16, 2, 3, 3, device="cuda", generator=g) * 0.1
st = torch.zeros(2, 16, dtype=torch.float64, device="cuda")
print(f"2->16 fwd (+ BN statistics):        {timeit(lambda: ops.conv3x3_fwd(x2, w0, None, None, None, st)):7.1f} us   HBM floor {(x2.numel() + B * 16 * H * H) * 4 / 6.5e12 * 1e6:5.1f} us")
dy1 = torch.randn(B, 1, H, H, device="cuda", generator=g)
xr = x16[:4].clone().requires_grad_(True)
F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="replicate"), w1).backward(dy1[:4])
dxo = ops.conv3x3_dgrad(dy1[:4].contiguous(), w1)
print("16->1 dgrad rel err", float((dxo - xr.grad).abs().max() / xr.grad.abs().max()))
print(f"16->1 dgrad (1 -> 16 channels):     {timeit(lambda: ops.conv3x3_dgrad(dy1, w1)):7.1f} us   HBM floor {(dy1.numel() + x16.numel()) * 4 / 6.5e12 * 1e6:5.1f} us")
xa = torch.relu(x16[:4] * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
wr = w1.clone().requires_grad_(True); br = bias.clone().requires_grad_(True)
F.conv2d(F.pad(xa, (1, 1, 1, 1), mode="replicate"), wr, br).backward(dy1[:4])
dw, db = ops.conv3x3_wgrad(x16[:4].contiguous(), dy1[:4].contiguous(), sc, sh, want_bias=True)
print("16->1 wgrad rel err", float((dw - wr.grad).abs().max() / wr.grad.abs().max()), "bias", float((db - br.grad).abs().max() / br.grad.abs().max()))
print(f"16->1 wgrad (+ prologue, bias grad):  {timeit(lambda: ops.conv3x3_wgrad(x16, dy1, sc, sh, want_bias=True)):7.1f} us   HBM floor {(x16.numel() + dy1.numel()) * 4 / 6.5e12 * 1e6:5.1f} us")
dy16 = torch.randn(B, 16, H, H, device="cuda", generator=g)
print(f"2->16 wgrad:                         {timeit(lambda: ops.conv3x3_wgrad(x2, dy16)):7.1f} us   HBM floor {(x2.numel() + dy16.numel()) * 4 / 6.5e12 * 1e6:5.1f} us")
