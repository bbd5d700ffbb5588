#!/usr/bin/env python3
"""Accuracy and speed of the opt-in BF16x3 kx-folded convolution (run with SIFNN_TC_BF16=1 and without)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import sifnn_b200
from sifnn_b200 import ops
torch.manual_seed(0)
def t(fn, reps=5):
    fn(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize(); return e0.elapsed_time(e1) / reps * 1e3
for ci, co, hw in [(16, 16, 256), (32, 16, 256), (16, 16, 128), (32, 16, 128)]:
    B = 32
    x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1; dy = torch.randn(B, ci, hw, hw, device="cuda")
    xs, ws = x[:2].double(), w.double()
    ref = F.conv2d(F.pad(xs, (1, 1, 1, 1), mode="replicate"), ws)
    y = ops.conv3x3_fwd_tc(x, w)
    err = float((y[:2].double() - ref).abs().max() / ref.abs().max())
    xp = torch.relu(x)   # all-positive inputs: worst case for systematic rounding
    refp = F.conv2d(F.pad(xp[:2].double(), (1, 1, 1, 1), mode="replicate"), ws.abs())
    yp = ops.conv3x3_fwd_tc(xp, w.abs())
    errp = float((yp[:2].double() - refp).abs().max() / refp.abs().max())
    print(f"{ci:3d}->{co:3d} @{hw}: fwd {t(lambda: ops.conv3x3_fwd_tc(x, w)):7.1f} us  rel err {err:.2e}  (all-positive {errp:.2e})")
