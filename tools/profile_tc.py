#!/usr/bin/env python3
"""Run the tcgen05 convolution on one layer shape (target of `ncu --set full -k regex:conv3x3_tc`).  usage: profile_tc.py CIxCOxHW [fwd|dgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops
ci, co, hw = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64x32x128").split("x"))
what = sys.argv[2] if len(sys.argv) > 2 else "fwd"
B = 32
x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
dy = torch.randn(B, co, hw, hw, device="cuda")
for _ in range(3):
    y = ops.conv3x3_fwd_tc(x, w) if what == "fwd" else ops.conv3x3_dgrad_tc(dy, w)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
