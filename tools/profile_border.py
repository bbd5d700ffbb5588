#!/usr/bin/env python3
"""Target for ncu on the padding-adjoint columns pass: dgrad of a 16->16 layer at 256x256 and of 128->64 at 64x64, batch 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops
for ci, co, hw in ((16, 16, 256), (128, 64, 64)):
    dy = torch.randn(32, co, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
    for _ in range(3):
        dx = ops.conv3x3_dgrad_tc(dy, w)
torch.cuda.synchronize()
print("ok")
