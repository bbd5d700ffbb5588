#!/usr/bin/env python3
"""Which operand-format pairs does kind::f16 accept in wgrad_km (0 FP16, 1 BF16)?  One sub-process per pair (an illegal instruction poisons the context)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0, %r)
import sifnn_b200
from sifnn_b200 import ops
fx, fd = int(sys.argv[1]), int(sys.argv[2])
sifnn_b200.load().sifnn_conv3x3_wgrad_km_config(fx, fd)
g = torch.Generator().manual_seed(1)
for (B, ci, co, H, W) in [(2, 16, 16, 8, 32), (2, 32, 16, 8, 64), (2, 16, 32, 8, 32), (1, 32, 32, 12, 128), (2, 128, 64, 4, 64)]:
    x = torch.randn(B, ci, H, W, generator=g); dy = torch.randn(B, co, H, W, generator=g)
    w = torch.zeros(co, ci, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="replicate"), w) * dy.double()).sum().backward()
    dw = ops.conv3x3_wgrad_km(x.cuda(), dy.cuda()).cpu().double()
    print(fx, fd, (B, ci, co, H, W), "rel err %%.2e" %% float((dw - w.grad).abs().max() / w.grad.abs().max()), flush=True)
''' % ROOT
for fx, fd in ((1, 1), (0, 0), (0, 1)):
    r = subprocess.run([sys.executable, "-c", code, str(fx), str(fd)], capture_output=True, text=True, timeout=200)
    print(r.stdout.strip() or "(no output)")
    if r.returncode != 0:
        print("  FAILED:", r.stderr.strip().splitlines()[-1][:200])
