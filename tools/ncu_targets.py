#!/usr/bin/env python3
"""One launch (after one warm-up) of each kernel that gets an `ncu --set full` capture this round, B = 32, the network's own layer shapes:
   wgrad_km 16->16 @256 and 32->16 @256, conv3x3_fs data gradient 32->16 @256 (two output groups) and 16->16 @256, conv3x3_fs forward 32->16 @256,
   conv3x3_ff data gradient 128->64 @64.
ncu --set full --clock-control none --import-source on -k regex:'wgrad_km_kernel|conv3x3_fs_kernel|conv3x3_ff_kernel' -o gpurun_out/r2_full python tools/ncu_targets.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops

B = 32
r = lambda *s: torch.randn(*s, device="cuda")
jobs = [("wgrad_km 16x16x256", lambda: ops.conv3x3_wgrad_km(x1, d1)), ("wgrad_km 32x16x256", lambda: ops.conv3x3_wgrad_km(x2, d1)),
        ("fs dgrad 32x16x256", lambda: ops.conv3x3_dgrad_fs(d1, w2)), ("fs dgrad 16x16x256", lambda: ops.conv3x3_dgrad_fs(d1, w1)),
        ("fs fwd 32x16x256", lambda: ops.conv3x3_fwd_fs(x2, w2)), ("ff dgrad 128x64x64", lambda: ops.conv3x3_dgrad_ff(d3, w3))]
x1, x2, d1 = r(B, 16, 256, 256), r(B, 32, 256, 256), r(B, 16, 256, 256)
w1, w2 = r(16, 16, 3, 3) * 0.1, r(16, 32, 3, 3) * 0.1
d3, w3 = r(B, 64, 64, 64), r(64, 128, 3, 3) * 0.1
for name, fn in jobs:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    print(name, "ok", flush=True)
