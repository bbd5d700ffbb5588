#!/usr/bin/env python3
"""Which half of act_upcat costs what: up-sampled half only / skip half only / both, batch 32, 128^2 -> 256^2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops

B = 32
def timed(fn, reps=10):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for c1, c2 in ((16, 16), (16, 1), (1, 16), (32, 32), (8, 8)):
    low, skip = torch.randn(B, c1, 128, 128, device="cuda"), torch.randn(B, c2, 256, 256, device="cuda")
    s1, h1 = torch.rand(c1, device="cuda") + 0.5, torch.randn(c1, device="cuda") * 0.1
    s2, h2 = torch.rand(c2, device="cuda") + 0.5, torch.randn(c2, device="cuda") * 0.1
    t = timed(lambda: ops.act_upcat_fwd(low, s1, h1, skip, s2, h2))
    nb = (low.numel() + skip.numel() + B * (c1 + c2) * 256 * 256) * 4
    print(f"C1={c1:3d} C2={c2:3d}: {t:7.1f} us  {nb / 1e6:6.0f} MB  {nb / t / 1e3:6.0f} GB/s")
