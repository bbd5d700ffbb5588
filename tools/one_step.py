#!/usr/bin/env python3
"""Two eager SR1 training steps at B = 32 (target of the `ncu --metrics gpu__time_duration.sum` launch list: the second step is steady state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200, model as model_mod
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = model_mod.ModelB_2(2).cuda().train()
tr = sifnn_b200.Trainer(m, "sr1", 0.99, -0.5, 1e-3)
lst, ndvi = torch.randn(B, 1, 64, 64, device="cuda"), torch.randn(B, 1, 256, 256, device="cuda")
for _ in range(2):
    out = tr.step(lst, ndvi)
torch.cuda.synchronize()
print("ok", out.tolist())
