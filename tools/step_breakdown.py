#!/usr/bin/env python3
"""CUDA-event breakdown of one training step (B=32, SR1): input stage, forward, loss, decoder backward, encoder backward, Adam."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200, model as model_mod
from sifnn_b200 import _lib
from sifnn_b200.losses import loss_fwd_bwd
from sifnn_b200.model import bicubic4_cat, _stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = model_mod.ModelB_2(2).cuda().train()
tr = sifnn_b200.Trainer(m, "sr1", 0.99, -0.5, 1e-3)
lst, ndvi = torch.randn(B, 1, 64, 64, device="cuda"), torch.randn(B, 1, 256, 256, device="cuda")
for _ in range(3):
    tr.step(lst, ndvi)
st, opt = tr._opt_state(lst.device)
names = ["bicubic+cat", "forward", "loss fwd+bwd", "backward decoder", "backward encoder", "adam"]
acc = [0.0] * 6
R = 10
for _ in range(R):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record()
    x = bicubic4_cat(lst, ndvi); ev[1].record()
    y, ws, key = m._run_forward(x, train=True, keep=True); ev[2].record()
    losses, dsr = loss_fwd_bwd("sr1", y, lst, ndvi, 0.99, -0.5); ev[3].record()
    m._run_backward(x, dsr, ws, phase=1); ev[4].record()
    m._run_backward(x, dsr, ws, phase=2); ev[5].record()
    _lib.call("sifnn_adam_step", st["flat"].data_ptr(), st["fgrad"].data_ptr(), opt["m"].data_ptr(), opt["v"].data_ptr(), opt["t"].data_ptr(),
              1e-3, 0.9, 0.999, 1e-8, 1.0, st["n"], _stream()); ev[6].record()
    m._ws.give(key, ws)
    torch.cuda.synchronize()
    for i in range(6):
        acc[i] += ev[i].elapsed_time(ev[i + 1]) / R
tot = sum(acc)
for n, t in zip(names, acc):
    print(f"{n:18s} {t:7.3f} ms  {t / tot * 100:5.1f}%")
print(f"{'total':18s} {tot:7.3f} ms   ({B / tot * 1e3:.0f} patches/s eager)")
