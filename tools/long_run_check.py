#!/usr/bin/env python3
"""300 SR1 + 300 SR2 training steps at batch 32 on the smoother synthetic distribution, tensor-core kernels vs strict fp32 (SIMT) from the same
initialisation: both must stay finite and decrease, and the two loss curves must stay together at the level of fp32 training noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import sifnn_b200, model as model_mod
import sifnn_oracle as O

for kind, alpha, gamma, lr in (("sr1", 0.99, -0.5, 1e-3), ("sr2", 0.5, -0.25, 1e-4)):
    curves = {}
    for tc in (True, False):
        sifnn_b200.set_tensor_cores(tc)
        sd = O.init_state_dict(0)
        m = model_mod.ModelB_2(2).cuda(); m.load_state_dict(sd); m.train()
        tr = sifnn_b200.Trainer(m, kind, alpha, gamma, lr)
        batches = [O.smooth_batch(32, seed=50 + i) for i in range(6)]
        batches = [(l.cuda(), n.cuda()) for l, _, n in batches]
        rec = torch.stack([tr.step(*batches[i % 6]) for i in range(300)]).cpu().numpy()
        assert np.isfinite(rec).all()
        curves[tc] = rec[:, 2]
    a, b = curves[True], curves[False]
    rel = np.abs(a - b) / np.abs(b)
    print(f"{kind}: loss {b[0]:.4f} -> tc {a[-1]:.5f} / strict {b[-1]:.5f}; |tc - strict| / strict: first 50 max {rel[:50].max():.2e}, all max {rel.max():.2e}, last 50 mean {rel[-50:].mean():.2e}")
    assert a[-1] < 0.7 * a[0] and b[-1] < 0.7 * b[0]
sifnn_b200.set_tensor_cores(True)
