#!/usr/bin/env python3
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops, _lib
from sifnn_b200.ops import _p, _s
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
B, Cin, Cout, H, W = 1, 16, 16, 4, 16
for mode in ("ones", "ramp"):
    if mode == "ones":
        x = torch.ones(B, Cin, H, W, device="cuda"); dy = torch.ones(B, Cout, H, W, device="cuda")
    else:
        x = torch.arange(Cin, device="cuda", dtype=torch.float32).view(1, Cin, 1, 1).expand(B, Cin, H, W).contiguous() + 1
        dy = (torch.arange(Cout, device="cuda", dtype=torch.float32).view(1, Cout, 1, 1).expand(B, Cout, H, W).contiguous() + 1) * 0.5
    n = Cout * Cin * 9
    ws = torch.full((148 * n,), 7.0, device="cuda")
    dw = torch.full((Cout, Cin, 3, 3), -1.0, device="cuda")
    _lib.call("sifnn_conv3x3_wgrad_tc", _p(x), None, None, _p(dy), _p(dw), _p(ws), B, Cin, Cout, H, W, _s())
    torch.cuda.synchronize()
    print(mode, "dw[0,0]", dw[0, 0].tolist(), "dw[3,5]", dw[3, 5].tolist())
    p = ws[:n].view(Cout, Cin, 9)
    print(" partial slice0: untouched", int((p == 7.0).sum()), "zeros", int((p == 0).sum()), "of", n, " p[0,0]", p[0, 0].tolist(), "p[3,5]", p[3, 5].tolist())
    print(" ref dw[3,5]", ops.conv3x3_wgrad(x, dy)[3, 5].tolist())
