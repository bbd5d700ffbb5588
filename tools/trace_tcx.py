#!/usr/bin/env python3
"""Pipeline trace of conv3x3_tcx_kernel (build with SIFNN_NVCC_EXTRA=-DSIFNN_TC_TRACE): clock64 stamps of CTA 0 for the first chunks / rows.
events: 0 loader issues TMA(g) | 1 transformer got stage (ab_empty) | 2 transformer got raw data | 3 transform done | 4 MMA thread got ab_full |
5 MMAs of chunk issued | 8 MMA thread got row slot (acc_empty) | 6 epilogue got acc_full(row) | 7 epilogue row math done"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops
ci, co, hw = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16x16x256").split("x"))
B = 32
x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
lib = sifnn_b200.load()
for _ in range(2):
    ops.conv3x3_fwd_tc(x, w)
tr = torch.zeros(9 * 64, dtype=torch.int64, device="cuda")
lib.sifnn_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.sifnn_debug_set_trace(tr.data_ptr())
ops.conv3x3_fwd_tc(x, w)
torch.cuda.synchronize()
lib.sifnn_debug_set_trace(None)
t = tr.cpu().view(9, 64)
t0 = int(t[0, 0])
names = {0: "load", 1: "xf:stage", 2: "xf:raw", 3: "xf:done", 4: "mma:abfull", 5: "mma:issued", 8: "mma:slot", 6: "epi:full", 7: "epi:done"}
print("chunk-indexed events (clocks since first TMA issue); 2 chunks per tile")
print(" g  " + "".join(f"{names[e]:>12s}" for e in (0, 1, 2, 3, 4, 5)))
for g in range(16, 32):
    print(f"{g:3d} " + "".join(f"{int(t[e, g]) - t0:12d}" for e in (0, 1, 2, 3, 4, 5)))
print("row-indexed events (4 rows per tile)")
print(" gr " + "".join(f"{names[e]:>12s}" for e in (8, 6, 7)))
for gr in range(32, 64):
    print(f"{gr:3d} " + "".join(f"{int(t[e, gr]) - t0:12d}" for e in (8, 6, 7)))
