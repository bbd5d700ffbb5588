#!/usr/bin/env python3
"""Per-layer timing of the weight-gradient kernels on the network's layer shapes, B = 32: wgrad_km (16-bit K-major, round 2) against wgrad_tc (TF32, round 1).
C-ABI calls with pre-allocated buffers, CUDA events, 10 repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import _lib

lib = sifnn_b200.load()
B = 32
LAYERS = [(16, 16, 256), (32, 16, 256), (16, 16, 128), (16, 32, 128), (64, 32, 128), (32, 16, 128), (32, 32, 64), (32, 64, 64), (128, 64, 64), (64, 32, 64), (64, 64, 32)]
tot = {"km": 0.0, "tc": 0.0}
for ci, co, hw in LAYERS:
    x = torch.randn(B, ci, hw, hw, device="cuda"); dy = torch.randn(B, co, hw, hw, device="cuda"); dw = torch.empty(co, ci, 3, 3, device="cuda")
    ws = torch.empty(max(lib.sifnn_conv3x3_wgrad_km_workspace(B, ci, co, hw, hw), lib.sifnn_conv3x3_wgrad_tc_workspace(B, ci, co, hw, hw), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    row = []
    for name in ("km", "tc"):
        fn = lambda: _lib.call(f"sifnn_conv3x3_wgrad_{name}", x.data_ptr(), None, None, dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), B, ci, co, hw, hw, st)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); e1.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e-3
        tot[name] += t
        fl = 2.0 * B * ci * co * 9 * hw * hw
        by = 4.0 * B * (ci + co) * hw * hw
        row.append(f"{name} {t * 1e6:7.1f} us {fl / t / 1e12:6.1f} TF {by / t / 1e9:6.0f} GB/s")
    print(f"{ci:3d}->{co:3d} @{hw:3d}: " + " | ".join(row), flush=True)
print("totals (us):", {k: round(v * 1e6, 1) for k, v in tot.items()})
