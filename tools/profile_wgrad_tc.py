#!/usr/bin/env python3
"""Time the tensor-core weight gradient against the SIMT kernel on the network's layer shapes (B = 32), or run one
shape (argument CixCoxHW) as a target for `ncu --set full -k regex:wgrad_tc_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops

B = 32
LAYERS = [(16, 16, 256), (32, 16, 256), (16, 16, 128), (16, 32, 128), (64, 32, 128), (32, 16, 128), (32, 32, 64), (32, 64, 64), (128, 64, 64),
          (64, 32, 64), (64, 64, 32)]
COUNT = {(16, 16, 256): 2, (16, 16, 128): 2, (32, 32, 64): 2, (64, 64, 32): 3}


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


if len(sys.argv) > 1:
    ci, co, hw = (int(v) for v in sys.argv[1].split("x"))
    x, dy = torch.randn(B, ci, hw, hw, device="cuda"), torch.randn(B, co, hw, hw, device="cuda")
    for _ in range(3):
        dw = ops.conv3x3_wgrad_tc(x, dy)
    torch.cuda.synchronize()
    print("ok", float(dw.abs().mean()))
    sys.exit(0)

tot_tc = tot_simt = 0.0
for ci, co, hw in LAYERS:
    x, dy = torch.randn(B, ci, hw, hw, device="cuda"), torch.randn(B, co, hw, hw, device="cuda")
    sc, sh = torch.rand(ci, device="cuda") + 0.5, torch.randn(ci, device="cuda") * 0.1
    fl = 2.0 * B * hw * hw * ci * co * 9
    t_s = timeit(lambda: ops.conv3x3_wgrad(x, dy))
    t_t = timeit(lambda: ops.conv3x3_wgrad_tc(x, dy))
    t_ta = timeit(lambda: ops.conv3x3_wgrad_tc(x, dy, sc, sh))
    err = float((ops.conv3x3_wgrad_tc(x, dy) - ops.conv3x3_wgrad(x, dy)).abs().max() / ops.conv3x3_wgrad(x, dy).abs().max())
    n = COUNT.get((ci, co, hw), 1)
    tot_tc += n * t_ta; tot_simt += n * t_s
    print(f"{ci:3d}->{co:3d} @{hw:3d}: simt {t_s:7.1f} us {fl / t_s * 1e-6:6.1f} TF | tc {t_t:7.1f} us {fl / t_t * 1e-6:6.1f} TF | tc+affine {t_ta:7.1f} us "
          f"{fl / t_ta * 1e-6:6.1f} TF | rel diff {err:.1e}")
print(f"sum over the network's eligible layers: simt {tot_simt / 1e3:.3f} ms, tc {tot_tc / 1e3:.3f} ms")
