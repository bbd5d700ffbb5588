#!/usr/bin/env python3
"""Run the heavy per-op kernels on the benchmark's own layer shapes (B=32) -- the target of
`ncu --set full -k regex:...` captures and of quick per-layer timing tables.

    python tools/profile_ops.py [--batch 32] [--time]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sifnn_b200  # noqa: E402
from sifnn_b200 import ops  # noqa: E402

LAYERS = [(2, 16, 256), (16, 16, 256), (16, 16, 128), (16, 32, 128), (32, 32, 64), (32, 64, 64), (64, 64, 32),
          (128, 64, 64), (64, 32, 64), (64, 32, 128), (32, 16, 128), (32, 16, 256), (16, 1, 256)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--tc", action="store_true", help="compare the tcgen05 kernels with the SIMT ones on the eligible layers")
    a = ap.parse_args()
    B = a.batch
    peak = ops.fp32_peak_tflops() if a.time else 0.0
    if a.time:
        print(f"fp32 FFMA peak {peak:.1f} TFLOP/s")
    for ci, co, hw in LAYERS:
        if a.only and a.only != f"{ci}x{co}x{hw}":
            continue
        x = torch.randn(B, ci, hw, hw, device="cuda")
        dy = torch.randn(B, co, hw, hw, device="cuda")
        w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
        sc, sh = torch.rand(ci, device="cuda") + 0.5, torch.randn(ci, device="cuda") * 0.1
        fl = 2.0 * B * ci * co * 9 * hw * hw
        fns = {"fwd": lambda: ops.conv3x3_fwd(x, w), "fwd_aff": lambda: ops.conv3x3_fwd(x, w, None, sc, sh),
               "dgrad": lambda: ops.conv3x3_dgrad(dy, w), "wgrad": lambda: ops.conv3x3_wgrad(x, dy, want_bias=(co == 1)),
               "wgrad_aff": lambda: ops.conv3x3_wgrad(x, dy, sc, sh, want_bias=(co == 1))}
        lib = sifnn_b200.load()
        if a.tc:
            fns = {}
            if co <= 64 and lib.sifnn_conv3x3_tc_supported(ci, co, hw, hw):
                fns["fwd"] = lambda: ops.conv3x3_fwd(x, w)
                fns["fwd_tc"] = lambda: ops.conv3x3_fwd_tc(x, w)
                fns["fwd_aff_tc"] = lambda: ops.conv3x3_fwd_tc(x, w, None, sc, sh)
            if lib.sifnn_conv3x3_tc_supported(co, ci, hw, hw):
                fns["dgrad"] = lambda: ops.conv3x3_dgrad(dy, w)
                fns["dgrad_tc"] = lambda: ops.conv3x3_dgrad_tc(dy, w)
            if not fns:
                continue
        row = []
        for name, fn in fns.items():
            fn()
            if a.time:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    fn()
                e1.record()
                e1.synchronize()
                t = e0.elapsed_time(e1) / 3 * 1e-3
                row.append(f"{name} {t * 1e6:8.1f} us {fl / t / 1e12:5.1f} TF ({fl / t / 1e12 / peak * 100:4.1f}%)")
            else:
                fn()
        torch.cuda.synchronize()
        if a.time:
            print(f"{ci:3d}->{co:3d} @{hw:3d}: " + " | ".join(row))


if __name__ == "__main__":
    main()
