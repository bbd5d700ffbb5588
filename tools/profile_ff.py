#!/usr/bin/env python3
"""Per-layer timing of the full-fold convolution (csrc/conv3x3_ff.cu) against the round-1 tensor-core kernels on the network's own layer
shapes, B = 32, CUDA events.  Also the per-layer roofline: max(FLOP / tensor roof of the split used, algorithmic bytes / HBM peak).

    python tools/profile_ff.py [--batch 32] [--tf32] [--only CIxCOxHW] [--reps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sifnn_b200  # noqa: E402
from sifnn_b200 import ops  # noqa: E402

LAYERS = [(16, 16, 256), (32, 16, 256), (16, 16, 128), (16, 32, 128), (64, 32, 128), (32, 16, 128), (32, 32, 64), (32, 64, 64),
          (64, 32, 64), (64, 64, 32)]
HBM = 6457.4e9


def timeit(fn, reps):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--kind", type=int, default=-1, help="0 BF16, 1 TF32, 2 FP16 forward + BF16 data gradient (default)")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--old", action="store_true", help="also time the round-1 kernels")
    ap.add_argument("--ff", action="store_true", help="also time the full-fold kernel on the wide layers")
    a = ap.parse_args()
    B = a.batch
    lib = sifnn_b200.load()
    kind = a.kind if a.kind >= 0 else (1 if a.tf32 else 2)
    lib.sifnn_conv3x3_ff_config(kind, 0)
    lib.sifnn_conv3x3_fs_config(kind, 0)
    roof = (732.4e12 if kind == 1 else 1670.5e12) / 3
    print(f"# full-fold convolution, {['BF16', 'TF32', 'FP16 forward / BF16 data gradient'][kind]} 3-term split, B = {B}; roofline = max(FLOP / {roof / 1e12:.0f} TFLOP/s, bytes / 6457 GB/s)")
    tot = {}
    for ci, co, hw in LAYERS:
        if a.only and a.only != f"{ci}x{co}x{hw}":
            continue
        x = torch.randn(B, ci, hw, hw, device="cuda")
        dy = torch.randn(B, co, hw, hw, device="cuda")
        w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
        sc, sh = torch.rand(ci, device="cuda") + 0.5, torch.randn(ci, device="cuda") * 0.1
        stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
        fl = 2.0 * B * ci * co * 9 * hw * hw
        byts = 4.0 * B * (ci + co) * hw * hw
        floor = max(fl / roof, byts / HBM)
        if lib.sifnn_conv3x3_fs_supported(ci, co, hw, hw):
            fns = {"fwd_fs": lambda: ops.conv3x3_fwd_fs(x, w, sc, sh, stats), "dgrad_fs": lambda: ops.conv3x3_dgrad_fs(dy, w)}
            if a.ff:
                fns["fwd_ff"] = lambda: ops.conv3x3_fwd_ff(x, w, sc, sh, stats)
                fns["dgrad_ff"] = lambda: ops.conv3x3_dgrad_ff(dy, w)
        else:
            fns = {"fwd_ff": lambda: ops.conv3x3_fwd_ff(x, w, sc, sh, stats), "dgrad_ff": lambda: ops.conv3x3_dgrad_ff(dy, w)}
        if a.old:
            fns["fwd_tc"] = lambda: ops.conv3x3_fwd_tc(x, w, None, sc, sh, stats)
            fns["dgrad_tc"] = lambda: ops.conv3x3_dgrad_tc(dy, w)
        row = []
        for name, fn in fns.items():
            t = timeit(fn, a.reps)
            tot[name] = tot.get(name, 0.0) + t
            row.append(f"{name} {t * 1e6:7.1f} us {fl / t / 1e12:6.1f} TF {byts / t / 1e9:6.0f} GB/s frac {floor / t:4.2f}")
        print(f"{ci:3d}->{co:3d} @{hw:3d} (floor {floor * 1e6:5.1f} us): " + " | ".join(row), flush=True)
    print("# totals (us): " + json.dumps({k: round(v * 1e6, 1) for k, v in tot.items()}))


if __name__ == "__main__":
    main()
