#!/usr/bin/env python3
"""Time the fold + shift convolution on one layer in its variants (plain / affine / affine + stats / data gradient), C-ABI calls, 20 repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import _lib

lib = sifnn_b200.load()
kind = int(os.environ.get('KIND', '2'))
lib.sifnn_conv3x3_fs_config(kind, 0)
lib.sifnn_conv3x3_ff_config(kind, 0)
if os.environ.get("NARROW") == "1":      # M = 64 fold + shift path for the 64- / 32-pixel-wide levels, next to the full-fold kernel
    lib.sifnn_conv3x3_fs_narrow(1)
B = 32
for sh in (sys.argv[1:] or ["16x16x256", "32x16x256", "64x32x128"]):
    ci, co, hw = (int(v) for v in sh.split("x"))
    x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
    dy = torch.randn(B, co, hw, hw, device="cuda"); dx = torch.empty_like(x)
    out = torch.empty(B, co, hw, hw, device="cuda")
    sc, shf = torch.rand(ci, device="cuda") + 0.5, torch.randn(ci, device="cuda") * 0.1
    stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(max(ci, co), max(ci, co)) + 2 * ci * co * 12, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fns = {
        "plain": lambda: _lib.call("sifnn_conv3x3_fwd_fs", x.data_ptr(), None, None, w.data_ptr(), out.data_ptr(), None, wprep.data_ptr(), B, ci, co, hw, hw, st),
        "affine": lambda: _lib.call("sifnn_conv3x3_fwd_fs", x.data_ptr(), sc.data_ptr(), shf.data_ptr(), w.data_ptr(), out.data_ptr(), None, wprep.data_ptr(), B, ci, co, hw, hw, st),
        "stats": lambda: _lib.call("sifnn_conv3x3_fwd_fs", x.data_ptr(), None, None, w.data_ptr(), out.data_ptr(), stats.data_ptr(), wprep.data_ptr(), B, ci, co, hw, hw, st),
        "affine+stats": lambda: _lib.call("sifnn_conv3x3_fwd_fs", x.data_ptr(), sc.data_ptr(), shf.data_ptr(), w.data_ptr(), out.data_ptr(), stats.data_ptr(), wprep.data_ptr(), B, ci, co, hw, hw, st),
        "dgrad": lambda: _lib.call("sifnn_conv3x3_dgrad_fs", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), 0, wprep.data_ptr(), B, ci, co, hw, hw, st),
        "dgrad+acc": lambda: _lib.call("sifnn_conv3x3_dgrad_fs", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), 1, wprep.data_ptr(), B, ci, co, hw, hw, st),
    }
    if lib.sifnn_conv3x3_ff_supported(ci, co, hw, hw) and ci <= 64:
        fns["ff affine+stats"] = lambda: _lib.call("sifnn_conv3x3_fwd_ff", x.data_ptr(), sc.data_ptr(), shf.data_ptr(), w.data_ptr(), out.data_ptr(), stats.data_ptr(), wprep.data_ptr(), B, ci, co, hw, hw, st)
    if lib.sifnn_conv3x3_ff_supported(co, ci, hw, hw):
        fns["ff dgrad"] = lambda: _lib.call("sifnn_conv3x3_dgrad_ff", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), 0, wprep.data_ptr(), B, ci, co, hw, hw, st)
    if not lib.sifnn_conv3x3_fs_supported(ci, co, hw, hw):
        for k in ("plain", "affine", "stats", "affine+stats"):
            fns.pop(k)
    if not lib.sifnn_conv3x3_fs_supported(co, ci, hw, hw):
        for k in ("dgrad", "dgrad+acc"):
            fns.pop(k)
    row = []
    for name, fn in fns.items():
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); e1.synchronize()
        row.append(f"{name} {e0.elapsed_time(e1) / 20 * 1e3:6.1f}")
    print(f"{sh:>10s} us | " + " | ".join(row), flush=True)
