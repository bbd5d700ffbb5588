#!/usr/bin/env python3
"""Pipeline study of the full-fold convolution: time one layer with parts of the kernel switched off (sifnn_conv3x3_ff_debug bits:
1 no MMAs, 2 no epilogue math, 4 no TMEM loads, 8 no transform, 16 no global stores, 32 loads always hit L2).  usage: ablate_ff.py [CIxCOxHW ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import ops, _lib

lib = sifnn_b200.load()
B = 32
shapes = [a for a in sys.argv[1:] if "x" in a] or ["16x16x256", "64x32x128"]
sets = [0, 1, 2, 4 | 2, 8, 16, 32, 1 | 8, 1 | 2 | 4, 1 | 2 | 4 | 8, 1 | 2 | 4 | 8 | 32, 2 | 4 | 16]
for tf32 in (0, 1):
    lib.sifnn_conv3x3_ff_config(tf32, 0)
    for sh in shapes:
        ci, co, hw = (int(v) for v in sh.split("x"))
        x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
        out = torch.empty(B, co, hw, hw, device="cuda")
        wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(ci, co), dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        def run():
            _lib.call("sifnn_conv3x3_fwd_ff", x.data_ptr(), None, None, w.data_ptr(), out.data_ptr(), None, wprep.data_ptr(), B, ci, co, hw, hw, st)
        row = []
        for ab in sets:
            lib.sifnn_conv3x3_ff_debug(ab)
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record(); e1.synchronize()
            row.append(f"{ab:2d}:{e0.elapsed_time(e1) / 10 * 1e3:6.1f}")
        lib.sifnn_conv3x3_ff_debug(0)
        print(f"{'tf32' if tf32 else 'bf16'} {sh:>10s} us by ablation mask | " + " ".join(row), flush=True)
