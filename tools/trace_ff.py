#!/usr/bin/env python3
"""clock64 trace of one CTA of the full-fold convolution: when does each role see each pipeline step?  usage: trace_ff.py [CIxCOxHW] [ablate] [tf32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import _lib

lib = sifnn_b200.load()
sh = sys.argv[1] if len(sys.argv) > 1 else "16x16x256"
ab = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tf32 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ci, co, hw = (int(v) for v in sh.split("x"))
B = 32
lib.sifnn_conv3x3_ff_config(tf32, 0)
x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
out = torch.empty(B, co, hw, hw, device="cuda")
wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(ci, co), dtype=torch.uint8, device="cuda")
tr = torch.zeros(16 * 256, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.call("sifnn_conv3x3_fwd_ff", x.data_ptr(), None, None, w.data_ptr(), out.data_ptr(), None, wprep.data_ptr(), B, ci, co, hw, hw, st)
lib.sifnn_conv3x3_ff_debug(ab)
run(); torch.cuda.synchronize()
lib.sifnn_conv3x3_ff_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
lib.sifnn_conv3x3_ff_trace(None); lib.sifnn_conv3x3_ff_debug(0)
t = tr.cpu().view(16, 256)
t0 = int(t[0, 0])
names = ["ld:top", "ld:issue", "xf:aempty", "xf:rawfull", "xf:done", "mma:accempty", "mma:afull", "mma:issued", "epi:accfull", "epi:loaded", "epi:barrier", "epi:done", "ld:issued"]
order = [11, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12]
print(f"# {sh} ablate={ab} tf32={tf32}; clocks since the first TMA issue")
print("step " + " ".join(f"{n:>12s}" for n in names))
for s in list(range(0, 12)) + list(range(40, 52)):
    print(f"{s:4d} " + " ".join(f"{int(t[e, s]) - t0:12d}" for e in order))
