#!/usr/bin/env python3
"""clock64 trace of one CTA of the fold + shift convolution: when does each role reach / leave each wait of each pipeline step?
usage: trace_fs.py [CIxCOxHW] [tf32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sifnn_b200
from sifnn_b200 import _lib

lib = sifnn_b200.load()
sh = sys.argv[1] if len(sys.argv) > 1 else "16x16x256"
tf32 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dgrad = len(sys.argv) > 3 and sys.argv[3] == "dgrad"   # CIxCOxHW names the LAYER: its data gradient convolves co -> ci channels
ci, co, hw = (int(v) for v in sh.split("x"))
B = 32
lib.sifnn_conv3x3_fs_config(tf32, 0)
x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
out = torch.empty(B, co, hw, hw, device="cuda")
wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(max(ci, co), max(ci, co)) + 2 * ci * co * 12, dtype=torch.uint8, device="cuda")
dy = torch.randn(B, co, hw, hw, device="cuda"); dx = torch.empty_like(x)
tr = torch.zeros(16 * 256, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    if dgrad:
        return _lib.call("sifnn_conv3x3_dgrad_fs", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), 0, wprep.data_ptr(), B, ci, co, hw, hw, st)
    _lib.call("sifnn_conv3x3_fwd_fs", x.data_ptr(), None, None, w.data_ptr(), out.data_ptr(), None, wprep.data_ptr(), B, ci, co, hw, hw, st)
run(); torch.cuda.synchronize()
lib.sifnn_conv3x3_fs_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
lib.sifnn_conv3x3_fs_trace(None)
t = tr.cpu().view(16, 256)
t0 = int(t[0, 0])
cols = [(11, "ld:top"), (0, "ld:issue"), (13, "xf:top"), (1, "xf:aempty"), (2, "xf:rawfull"), (3, "xf:done"), (12, "mma:top"), (4, "mma:accempty"), (5, "mma:afull"),
        (6, "mma:issued"), (14, "epi:top"), (7, "epi:accfull"), (8, "epi:loaded"), (10, "epi:done")]
if dgrad:
    cols += [(9, "edge:top"), (15, "edge:done")]
print(f"# {sh} kind={tf32} {'dgrad' if dgrad else 'fwd'}; clocks since the first TMA issue")
print("step " + " ".join(f"{n:>12s}" for _, n in cols))
for s in list(range(0, 10)) + list(range(40, 56)):
    print(f"{s:4d} " + " ".join(f"{int(t[e, s]) - t0:12d}" for e, _ in cols))
