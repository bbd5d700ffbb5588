import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sifnn_b200
from sifnn_b200 import ops
lib = sifnn_b200.load()
lib.sifnn_tc_debug.argtypes = [ctypes.c_int]
lib.sifnn_tc_debug.restype = None
for ci, co, hw in [(16, 16, 256), (64, 32, 128), (32, 64, 128), (64, 64, 128)]:
    B = 32
    x = torch.randn(B, ci, hw, hw, device="cuda"); w = torch.randn(co, ci, 3, 3, device="cuda") * 0.1
    row = []
    for mode in (0, 1, 2):
        lib.sifnn_tc_debug(mode)
        ops.conv3x3_fwd_tc(x, w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): ops.conv3x3_fwd_tc(x, w)
        e1.record(); e1.synchronize()
        row.append(e0.elapsed_time(e1) / 3 * 1e3)
    lib.sifnn_tc_debug(0)
    units = B * hw * hw * ci / 1024
    stages = units / 2
    cyc = lambda us: us * 1e-6 * 1.92e9 * 148 / stages
    print(f"{ci}->{co}@{hw}: normal {row[0]:.0f} us, no-MMA {row[1]:.0f} us, 2x MMA {row[2]:.0f} us | cycles/stage(36 MMA): {cyc(row[0]):.0f} / {cyc(row[1]):.0f} / {cyc(row[2]):.0f} -> per extra MMA {(cyc(row[2]) - cyc(row[0])) / 36:.0f} cyc")
