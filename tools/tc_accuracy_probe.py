import sys, os
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
import sifnn_b200
from sifnn_b200 import ops
torch.manual_seed(0)
for (B, ci, co, H, W) in [(2, 16, 16, 16, 128), (2, 64, 32, 16, 128), (2, 128, 64, 8, 128)]:
    x = torch.randn(B, ci, H, W); w = torch.randn(co, ci, 3, 3) * 0.1
    xa = x.abs() + 0.5   # all-positive data: a rounding BIAS shows up as a non-zero mean error
    for name, xx in (("randn", x), ("positive", xa)):
        ww = w if name == "randn" else w.abs()
        ref = F.conv2d(F.pad(xx.double(), (1, 1, 1, 1), mode="replicate"), ww.double())
        a = ops.conv3x3_fwd(xx.cuda(), ww.cuda()).cpu().double()
        b = ops.conv3x3_fwd_tc(xx.cuda(), ww.cuda()).cpu().double()
        c = F.conv2d(F.pad(xx, (1, 1, 1, 1), mode="replicate"), ww).double()
        sc = ref.abs().max()
        print(f"{ci}->{co} {name:8s}: max-rel err  simt {float((a-ref).abs().max()/sc):.2e}  tc {float((b-ref).abs().max()/sc):.2e}  torch-cpu {float((c-ref).abs().max()/sc):.2e} | mean signed err/scale simt {float((a-ref).mean()/sc):+.2e} tc {float((b-ref).mean()/sc):+.2e} cpu {float((c-ref).mean()/sc):+.2e}")
