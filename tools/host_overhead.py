#!/usr/bin/env python3
"""Host-side cost per call of the inference entry points at batch 4 (eager launches, graph replay, pipelined submit): host time vs total time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sifnn_b200, model as model_mod
m = model_mod.ModelB_2(2).cuda().eval()
l, n = torch.randn(4, 1, 64, 64).cuda(), torch.randn(4, 1, 256, 256).cuda()
def host_time(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6
with torch.inference_mode():
    print("eager  host %.0f us / total %.0f us" % host_time(lambda: m.forward_from_lowres(l, n)))
    m.enable_eval_graphs(True)
    print("graph  host %.0f us / total %.0f us" % host_time(lambda: m.forward_from_lowres(l, n)))
    lp, np_, op = l.cpu().pin_memory(), n.cpu().pin_memory(), torch.empty(4, 1, 256, 256).pin_memory()
    pipe = sifnn_b200.PipelinedInference(m)
    print("pipe+graph host %.0f us / total %.0f us" % host_time(lambda: pipe.submit(lp, np_, op)))
    m.enable_eval_graphs(False)
    print("pipe+eager host %.0f us / total %.0f us" % host_time(lambda: pipe.submit(lp, np_, op)))
