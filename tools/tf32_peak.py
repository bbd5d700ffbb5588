"""Measured dense TF32 tensor-core peak on this GPU (cuBLAS through torch.matmul, 8192^3, best of 10, CUDA events),
the same method MEASURED_PEAKS.json uses for bf16.  Prints one JSON line."""
import json
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
best = {}
for name, (x, y) in {"tf32": (a, b), "bf16": (a.bfloat16(), b.bfloat16())}.items():
    for _ in range(3):
        x @ y
    t = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x @ y; e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    best[name + "_tflops"] = round(2 * n ** 3 / (min(t) * 1e-3) / 1e12, 1)
print(json.dumps(best))
