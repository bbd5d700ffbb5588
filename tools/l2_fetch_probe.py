#!/usr/bin/env python3
"""Does cudaLimitMaxL2FetchGranularity change the column-strided kernels?  Sets the limit (argv[1] = 32|64|128),
prints what the driver reports back, then runs bench.py with the remaining arguments."""
import ctypes, os, runpy, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
if rt is None:
    import glob
    rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0])
LIMIT = 0x05  # cudaLimitMaxL2FetchGranularity
want = int(sys.argv[1])
v = ctypes.c_size_t(0)
rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT); before = v.value
rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(want))
rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
print(f"l2 fetch granularity: before={before} set({want}) rc={rc} after={v.value}", file=sys.stderr)
sys.argv = [os.path.join(root, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
