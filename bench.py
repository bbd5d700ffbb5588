#!/usr/bin/env python3
"""Benchmark of the SIF-NN-SR ModelB hot path (BASELINE.json: "ModelB train patches/s (64->256) at
1/2/4/8 B200; inference Mpix/s vs CPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train_sr1|train_sr2|infer]

A step = one full training step (bicubic input stage, forward, fused loss, backward, Adam) on one batch
of synthetic z-scored patches of the paper's shape: per GPU 32 x {LST (1,64,64), NDVI (1,256,256)} ->
(1,256,256) -- BASELINE.json configs[1] at N=1 ("SR1 training step, batch 32, fp32, 1xB200"); for N>1 the
per-GPU batch stays 32 (weak scaling, data parallel, two-bucket NCCL all-reduce).  Prints ONE JSON line.

--impl reference times the CPU implementation the reference would run (PyTorch CPU ops through the oracle
port -- the reference is pure Python/PyTorch and cannot travel to the GPU box) on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

PATCH_MPIX = 256 * 256 / 1e6
FWD_GFLOP, STEP_GFLOP = 3.605, 10.78  # per patch, SURVEY section 8d
# ncu --set full summaries (profiles/) of one launch of each kernel class, used for roofline.traffic
NCU_SUMMARIES = {"wgrad_km_kernel": "r2_ncu_full_wgrad_km_32x16x256_summary.csv",
                 "conv3x3_fs_kernel fwd": "r2_ncu_full_conv3x3_fs_fwd_32x16x256_summary.csv",
                 "conv3x3_fs_kernel dgrad": "r2_ncu_full_conv3x3_fs_dgrad_32x16x256_summary.csv",
                 "conv3x3_ff_kernel fwd": "r2j_ncu_full_conv3x3_ff_16x16x256_summary.csv",
                 "conv3x3_ff_kernel dgrad": "r2_ncu_full_conv3x3_ff_dgrad_128x64x64_summary.csv",
                 "wgrad_tc_kernel": "r1l_ncu_full_wgrad_tc_32x16x256_summary.csv",
                 "wgrad_kernel": "r1i_ncu_full_wgrad_16x16x256_summary.csv",
                 "conv3x3_tc_kernel fwd": "r1j_ncu_full_conv3x3_tc_64x32x128_summary.csv",
                 "conv3x3_tc_kernel dgrad": "r1q_ncu_full_conv3x3_tcx_dgrad_16x16x256_summary.csv"}
HYPER = {"train_sr1": ("sr1", 0.99, -0.5, 1e-3), "train_sr2": ("sr2", 0.5, -0.25, 1e-4)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops", 1590.0), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_batches(n_batches, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 1, 64, 64, generator=g), torch.randn(batch, 1, 256, 256, generator=g)) for _ in range(n_batches)]


# --------------------------------------------------------------------------------------------------
# reference arm: the CPU path the reference would run
# --------------------------------------------------------------------------------------------------
def reference_impl():
    """The CPU implementation timed by the reference arm and the cpu_baseline leg: the reference's OWN model.py / utils.py (byte copies under
    oracle/_ref/, made by oracle/make_ref.py wherever /root/reference exists) driven in the order of its train_step; the oracle port otherwise."""
    try:
        import ref_runner as R
        if R.available():
            R.load()
            return "reference", R
    except Exception as e:   # a missing third-party module of utils.py on this host: say so and fall back
        sys.stderr.write(f"bench: reference modules unusable ({e!r}); timing the oracle port\n")
    return "port", None


def cpu_train_stepper(mode, batch):
    import sifnn_oracle as O
    kind, alpha, gamma, lr = HYPER[mode]
    which, R = reference_impl()
    sd = O.init_state_dict(0)
    tr = R.RefTrainer(kind, alpha, gamma, lr, state_dict=sd) if R else O.Trainer(sd, kind, alpha, gamma, lr)
    lst, up, ndvi = O.synthetic_batch(batch)
    return which, (lambda: tr.step(lst, up, ndvi))


def cpu_infer_stepper(batch):
    import sifnn_oracle as O
    which, R = reference_impl()
    sd = O.init_state_dict(0)
    x = torch.randn(batch, 2, 256, 256, generator=torch.Generator().manual_seed(1234))
    if R:
        m = R.build_model(sd).eval()

        def step():
            with torch.inference_mode():
                m(x)
    else:
        def step():
            with torch.inference_mode():
                O.forward(sd, x, train=False)
    return which, step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mode = args.mode
    if mode in ("infer", "tile"):
        bs = 1
        which, step = cpu_infer_stepper(bs)
        per_step, unit, metric = bs * PATCH_MPIX, "Mpix/s", "ModelB inference Mpix/s"
        sample = "eval forward of 1 synthetic patch per step (batch 1, the way predict.py:86-103 calls the model)"
    else:
        kind = HYPER[mode][0]
        bs = args.ref_batch
        which, step = cpu_train_stepper(mode, bs)
        per_step, unit, metric = bs, "patches/s", f"ModelB {kind.upper()} train patches/s"
        sample = f"one full {kind.upper()} train step on {bs} synthetic patches per step (the benchmark's own batch), torch CPU fp32, all host threads"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    path = ("the reference's own model.py + utils.py (oracle/_ref), train_step order of train_model_B_*.py" if which == "reference"
            else "PyTorch CPU fp32 (oracle port of model.py + loss helpers)")
    out = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(mode, args.batch), "batch_per_step": bs, "reference_path": path},
           "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": which, "sample": sample},
           "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def workload_name(mode, batch):
    if mode == "tile":
        return f"ModelB whole-tile inference: LST 1200x1200 + NDVI 4800x4800 -> 4800x4800, 324 windows in batches of {batch}, block-partitioned over the ranks, fp32"
    if mode == "infer":
        return f"ModelB eval forward, batch {batch} synthetic 64x64 LST + 256x256 NDVI -> 256x256, fp32"
    return f"ModelB SIF-NN-{HYPER[mode][0].upper()} training step, batch {batch}/GPU synthetic patches (64->256), fp32"


# --------------------------------------------------------------------------------------------------
# per-kernel roofline (measured live, CUDA events on the launch stream)
# --------------------------------------------------------------------------------------------------
def tf32_peak_tflops():
    """Dense TF32 tensor-core peak of this GPU, measured the way MEASURED_PEAKS.json measures bf16: cuBLAS 8192^3, best of 10, CUDA events."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


LAYERS = [(2, 16, 256), (16, 16, 256), (16, 16, 128), (16, 16, 128), (16, 32, 128), (32, 32, 64), (32, 32, 64), (32, 64, 64),
          (64, 64, 32), (64, 64, 32), (64, 64, 32), (128, 64, 64), (64, 32, 64), (64, 32, 128), (32, 16, 128), (32, 16, 256),
          (16, 16, 256), (16, 1, 256)]


_RESULT_OUT = None


def emit(record):
    """The one JSON line, on the process's original stdout (see main)."""
    f = _RESULT_OUT or sys.stdout
    f.write(json.dumps(record) + "\n")
    f.flush()


def kernel_rooflines(batch, fp32_peak, hbm_gbs, bf16_tflops, tf32_tflops, forward_only=False):
    """Times every convolution kernel class on the workload's own 18 layer shapes through the per-op C-ABI (pre-allocated buffers, CUDA events on
    the launch stream), with the kernel the network plan picks for each layer.  Per layer the roofline time is
    max(FLOPs / peak of the unit used, algorithmic bytes / measured HBM bandwidth); a class reports sum(roofline) / sum(measured).
    Unit peaks: 16-bit 3-term split (FP16 forward, BF16 data gradient) = bf16 / 3; TF32 3-term split (round-1 kernels, weight gradient) =
    measured TF32 / 3; SIMT = measured FFMA peak.  Algorithmic bytes of a layer: read the input once + write the output once (fp32)."""
    import sifnn_b200
    from sifnn_b200 import _lib
    lib = sifnn_b200.load()
    dev = "cuda"
    use_tc = sifnn_b200.tensor_cores_enabled()
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=5):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    res = {}

    def add(name, unit, t, fl, by):
        peak = {"tensor16": bf16_tflops / 3.0, "tensor32": tf32_tflops / 3.0, "fp32": fp32_peak}[unit]
        r = res.setdefault(name, {"t": 0.0, "fl": 0.0, "by": 0.0, "roof": 0.0, "roof_c": 0.0, "roof_m": 0.0, "n": 0, "unit": unit, "peak": peak})
        tc, tm = fl / (peak * 1e12), by / (hbm_gbs * 1e9)
        r["t"] += t; r["fl"] += fl; r["by"] += by; r["roof"] += max(tc, tm); r["roof_c"] += tc; r["roof_m"] += tm; r["n"] += 1

    for i, (ci, co, hw) in enumerate(LAYERS):
        x = torch.randn(batch, ci, hw, hw, device=dev)
        dy = torch.randn(batch, co, hw, hw, device=dev)
        y = torch.empty_like(dy)
        dx = torch.empty_like(x)
        w = torch.randn(co, ci, 3, 3, device=dev) * 0.1
        dw = torch.empty_like(w)
        wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(max(ci, 8), max(co, 8)) + lib.sifnn_conv3x3_tc_wprep_bytes(max(co, 8), max(ci, 8)) + 2 * ci * co * 12 + 256,
                            dtype=torch.uint8, device=dev)
        ws = torch.empty(max(lib.sifnn_conv3x3_wgrad_workspace(batch, ci, co, hw, hw), lib.sifnn_conv3x3_wgrad_tc_workspace(batch, ci, co, hw, hw),
                             lib.sifnn_conv3x3_wgrad_km_workspace(batch, ci, co, hw, hw), 16),
                         dtype=torch.uint8, device=dev)
        fl = 2.0 * batch * ci * co * 9 * hw * hw
        by = 4.0 * batch * (ci + co) * hw * hw
        P = lambda t: t.data_ptr()
        # forward
        if use_tc and (lib.sifnn_conv3x3_fs_supported(ci, co, hw, hw) or (hw == 64 and ci == 32 and co in (32, 64))):   # + the plan's M = 64 forward layers
            add("conv3x3_fs_kernel (fwd, FP16 split)", "tensor16", timed(lambda: _lib.call("sifnn_conv3x3_fwd_fs", P(x), None, None, P(w), P(y), None, P(wprep), batch, ci, co, hw, hw, st)), fl, by)
        elif use_tc and lib.sifnn_conv3x3_ff_supported(ci, co, hw, hw):
            add("conv3x3_ff_kernel (fwd, FP16 split)", "tensor16", timed(lambda: _lib.call("sifnn_conv3x3_fwd_ff", P(x), None, None, P(w), P(y), None, P(wprep), batch, ci, co, hw, hw, st)), fl, by)
        elif use_tc and co <= 64 and lib.sifnn_conv3x3_tc_supported(ci, co, hw, hw):
            add("conv3x3_tc_kernel (fwd, TF32 split, round 1)", "tensor32", timed(lambda: _lib.call("sifnn_conv3x3_fwd_tc", P(x), None, None, P(w), None, P(y), None, P(wprep), batch, ci, co, hw, hw, st)), fl, by)
        else:
            add("conv3x3_kernel 2->16 + conv3x3_to1_kernel 16->1 (fwd, SIMT)", "fp32", timed(lambda: _lib.call("sifnn_conv3x3_fwd", P(x), None, None, P(w), None, P(y), None, batch, ci, co, hw, hw, st)), fl, by)
        if not forward_only:
            if i > 0:   # no data gradient for the first layer
                if use_tc and lib.sifnn_conv3x3_fs_supported(co, ci, hw, hw):
                    add("conv3x3_fs_kernel (dgrad, BF16 split)", "tensor16", timed(lambda: _lib.call("sifnn_conv3x3_dgrad_fs", P(dy), P(w), P(dx), 0, P(wprep), batch, ci, co, hw, hw, st)), fl, by)
                elif use_tc and lib.sifnn_conv3x3_ff_supported(co, ci, hw, hw):
                    add("conv3x3_ff_kernel (dgrad, BF16 split)", "tensor16", timed(lambda: _lib.call("sifnn_conv3x3_dgrad_ff", P(dy), P(w), P(dx), 0, P(wprep), batch, ci, co, hw, hw, st)), fl, by)
                elif use_tc and lib.sifnn_conv3x3_tc_supported(co, ci, hw, hw):
                    add("conv3x3_tc_kernel (dgrad, TF32 split, round 1)", "tensor32", timed(lambda: _lib.call("sifnn_conv3x3_dgrad_tc", P(dy), P(w), P(dx), 0, P(wprep), batch, ci, co, hw, hw, st)), fl, by)
                else:
                    add("dgrad_from1_kernel 16->1 (dgrad, SIMT)", "fp32", timed(lambda: _lib.call("sifnn_conv3x3_dgrad", P(dy), P(w), P(dx), 0, batch, ci, co, hw, hw, st)), fl, by)
            if use_tc and co > 1 and lib.sifnn_conv3x3_wgrad_km_supported(ci, co, hw, hw):   # the plan's rule (modelb.cu wgrad)
                # the kernel itself, like in the network plan, where the per-CTA partials of all layers are summed by one launch per backward phase
                slots = ctypes.c_int(0)
                add("wgrad_km_kernel (BF16 split)", "tensor16", timed(lambda: _lib.call("sifnn_conv3x3_wgrad_km_partials", P(x), None, None, P(dy), P(ws), batch, ci, co, hw, hw, st, ctypes.addressof(slots))), fl, by)
                _lib.call("sifnn_wgrad_reduce", P(ws), P(dw), co * ci * 9, slots.value, st)
            elif use_tc and lib.sifnn_conv3x3_wgrad_tc_supported(ci, co, hw, hw):
                add("wgrad_tc_kernel (TF32 split)", "tensor32", timed(lambda: _lib.call("sifnn_conv3x3_wgrad_tc", P(x), None, None, P(dy), P(dw), P(ws), batch, ci, co, hw, hw, st)), fl, by)
            else:
                db = torch.empty(co, device=dev) if co == 1 else None
                add("wgrad_kernel (SIMT)", "fp32", timed(lambda: _lib.call("sifnn_conv3x3_wgrad", P(x), None, None, P(dy), P(dw), None if db is None else P(db), P(ws), batch, ci, co, hw, hw, st)), fl, by)
        del x, dy, y, dx, w, dw, wprep, ws
    out = []
    for name, r in res.items():
        hbm_bound = r["roof_m"] > r["roof_c"]
        out.append({"kernel": name, "bound": "hbm" if hbm_bound else ("tensor" if r["unit"].startswith("tensor") else "fp32"), "launches_per_step": r["n"],
                    "ms_per_step": r["t"] * 1e3, "achieved_tflops": r["fl"] / r["t"] / 1e12, "achieved_gbs": r["by"] / r["t"] / 1e9,
                    "unit_peak_tflops": r["peak"], "hbm_peak_gbs": hbm_gbs, "roofline_ms": r["roof"] * 1e3, "frac": r["roof"] / r["t"]})
    return out


def timed_events(fn, steps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / steps


def sr2_subrecord(model_mod, dev, B, world, rank, devb, nb, steps):
    """BASELINE.json configs[2]'s loss: the SR2 (gradFTM) training step at the same per-GPU batch, CUDA-graph replay, device time."""
    import sifnn_b200
    torch.manual_seed(0)
    m2 = model_mod.ModelB_2(in_channels=2).to(dev).train()
    kind, alpha, gamma, lr = HYPER["train_sr2"]
    tr2 = sifnn_b200.Trainer(m2, kind, alpha, gamma, lr)
    tr2.broadcast_parameters(0)
    tr2.capture(*devb[0])
    dt = timed_events(lambda i: tr2.step_graph(*devb[i % nb]), steps)
    tr2.release_graphs()   # captured NCCL collectives pin the communicator until their graphs are gone
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"metric": "ModelB SR2 train patches/s", "value": B * world / float(t[0]), "unit": "patches/s", "ms_per_step": float(t[0]) * 1e3,
            "batch_per_gpu": B, "global_batch": B * world, "steps": steps, "hyper": {"alpha": alpha, "gamma": gamma, "lr": lr}}


def inference_subrecord(m, dev, steps_small=50):
    """The inference half of the headline metric (BASELINE.json configs[0], [3]): eval forward from the 64x64 LST + 256x256 NDVI patches at
    batch 1 (predict.py's own call pattern; CUDA-graph replay), 32 and 4096 (chunks of 32 inside the module), device Mpix/s and end to end
    from pinned host buffers through PipelinedInference, plus the reference's CPU path at batch 1."""
    import sifnn_b200
    was = m.training
    m.eval()
    rows = []
    g = torch.Generator().manual_seed(99)
    for bsz, steps in ((1, steps_small), (32, 10), (4096, 2)):
        nbuf = 8 if bsz <= 32 else 1   # rotate distinct inputs (at 4096 one batch is 1.1 GB: far beyond L2 anyway)
        host = [(torch.randn(bsz, 1, 64, 64, generator=g).pin_memory(), torch.randn(bsz, 1, 256, 256, generator=g).pin_memory()) for _ in range(nbuf)]
        devb = [(l.to(dev), n.to(dev)) for l, n in host]
        m.enable_eval_graphs(bsz <= 8, max_batch=8)

        def step(i):
            with torch.inference_mode():
                return m.forward_from_lowres(*devb[i % nbuf])
        dt = timed_events(step, steps, warm=2)
        out_h = [torch.empty((bsz, 1, 256, 256), dtype=torch.float32).pin_memory() for _ in range(2)]
        pipe = sifnn_b200.PipelinedInference(m)
        for i in range(8 if bsz <= 32 else 2):
            pipe.submit(*host[i % nbuf], out_h[i % 2])
        pipe.flush(); torch.cuda.synchronize()
        windows = []   # three timed windows, the median is reported (a 13 ms window at batch 1 is at the mercy of one host hiccup)
        for _ in range(3 if bsz <= 32 else 1):
            t0 = time.perf_counter()
            for i in range(steps):
                pipe.submit(*host[i % nbuf], out_h[i % 2])
            pipe.flush(); torch.cuda.synchronize()
            windows.append((time.perf_counter() - t0) / steps)
        e2e = sorted(windows)[len(windows) // 2]
        rows.append({"batch": bsz, "mpix_s_device": bsz * PATCH_MPIX / dt, "mpix_s_e2e": bsz * PATCH_MPIX / e2e, "ms_per_batch": dt * 1e3,
                     "h2d_bytes_per_batch": bsz * (64 * 64 + 256 * 256) * 4, "d2h_bytes_per_batch": bsz * 256 * 256 * 4,
                     "launch": "cuda-graph replay" if bsz <= 8 else "eager, chunks of 32"})
        del host, devb, out_h, pipe
        torch.cuda.empty_cache()
    m.enable_eval_graphs(False)
    m.train(was)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    which, cstep = cpu_infer_stepper(1)
    cstep()
    t0 = time.perf_counter()
    for _ in range(10):
        cstep()
    cdt = (time.perf_counter() - t0) / 10
    return {"metric": "ModelB inference Mpix/s", "unit": "Mpix/s", "operand_split": "FP16 3-term (22 significant bits)", "sweep": rows,
            "cpu_baseline": {"value": PATCH_MPIX / cdt, "unit": "Mpix/s", "cores": cores, "kind": which, "sample": "10 eval forwards of 1 patch after 1 warm-up"}}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import sifnn_b200
    import model as model_mod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a GPU: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints a banner)
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    lib = sifnn_b200.load()
    torch.manual_seed(0)
    m = model_mod.ModelB_2(in_channels=2, downchannels=[16, 32, 64, 128], padding_mode="replicate", activation="ReLU",
                           bilinear=1, n_bridge_blocks=1).to(dev)
    B = args.batch
    mode = args.mode
    nb = 24 if mode != "infer" else 8  # distinct input batches rotated through: 24 x 8.9 MB = 214 MB > 126 MB L2
    host = [(l.pin_memory(), n.pin_memory()) for l, n in make_batches(nb, B, 1234 + rank)]
    devb = [(l.to(dev), n.to(dev)) for l, n in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if mode == "tile":
        # BASELINE.json configs[4]: one MODIS tile (LST 1200x1200 K, NDVI 4800x4800) -> 4800x4800, 324 windows of 64x64 block-partitioned over
        # the ranks (predict.py:81-103); every rank holds the tile and writes its own windows, no collective
        m.eval()
        stats = dict(mean_lst=307.24, std_lst=5.57, mean_ndvi=0.645, std_ndvi=0.168)
        g = torch.Generator().manual_seed(4321)
        tiles_h = [((300 + 8 * torch.rand(1200, 1200, generator=g)).pin_memory(), (0.6 + 0.3 * torch.randn(4800, 4800, generator=g)).pin_memory())
                   for _ in range(2)]
        tiles_d = [(l.to(dev), n.to(dev)) for l, n in tiles_h]
        out_d = torch.zeros(4800, 4800, dtype=torch.float32, device=dev)
        out_h = torch.empty(4800, 4800, dtype=torch.float32).pin_memory()
        nwin = 324
        unit, per_step, metric = "Mpix/s", nwin * PATCH_MPIX, "ModelB whole-tile inference Mpix/s"

        def step(i):
            l, n = tiles_d[i % 2]
            return sifnn_b200.super_resolve_tile(m, l, n, stats, batch=B, rank=rank, world_size=world, out=out_d)

        def step_host(i):   # each rank moves only the rows its windows touch (super_resolve_tile_host)
            l, n = tiles_h[i % 2]
            return sifnn_b200.super_resolve_tile_host(m, l, n, stats, batch=B, rank=rank, world_size=world, out_host=out_h, device=dev)[0]
        y0, y1, _, _ = sifnn_b200.owned_rows(1200, 1200, rank, world)
        h2d, d2h = (64 * 1200 + 256 * 4800) * 4 * (y1 - y0), 256 * 4800 * 4 * (y1 - y0)   # rank 0's share (upper bound of its result bytes)
        flop_per_step = FWD_GFLOP * 1e9 * nwin / n_gpus
    elif mode == "infer":
        m.eval()
        if B <= 8:
            m.enable_eval_graphs(True, max_batch=8)   # small batches (predict.py runs batch 1) are launch-bound: replay a captured graph
        unit, per_step, metric = "Mpix/s", B * PATCH_MPIX * n_gpus, "ModelB inference Mpix/s"

        def step(i):
            with torch.inference_mode():
                return m.forward_from_lowres(*devb[i % nb])

        host_out = [torch.empty((B, 1, 256, 256), dtype=torch.float32).pin_memory() for _ in range(2)]   # the caller's result buffers (pinned)
        pipe = sifnn_b200.PipelinedInference(m)   # H2D of step i+1 and D2H of step i-1 overlap the forward of step i; flushed by the barrier below

        def step_host(i):
            l, n = host[i % nb]
            pipe.submit(l, n, host_out[i % 2])
            return host_out[i % 2]
        h2d, d2h = B * (64 * 64 + 256 * 256) * 4, B * 256 * 256 * 4
        flop_per_step = FWD_GFLOP * 1e9 * B
    else:
        kind, alpha, gamma, lr = HYPER[mode]
        m.train()
        tr = sifnn_b200.Trainer(m, kind, alpha, gamma, lr)
        tr.broadcast_parameters(0)
        unit, per_step, metric = "patches/s", B * n_gpus, f"ModelB {kind.upper()} train patches/s"

        use_graph = not args.no_graph
        if use_graph:
            tr.capture(*devb[0])  # whole step as one CUDA graph; inputs are copied into its static buffers each step

            def step(i):
                return tr.step_graph(*devb[i % nb])
        else:
            def step(i):
                return tr.step(*devb[i % nb])

        losses_host = [torch.zeros(3, dtype=torch.float64).pin_memory() for _ in range(4)]

        def step_host(i):
            if use_graph:   # H2D of step i+1 overlaps step i (copy stream); the three loss scalars come back asynchronously every step
                tr.step_host_async(*host[i % nb], losses_host[i % 4])
                return losses_host[i % 4]
            return tr.step_host(*host[i % nb])
        h2d, d2h = B * (64 * 64 + 256 * 256) * 4, 3 * 8
        flop_per_step = STEP_GFLOP * 1e9 * B

    n_warm = max(args.warmup, 3) + (2 if world > 1 else 0)   # + 2 under NCCL: the first replays of the captured collectives still set up channels
    for i in range(n_warm):
        step(i)
    barrier()
    launches0 = lib.sifnn_launch_count()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    launches = lib.sifnn_launch_count() - launches0
    if mode.startswith("train") and getattr(tr, "_graph", None) is not None:
        launches = tr.graph_launches * args.steps  # graph replay: the library's counter only saw the capture
    # end-to-end: pinned host buffers in, result scalars (or the SR image) out, every step
    def flush_host():
        if mode == "infer":
            pipe.flush()   # every result has landed in its pinned host buffer
    for i in range(2):
        step_host(i)
    flush_host()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i)
    flush_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()

    out = None
    if rank == 0:
        hbm, bf16, how = peaks()
        value = per_step * args.steps / (ms * 1e-3)
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": n_gpus, "steps": args.steps, "warmup": n_warm,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if mode == "tile" else "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": workload_name(mode, B), "batch_per_gpu": B, "global_batch": B * n_gpus,
                          "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single", "batchnorm": "local (per-rank statistics)",
                          "weights": "random init, seed 0", "launch": "cuda-graph replay" if (mode.startswith("train") and not args.no_graph) else "eager", "l2": ("two 98 MB tiles alternated + 1.4 GB of activations per batch of windows (> 126 MB L2)" if mode == "tile" else
                                 f"{nb} distinct input batches rotated + {(2.7 if mode.startswith('train') else 1.4) * B / 32:.1f} GB of activations per step (> 126 MB L2)")},
               "e2e": {"value": per_step * args.steps / (e2e_ms * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": int(launches), "clocks": clk.summary(),
               "achieved_tflops": flop_per_step * n_gpus * args.steps / (ms * 1e-3) / 1e12}
    if rank == 0 and not args.no_roofline:
        from sifnn_b200 import ops
        fp32_peak = ops.fp32_peak_tflops()
        tf32 = tf32_peak_tflops()
        per_kernel = kernel_rooflines(B, fp32_peak, hbm, bf16, tf32, forward_only=mode in ("infer", "tile"))
        top = max(per_kernel, key=lambda r: r["ms_per_step"])
        traffic, traffic_source = None, None
        try:  # DRAM bytes of one launch of the dominant kernel class from the committed `ncu --set full` capture (profiles/)
            import csv
            key = top["kernel"].split(" ")[0] + (" dgrad" if "dgrad" in top["kernel"] else (" fwd" if "fwd" in top["kernel"] else ""))
            name = NCU_SUMMARIES.get(key)
            vals = {r[0]: r[1] for r in csv.reader(open(os.path.join(ROOT, "profiles", name))) if len(r) >= 2}
            traffic = (float(vals["dram__bytes_read.sum"]) + float(vals["dram__bytes_write.sum"])) * 1e6   # bytes of that one launch
            traffic_source = {"launch": name.replace("_summary.csv", ""), "file": "profiles/" + name,
                              "metric": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, ONE launch of the dominant kernel class (the layer named in the file)"}
        except Exception:
            pass
        hbm_bound = top["bound"] == "hbm"
        out["roofline"] = {"bound": "hbm" if hbm_bound else "tensor", "kernel": top["kernel"],
                           "achieved": top["achieved_gbs"] if hbm_bound else top["achieved_tflops"],
                           "peak": hbm if hbm_bound else top["unit_peak_tflops"], "unit": "GB/s" if hbm_bound else "TFLOP/s",
                           "frac": top["frac"], "traffic": traffic, "traffic_source": traffic_source,
                           "peaks": {"hbm_gbs": hbm, "bf16_tflops": bf16, "tf32_tflops_measured_live": tf32, "fp32_ffma_tflops_measured_live": fp32_peak, "source": how},
                           "note": "dominant = the kernel class with the largest ms_per_step. Per layer, roofline time = max(conv FLOPs / unit peak, (input + output) bytes / "
                                   "HBM peak); frac = sum of roofline times / sum of CUDA-event times over the class's launches of one step, measured live through the "
                                   "per-op C-ABI on the network's 18 layer shapes. Unit peaks: 16-bit 3-term split = bf16 / 3, TF32 3-term split = measured TF32 / 3, "
                                   "SIMT = measured FFMA. For an hbm-bound class achieved / peak are GB/s of algorithmic bytes (frac also counts its compute-bound layers).",
                           "per_kernel": per_kernel}
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        if mode in ("infer", "tile"):
            which, cstep = cpu_infer_stepper(1)
            cstep()
            t0 = time.perf_counter()
            for _ in range(20):
                cstep()
            dt = (time.perf_counter() - t0) / 20
            out["cpu_baseline"] = {"value": PATCH_MPIX / dt, "unit": unit, "cores": cores, "kind": which,
                                   "sample": "20 eval forwards of 1 patch (batch 1, like predict.py), after 1 warm-up"}
        else:
            kind = HYPER[mode][0]
            which, cstep = cpu_train_stepper(mode, B)
            cstep()
            t0 = time.perf_counter()
            for _ in range(5):
                cstep()
            dt = (time.perf_counter() - t0) / 5
            out["cpu_baseline"] = {"value": B / dt, "unit": unit, "cores": cores, "kind": which,
                                   "sample": f"5 full {kind.upper()} train steps of {B} patches (the benchmark's own batch) after 1 warm-up, torch CPU fp32"}
    ident = None
    if world > 1 and mode.startswith("train"):
        # every rank must hold bit-identical weights after the timed steps: all-reduce MIN / MAX of two checksums of the flat parameter buffer
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).double()
        cs = torch.stack([flat.sum(), (flat * flat).sum()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ident = bool(torch.equal(lo, hi))
    extras = {}
    if mode == "train_sr1" and not args.no_extras:
        extras["sr2"] = sr2_subrecord(model_mod, dev, B, world, rank, devb, nb, max(5, args.steps // 2))
        if world == 1:
            extras["inference"] = inference_subrecord(m, dev)
    if rank == 0:
        if ident is not None:
            out["ranks_identical"] = ident
        out.update(extras)
        emit(out)
    if world > 1:
        # teardown: the captured graphs hold persistent references on the NCCL communicator (ncclCommDestroy would wait for them), so release them
        # first; every collective of this run has completed on every rank by now (the checksum / sub-record reductions above were the last ones).
        sys.stdout.flush()
        sys.stderr.flush()
        killer = threading.Timer(30.0, lambda: os._exit(0))   # the result is out: a teardown that hangs must not hang the benchmark
        killer.daemon = True
        killer.start()
        if mode.startswith("train"):
            tr.release_graphs()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train_sr1", choices=["train_sr1", "train_sr2", "infer", "tile"])
    ap.add_argument("--batch", type=int, default=32, help="patches per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=32, help="patches per step of the CPU reference arm (default: the benchmark's own batch)")
    ap.add_argument("--no-graph", action="store_true", help="launch the training step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SR2 and inference sub-records of the default line")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON record): anything a library writes to file descriptor 1 on the way (NCCL prints its version banner
    # there when NCCL_DEBUG=VERSION is set on the box) goes to stderr instead
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
