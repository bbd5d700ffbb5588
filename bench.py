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
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import sifnn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mode = args.mode
    if mode == "infer":
        sd = O.init_state_dict(0)
        bs = 1
        x = torch.randn(bs, 2, 256, 256, generator=torch.Generator().manual_seed(1234))

        def step():
            with torch.inference_mode():
                O.forward(sd, x, train=False)
        per_step, unit, metric = bs * PATCH_MPIX, "Mpix/s", "ModelB inference Mpix/s"
        sample = "eval forward of 1 synthetic patch per step (batch 1, like predict.py)"
    else:
        kind, alpha, gamma, lr = HYPER[mode]
        bs = args.ref_batch
        tr = O.Trainer(O.init_state_dict(0), kind, alpha, gamma, lr)
        lst, up, ndvi = O.synthetic_batch(bs)

        def step():
            tr.step(lst, up, ndvi)
        per_step, unit, metric = bs, "patches/s", f"ModelB {kind.upper()} train patches/s"
        sample = f"{kind.upper()} train step on {bs} of the 32 patches per step (bounded CPU sample)"
    for _ in range(min(args.warmup, 1) if args.warmup else 0):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    out = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(mode, 32), "reference_path": "PyTorch CPU fp32 (oracle port of model.py + loss helpers)"},
           "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def workload_name(mode, batch):
    if mode == "tile":
        return f"ModelB whole-tile inference: LST 1200x1200 + NDVI 4800x4800 -> 4800x4800, 324 windows in batches of {batch}, block-partitioned over the ranks, fp32"
    if mode == "infer":
        return f"ModelB eval forward, batch {batch} synthetic 64x64 LST + 256x256 NDVI -> 256x256, fp32"
    return f"ModelB SIF-NN-{HYPER[mode][0].upper()} training step, batch {batch}/GPU synthetic patches (64->256), fp32"


# --------------------------------------------------------------------------------------------------
# per-kernel roofline (measured live, CUDA events on the launch stream)
# --------------------------------------------------------------------------------------------------
def kernel_rooflines(batch, fp32_peak, tensor_peak_eff):
    """Times every convolution kernel class on the workload's own 18 layer shapes through the per-op C-ABI, using the
    same kernel choice as the network plan (tcgen05 where the shape is eligible, fp32 SIMT elsewhere)."""
    import sifnn_b200
    from sifnn_b200 import ops
    lib = sifnn_b200.load()
    layers = [(2, 16, 256), (16, 16, 256), (16, 16, 128), (16, 16, 128), (16, 32, 128), (32, 32, 64), (32, 32, 64), (32, 64, 64),
              (64, 64, 32), (64, 64, 32), (64, 64, 32), (128, 64, 64), (64, 32, 64), (64, 32, 128), (32, 16, 128), (32, 16, 256),
              (16, 16, 256), (16, 1, 256)]
    dev = "cuda"
    use_tc = sifnn_b200.tensor_cores_enabled()

    def timed(fn, reps=3):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    res = {}

    def add(name, t, fl):
        r = res.setdefault(name, [0.0, 0.0, 0])
        r[0] += t; r[1] += fl; r[2] += 1

    for i, (ci, co, hw) in enumerate(layers):
        x = torch.randn(batch, ci, hw, hw, device=dev)
        dy = torch.randn(batch, co, hw, hw, device=dev)
        w = torch.randn(co, ci, 3, 3, device=dev) * 0.1
        fl = 2.0 * batch * ci * co * 9 * hw * hw
        if use_tc and co <= 64 and lib.sifnn_conv3x3_tc_supported(ci, co, hw, hw):
            add("conv3x3_tc_kernel (fwd)", timed(lambda: ops.conv3x3_fwd_tc(x, w)), fl)
        else:
            add("conv3x3_kernel 2->16 + conv3x3_to1_kernel 16->1 (fwd, SIMT)", timed(lambda: ops.conv3x3_fwd(x, w)), fl)
        if i > 0:
            if use_tc and lib.sifnn_conv3x3_tc_supported(co, ci, hw, hw):
                add("conv3x3_tc_kernel (dgrad)", timed(lambda: ops.conv3x3_dgrad_tc(dy, w)), fl)
            else:
                add("dgrad_from1_kernel 16->1 (dgrad, SIMT)", timed(lambda: ops.conv3x3_dgrad(dy, w)), fl)
        if use_tc and lib.sifnn_conv3x3_wgrad_tc_supported(ci, co, hw, hw):
            add("wgrad_tc_kernel", timed(lambda: ops.conv3x3_wgrad_tc(x, dy)), fl)
        else:
            add("wgrad_kernel (SIMT)", timed(lambda: ops.conv3x3_wgrad(x, dy, want_bias=(co == 1))), fl)
        del x, dy, w
    out = []
    for name, (t, fl, n) in res.items():
        tensor = "tc_kernel" in name
        peak = tensor_peak_eff if tensor else fp32_peak
        out.append({"kernel": name, "bound": "tensor" if tensor else "fp32", "launches_per_step": n, "ms_per_step": t * 1e3,
                    "achieved": fl / t / 1e12, "unit": "TFLOP/s", "peak": peak, "frac": fl / t / 1e12 / peak})
    return out


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

        def step_host(i):
            l, n = tiles_h[i % 2]
            o = sifnn_b200.super_resolve_tile(m, l.to(dev, non_blocking=True), n.to(dev, non_blocking=True), stats, batch=B, rank=rank,
                                              world_size=world, out=out_d)
            out_h.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out_h
        h2d, d2h = (1200 * 1200 + 4800 * 4800) * 4, 4800 * 4800 * 4
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
        tensor_eff = bf16 / 2.0 / 3.0  # TF32 dense = half the measured bf16 peak; fp32 parity costs 3 TF32 products per MAC
        per_kernel = kernel_rooflines(B, fp32_peak, tensor_eff)
        if mode in ("infer", "tile"):
            per_kernel = [k for k in per_kernel if "fwd" in k["kernel"]]
        top = max(per_kernel, key=lambda r: r["ms_per_step"])
        traffic, traffic_source = None, None
        try:  # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
            import csv
            name = {"wgrad_tc_kernel": "r1l_ncu_full_wgrad_tc_32x16x256_summary.csv",
                    "wgrad_kernel (SIMT)": "r1i_ncu_full_wgrad_16x16x256_summary.csv",
                    "conv3x3_tc_kernel (dgrad)": "r1q_ncu_full_conv3x3_tcx_dgrad_16x16x256_summary.csv",
                    "conv3x3_tc_kernel (fwd)": "r1q_ncu_full_conv3x3_tcx_fwd_32x16x256_summary.csv"}.get(top["kernel"], "r1j_ncu_full_conv3x3_tc_64x32x128_summary.csv")
            vals = {r[0]: r[1] for r in csv.reader(open(os.path.join(ROOT, "profiles", name))) if len(r) >= 2}
            traffic = (float(vals["dram__bytes_read.sum"]) + float(vals["dram__bytes_write.sum"])) * 1e6   # bytes of that one launch
            traffic_source = {"launch": name.replace("_summary.csv", ""), "file": "profiles/" + name,
                              "metric": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, one launch of the dominant kernel class"}
        except Exception:
            pass
        out["roofline"] = {"bound": top["bound"], "kernel": top["kernel"], "achieved": top["achieved"], "peak": top["peak"], "unit": "TFLOP/s",
                           "frac": top["frac"], "traffic": traffic, "traffic_source": traffic_source,
                           "note": "achieved = algorithmic conv FLOPs of the kernel class over the network's 18 layer shapes / CUDA-event time, measured "
                                   "live through the per-op C-ABI. fp32 peak = this GPU's measured FFMA throughput (sifnn_fp32_peak_kernel, best of 5; not in "
                                   f"MEASURED_PEAKS.json). tensor peak = {how} bf16 {bf16:.0f} TFLOP/s / 2 (TF32) / 3 (3-term split for fp32 parity) = {tensor_eff:.0f}.",
                           "per_kernel": per_kernel}
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        import sifnn_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        if mode in ("infer", "tile"):
            sd = O.init_state_dict(0)
            x = torch.randn(1, 2, 256, 256)
            with torch.inference_mode():
                O.forward(sd, x)
                t0 = time.perf_counter()
                for _ in range(20):
                    O.forward(sd, x)
                dt = (time.perf_counter() - t0) / 20
            out["cpu_baseline"] = {"value": PATCH_MPIX / dt, "unit": unit, "cores": cores, "kind": "port",
                                   "sample": "20 eval forwards of 1 patch (batch 1, like predict.py), after 1 warm-up"}
        else:
            kind, alpha, gamma, lr = HYPER[mode]
            cb = 16
            ref = O.Trainer(O.init_state_dict(0), kind, alpha, gamma, lr)
            lst, up, ndvi = O.synthetic_batch(cb)
            ref.step(lst, up, ndvi)
            t0 = time.perf_counter()
            for _ in range(2):
                ref.step(lst, up, ndvi)
            dt = (time.perf_counter() - t0) / 2
            out["cpu_baseline"] = {"value": cb / dt, "unit": unit, "cores": cores, "kind": "port",
                                   "sample": f"2 {kind.upper()} train steps of {cb} patches (half a batch) after 1 warm-up, torch CPU fp32"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train_sr1", choices=["train_sr1", "train_sr2", "infer", "tile"])
    ap.add_argument("--batch", type=int, default=32, help="patches per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=8, help="patches per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--no-graph", action="store_true", help="launch the training step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
