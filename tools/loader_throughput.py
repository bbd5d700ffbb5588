#!/usr/bin/env python3
"""Throughput of the input pipeline alone: N GeoTiff pairs (64x64 LST + 256x256 NDVI, float32, uncompressed -- what the
reference's save_GeoTiff writes) -> ModisDatasetB -> PinnedBatchLoader, no device work.  usage: loader_throughput.py [workers ...]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import sifnn_b200

N, B = 512, 32
with tempfile.TemporaryDirectory() as d:
    rows = []
    r = np.random.default_rng(0)
    for i in range(N):
        fl, fn = os.path.join(d, f"lst_day_{i}.tif"), os.path.join(d, f"ndvi_{i}.tif")
        sifnn_b200.save_geotiff((r.standard_normal((64, 64)) * 5 + 300).astype(np.float32), fl, "EPSG:32631", (0, 1000, 0, 0, 0, -1000))
        sifnn_b200.save_geotiff(r.random((256, 256)).astype(np.float32), fn, "EPSG:32631", (0, 250, 0, 0, 0, -250))
        rows.append({"LST": fl, "NDVI": fn, "split": "Train"})
    csv = os.path.join(d, "ds.csv"); pd.DataFrame(rows).to_csv(csv)
    stats = os.path.join(d, "statistics.json")
    json.dump({"mean_lst": 307.24, "std_lst": 5.57, "mean_ndvi": 0.645, "std_ndvi": 0.168, "maxi": 340.0}, open(stats, "w"))
    ds = sifnn_b200.ModisDatasetB(csv, stats_path=stats)
    t0 = time.perf_counter()
    for i in range(128):
        ds[i]
    dt = time.perf_counter() - t0
    print(f"__getitem__ (read + z-score + host bicubic), one thread: {128 / dt:8.0f} patches/s")
    for procs in (False, True):
        for up in (True, False):
            for w in [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 4, 8]:
                ld = sifnn_b200.PinnedBatchLoader(ds, B, shuffle=True, seed=0, workers=w, depth=3, with_upsampled=up, processes=procs)
                for _ in ld:
                    pass
                t0 = time.perf_counter()
                n = 0
                for _ in range(3):
                    for lst, _, _ in ld:
                        n += lst.shape[0]
                dt = time.perf_counter() - t0
                ld.close()
                print(f"PinnedBatchLoader B={B} {'processes' if procs else 'threads  '}={w} with_upsampled={up}: {n / dt:8.0f} patches/s")

    if "--train" in sys.argv or os.environ.get("LOADER_TRAIN"):
        import torch
        for procs, w in ((False, 1), (True, 4), (True, 8)):
            torch.manual_seed(0)
            m = sifnn_b200.ModelB_2(2, [16, 32, 64, 128], "replicate", "ReLU", 1, 1).cuda().train()
            t = sifnn_b200.Trainer(m, "sr1", 0.99, -0.5, 1e-3)
            ld = sifnn_b200.PinnedBatchLoader(ds, B, shuffle=True, seed=0, workers=w, depth=4, with_upsampled=False, processes=procs, drop_last=True)
            first = next(iter(ld))
            t.capture(first[0].cuda(), first[2].cuda())
            outs = [torch.empty(3, dtype=torch.float64).pin_memory() for _ in range(4)]
            def epoch():
                k = 0
                for lst, _, ndvi in ld:
                    t.step_host_async(lst, ndvi, outs[k % 4]); k += 1
                return k
            epoch(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            steps = sum(epoch() for _ in range(4))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ld.close()
            print(f"files -> PinnedBatchLoader({'processes' if procs else 'threads'}={w}) -> step_host_async, SR1 B={B}: {steps * B / dt:8.0f} patches/s"
                  f"  (loss {outs[(steps - 1) % 4][2].item():.4f})")
