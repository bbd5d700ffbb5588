"""Input pipeline on the GPU: GeoTiff files -> ModisDatasetB -> PinnedBatchLoader -> Trainer.  The loader's
pinned batches drive ``step_host_async`` and ``fit`` and give the same losses as the same samples fed as device tensors."""
import json
import sys

import numpy as np
import pytest
import torch

import sifnn_b200
from sifnn_b200 import ModelB_2, Trainer, PinnedBatchLoader, ModisDatasetB, fit

pytestmark = pytest.mark.gpu
STATS = {"mean_lst": 307.24, "std_lst": 5.57, "mean_ndvi": 0.645, "std_ndvi": 0.168, "maxi": 340.0}


def _files(tmp_path, n, h):
    import pandas as pd
    rows = []
    for i in range(n):
        r = np.random.default_rng(i)
        lst = (r.standard_normal((h, h)) * 5.57 + 307.24).astype(np.float32)
        ndvi = (r.standard_normal((4 * h, 4 * h)) * 0.168 + 0.645).astype(np.float32)
        fl, fn = tmp_path / f"lst_day_{i}.tif", tmp_path / f"ndvi_{i}.tif"
        sifnn_b200.save_geotiff(lst, fl, "EPSG:32631", (0.0, 1000.0, 0.0, 0.0, 0.0, -1000.0))
        sifnn_b200.save_geotiff(ndvi, fn, "EPSG:32631", (0.0, 250.0, 0.0, 0.0, 0.0, -250.0))
        rows.append({"LST": str(fl), "NDVI": str(fn), "split": "Train" if i < n - 4 else "Val"})
    csv = tmp_path / "ModisDatasetB.csv"
    pd.DataFrame(rows).to_csv(csv)
    stats = tmp_path / "statistics.json"
    stats.write_text(json.dumps(STATS))
    return str(csv), str(stats)


def _model():
    torch.manual_seed(0)
    return ModelB_2(2, [16, 32, 64, 128], "replicate", "ReLU", 1, 1).cuda().train()


@pytest.mark.parametrize("processes", [False, True])
def test_loader_feeds_step_host_async(tmp_path, processes):
    csv, stats = _files(tmp_path, 20, 16)
    d = ModisDatasetB(csv, stats_path=stats)
    assert len(d) == 16
    B = 4
    ld = PinnedBatchLoader(d, B, shuffle=True, seed=3, with_upsampled=False, workers=3, depth=2, processes=processes, chunk=2)
    # A: pinned batches through the asynchronous host path (two epochs: every slot is reused several times)
    ta = Trainer(_model(), "sr1", lr=1e-4)
    first = next(iter(ld))
    assert first[0].is_pinned() and first[2].is_pinned() and first[1] is None
    ta.capture(first[0].cuda(), first[2].cuda())        # capture() leaves weights, Adam state and BatchNorm buffers untouched
    ld.epoch = 0
    host_losses = []
    for _ in range(2):
        for lst, _, ndvi in ld:
            out = torch.empty(3, dtype=torch.float64).pin_memory()
            ta.step_host_async(lst, ndvi, out)
            host_losses.append(out)
    torch.cuda.synchronize()
    # B: the same order from the dataset items, as device tensors
    tb = Trainer(_model(), "sr1", lr=1e-4)
    dev_losses = []
    for ep in range(2):
        g = torch.Generator(); g.manual_seed(3 + ep)
        order = torch.randperm(len(d), generator=g).tolist()
        for i in range(0, len(order), B):
            items = [d[j] for j in order[i:i + B]]
            lst = torch.from_numpy(np.stack([it[0] for it in items])).cuda()
            ndvi = torch.from_numpy(np.stack([it[2] for it in items])).cuda()
            dev_losses.append(tb.step(lst, ndvi).cpu())
    ld.close()
    assert len(host_losses) == len(dev_losses) == 8
    for a, b in zip(host_losses, dev_losses):
        assert torch.allclose(a, b, rtol=2e-4, atol=0), (a, b)


def test_fit_from_files(tmp_path):
    csv, stats = _files(tmp_path, 20, 16)
    tr = ModisDatasetB(csv, stats_path=stats)
    va = ModisDatasetB(csv, split="Val", stats_path=stats)
    # with the host-side bicubic (the reference's triple) and without (device front-end): same epoch metrics
    res = []
    for up in (True, False):
        t = Trainer(_model(), "sr2", lr=1e-4, alpha=0.5, gamma=-0.25)
        ltr = PinnedBatchLoader(tr, 4, shuffle=True, seed=11, with_upsampled=up)
        lva = PinnedBatchLoader(va, 4, shuffle=False, with_upsampled=up)
        _, metrics = fit(t, lambda: ltr, lambda: lva, 2, quality="device")
        res.append(metrics)
    for k in ("train_loss", "val_loss", "train_psnr", "val_ssim"):
        a, b = np.asarray(res[0][k]), np.asarray(res[1][k])
        assert np.all(np.isfinite(a)) and a.shape == (2,)
        np.testing.assert_allclose(a, b, rtol=5e-4)
