"""Whole-tile driver (-m gpu): gather -> fused bicubic front-end -> ModelB -> scatter against the reference's
per-window loop (predict.py:84-103) restated with the oracle; block-partitioned sharding reproduces the full tile."""
import numpy as np
import pytest
import torch

import model as model_mod
import sifnn_b200
import sifnn_oracle as O
from conftest import load_ckpt, rel_err

pytestmark = pytest.mark.gpu
STATS = dict(mean_lst=O.MEAN_LST, std_lst=O.STD_LST, mean_ndvi=O.MEAN_NDVI, std_ndvi=O.STD_NDVI)


def make_tile(ht=176, wt=200, seed=3):
    g = torch.Generator().manual_seed(seed)
    lst = 300 + 8 * torch.rand(ht, wt, generator=g)
    ndvi = 0.6 + 0.6 * torch.randn(4 * ht, 4 * wt, generator=g)  # some values outside [-1, 1]: exercises the clip
    return lst, ndvi


def oracle_tile(sd, lst, ndvi):
    """predict.py:81-103 with the oracle forward: windows of 64, incomplete windows skipped, output starts at 0."""
    out = torch.zeros(ndvi.shape)
    for i in range(0, lst.shape[0], 64):
        for j in range(0, lst.shape[1], 64):
            lb = lst[i:i + 64, j:j + 64]
            if lb.shape != (64, 64):
                continue
            nb = ndvi[4 * i:4 * (i + 64), 4 * j:4 * (j + 64)].clamp(-1, 1)
            l = ((lb - O.MEAN_LST) / O.STD_LST)[None, None]
            n = ((nb - O.MEAN_NDVI) / O.STD_NDVI)[None, None]
            with torch.no_grad():
                sr = O.forward(sd, torch.cat((O.bicubic_up4(l), n), 1))[0, 0]
            out[4 * i:4 * (i + 64), 4 * j:4 * (j + 64)] = sr * O.STD_LST + O.MEAN_LST
    return out


def test_tile_matches_reference_loop_and_shards():
    sd = load_ckpt("1009")
    m = model_mod.ModelB_2(2)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    lst, ndvi = make_tile()
    ref = oracle_tile(sd, lst, ndvi)
    out = sifnn_b200.super_resolve_tile(m, lst.cuda(), ndvi.cuda(), STATS, batch=4)
    assert out.shape == ref.shape
    assert float((out.cpu() - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    assert float(out[:, 4 * 192:].abs().max()) == 0.0 and float(out[4 * 128:].abs().max()) == 0.0  # incomplete windows stay 0
    # 3-way block partition of the 2x3 = 6 windows, every rank writing into the same buffer == the full tile
    acc = torch.zeros_like(out)
    for r in range(3):
        sifnn_b200.super_resolve_tile(m, lst.cuda(), ndvi.cuda(), STATS, batch=8, rank=r, world_size=3, out=acc)
    assert torch.equal(acc, out)
    assert sifnn_b200.window_list(1200, 1200)[0].numel() == 324


def test_pipelined_inference_matches_direct_forward():
    """Three-stream pipelined serving path (serving.py): same bits as the direct forward, results in the caller's pinned buffers."""
    sd = load_ckpt("1009")
    m = model_mod.ModelB_2(2)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(12)
    batches = [(torch.randn(3, 1, 64, 64, generator=g).pin_memory(), torch.randn(3, 1, 256, 256, generator=g).pin_memory()) for _ in range(5)]
    outs = [torch.empty(3, 1, 256, 256).pin_memory() for _ in range(5)]
    pipe = sifnn_b200.PipelinedInference(m)
    for (l, n), o in zip(batches, outs):
        pipe.submit(l, n, o)
    pipe.flush()
    with torch.inference_mode():
        for (l, n), o in zip(batches, outs):
            assert torch.equal(m.forward_from_lowres(l.cuda(), n.cuda()).cpu(), o)
    with pytest.raises(sifnn_b200.SifnnError):
        pipe.submit(batches[0][0].clone(), batches[0][1], outs[0])   # not pinned


def test_tile_from_geotiff_files(tmp_path):
    """File-to-file driver: GeoTiff in, GeoTiff out on the NDVI grid, same pixels as the in-memory tile driver."""
    sd = load_ckpt("1009")
    m = model_mod.ModelB_2(2)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    lst, ndvi = make_tile(ht=64, wt=132, seed=5)
    gt_lst, gt_ndvi = (500000.0, 1000.0, 0.0, 4600000.0, 0.0, -1000.0), (500000.0, 250.0, 0.0, 4600000.0, 0.0, -250.0)
    fl, fn, fo = tmp_path / "lst.tif", tmp_path / "ndvi.tif", tmp_path / "prediction.tif"
    sifnn_b200.save_geotiff(lst.numpy(), fl, "EPSG:32631", gt_lst)
    sifnn_b200.save_geotiff(ndvi.numpy(), fn, "EPSG:32631", gt_ndvi)
    out = sifnn_b200.super_resolve_geotiff(m, fl, fn, fo, STATS, batch=2)
    ref = sifnn_b200.super_resolve_tile(m, lst.cuda(), ndvi.cuda(), STATS, batch=2)
    assert torch.equal(out, ref) and float(out[:, :4 * 128].abs().min()) > 0.0
    img, cols, rows, proj, gt = sifnn_b200.read_geotiff(fo)
    assert (rows, cols) == tuple(ndvi.shape) and proj == "EPSG:32631" and gt == gt_ndvi
    assert np.array_equal(img, ref.cpu().numpy())


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_host_sharded_tile_moves_only_owned_rows_and_matches(world):
    """super_resolve_tile_host over `world` simulated ranks (run one after the other on this GPU, one shared host result buffer) == the
    whole-tile device function, bit for bit; every rank uploads only its band of window rows."""
    m = model_mod.ModelB_2(2)
    m.load_state_dict(load_ckpt("1009"))
    m = m.cuda().eval()
    ht, wt = 64 * 5 + 16, 64 * 4       # 5 x 4 full windows + a partial strip that stays 0 (predict.py:95)
    g = torch.Generator().manual_seed(9)
    lst = 300 + 8 * torch.rand(ht, wt, generator=g)
    ndvi = 0.6 + 0.3 * torch.randn(4 * ht, 4 * wt, generator=g)
    stats = STATS
    ref = sifnn_b200.super_resolve_tile(m, lst.cuda(), ndvi.cuda(), stats, batch=7).cpu()
    out = torch.zeros(4 * ht, 4 * wt)
    covered = 0
    for r in range(world):
        y0, y1, wy, wx = sifnn_b200.owned_rows(ht, wt, r, world)
        covered += wy.numel()
        assert (y1 - y0) <= ((20 + world - 1) // world + 3) // 4 + 1     # a band of window rows, not the tile
        _, r0, r1 = sifnn_b200.super_resolve_tile_host(m, lst, ndvi, stats, batch=7, rank=r, world_size=world, out_host=out)
        assert (r0, r1) == (256 * y0, 256 * y1)
    assert covered == 20
    assert torch.equal(out, ref)
