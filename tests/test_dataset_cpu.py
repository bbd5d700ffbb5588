"""Input pipeline on the CPU: the GeoTiff reader against an independent decoder (OpenCV's libtiff) and against
the reference's own rasters when they are present, the host bicubic against the cv2 golden vectors, ``ModisDatasetB`` against
the reference class (dataset.py:29-142) and the pinned batch loader's ordering / content / error behaviour."""
import json
import os
import struct
import sys
import types
import zlib

import numpy as np
import pytest
import torch

import sifnn_b200
from sifnn_b200 import SifnnError

ds = sys.modules[sifnn_b200.ModisDatasetB.__module__]

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"
STATS = {"mean_lst": 307.24, "std_lst": 5.57, "mean_ndvi": 0.645, "std_ndvi": 0.168, "maxi": 340.0}


def _rand(shape, seed, dtype=np.float32):
    r = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return (r.standard_normal(shape) * 7 + 300).astype(dtype)
    return r.integers(0, np.iinfo(dtype).max, size=shape).astype(dtype)


def test_roundtrip_own_writer(tmp_path):
    img = _rand((70, 53), 0)
    gt = (399960.0, 250.0, 0.0, 4900020.0, 0.0, -250.0)
    f = tmp_path / "a.tif"
    assert ds.save_geotiff(img, f, "EPSG:32631", gt) is True
    out, cols, rows, proj, gt2 = ds.read_geotiff(f)
    assert (cols, rows) == (53, 70) and out.dtype == np.float32
    np.testing.assert_array_equal(out, img)
    assert proj == "EPSG:32631" and gt2 == gt
    # a taller image: several strips
    img = _rand((300, 64), 1)
    ds.save_geotiff(img, f, "", (0.0, 1.0, 0.0, 0.0, 0.0, -1.0))
    np.testing.assert_array_equal(ds.read_geotiff(f)[0], img)


def test_reader_vs_opencv_writer(tmp_path):
    cv2 = pytest.importorskip("cv2")
    for comp in (1, 5, 8, 32946):
        for dtype, shape in ((np.float32, (67, 45)), (np.uint16, (40, 90)), (np.uint8, (33, 31)), (np.float32, (256, 256))):
            img = _rand(shape, comp + shape[0], dtype)
            if dtype == np.float32 and shape == (256, 256):
                img = np.round(img)          # compressible: the LZW table fills and resets, codes reach 12 bits
            f = str(tmp_path / f"c{comp}_{np.dtype(dtype).name}_{shape[0]}.tif")
            if not cv2.imwrite(f, img, [cv2.IMWRITE_TIFF_COMPRESSION, comp]):
                pytest.skip("this OpenCV build cannot write TIFF")
            out, cols, rows, _, gt = ds.read_geotiff(f)
            assert (rows, cols) == shape
            np.testing.assert_array_equal(out, img.astype(np.float32), err_msg=f"compression {comp} {dtype} {shape}")
            assert gt == (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)


def _tiff(tags, payload, bo="<"):
    """A tiny TIFF builder independent of the package's writer: tags = [(tag, type, values)], strips/tiles appended raw."""
    fmt = {3: "H", 4: "I", 12: "d"}
    n = len(tags)
    head = 8 + 2 + 12 * n + 4
    extra = b""
    body = b""
    for tag, typ, vals in sorted(tags):
        p = struct.pack(bo + fmt[typ] * len(vals), *vals)
        if len(p) <= 4:
            body += struct.pack(bo + "HHI", tag, typ, len(vals)) + p.ljust(4, b"\0")
        else:
            body += struct.pack(bo + "HHII", tag, typ, len(vals), head + len(extra))
            extra += p
    return (b"II" if bo == "<" else b"MM") + struct.pack(bo + "HI", 42, 8) + struct.pack(bo + "H", n) + body + struct.pack(bo + "I", 0) + extra + payload


def test_tiled_deflate_big_endian_pixel_is_point(tmp_path):
    img = _rand((40, 50), 5)
    tw = th = 16
    across, down = 4, 3
    tiles = []
    for ty in range(down):
        for tx in range(across):
            t = np.zeros((th, tw), dtype=">f4")
            blk = img[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            t[:blk.shape[0], :blk.shape[1]] = blk
            tiles.append(zlib.compress(t.tobytes()))
    geokeys = [1, 1, 0, 2, 1025, 0, 1, 2, 2048, 0, 1, 4326]
    tags = [(256, 3, [50]), (257, 3, [40]), (258, 3, [32]), (259, 3, [8]), (277, 3, [1]), (339, 3, [3]), (322, 3, [tw]), (323, 3, [th]),
            (324, 4, [0] * 12), (325, 4, [len(t) for t in tiles]), (33550, 12, [0.5, 0.25, 0.0]),
            (33922, 12, [0.0, 0.0, 0.0, 10.0, 50.0, 0.0]), (34735, 3, geokeys)]
    base = len(_tiff(tags, b"", ">"))
    offs, o = [], base
    for t in tiles:
        offs.append(o)
        o += len(t)
    tags[8] = (324, 4, offs)
    f = tmp_path / "t.tif"
    f.write_bytes(_tiff(tags, b"".join(tiles), ">"))
    out, cols, rows, proj, gt = ds.read_geotiff(f)
    np.testing.assert_array_equal(out, img)
    assert proj == "EPSG:4326"
    assert gt == (10.0 - 0.25, 0.5, 0.0, 50.0 + 0.125, 0.0, -0.25)      # PixelIsPoint: shifted by half a pixel


def test_floating_point_predictor(tmp_path):
    """Predictor 3 as GDAL writes it with PREDICTOR=3: per row, the big-endian bytes are split into planes (most significant
    first) and differenced bytewise; encoded here from that definition, Deflate on top, two strips."""
    img = _rand((12, 37), 9)
    rows, cols = img.shape
    be = img.astype(">f4").view(np.uint8).reshape(rows, cols, 4)
    planes = np.ascontiguousarray(be.transpose(0, 2, 1)).reshape(rows, 4 * cols)
    diff = planes.copy()
    diff[:, 1:] = planes[:, 1:] - planes[:, :-1]
    strips = [zlib.compress(diff[:8].tobytes()), zlib.compress(diff[8:].tobytes())]
    tags = [(256, 3, [cols]), (257, 3, [rows]), (258, 3, [32]), (259, 3, [8]), (277, 3, [1]), (339, 3, [3]), (317, 3, [3]),
            (278, 3, [8]), (273, 4, [0, 0]), (279, 4, [len(x) for x in strips])]
    base = len(_tiff(tags, b""))
    tags[8] = (273, 4, [base, base + len(strips[0])])
    f = tmp_path / "p3.tif"
    f.write_bytes(_tiff(tags, b"".join(strips)))
    np.testing.assert_array_equal(ds.read_geotiff(f)[0], img)


def test_reader_errors(tmp_path):
    f = tmp_path / "x.tif"
    f.write_bytes(b"not a tiff at all")
    with pytest.raises(SifnnError, match="not a TIFF"):
        ds.read_geotiff(f)
    tags = [(256, 3, [4]), (257, 3, [4]), (258, 3, [32]), (259, 3, [7]), (339, 3, [3]), (273, 4, [200]), (278, 3, [4]), (279, 4, [64])]
    f.write_bytes(_tiff(tags, b"\0" * 200))
    with pytest.raises(SifnnError, match="compression 7"):
        ds.read_geotiff(f)
    with pytest.raises(SifnnError, match="north-up"):
        ds.save_geotiff(np.zeros((4, 4)), f, "", (0, 1, 0.1, 0, 0, -1))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "test_data_formatted", "data")), reason="reference rasters not on this box")
def test_reader_on_reference_rasters():
    """The GDAL-written ASTER rasters shipped with the reference: same pixels as OpenCV's decoder; sane georeferencing."""
    cv2 = pytest.importorskip("cv2")
    d = os.path.join(REF, "test_data_formatted", "data")
    files = sorted(f for f in os.listdir(d) if f.endswith(".tif"))[:12]
    assert files
    for name in files:
        img, cols, rows, proj, gt = ds.read_geotiff(os.path.join(d, name))
        want = cv2.imread(os.path.join(d, name), cv2.IMREAD_UNCHANGED)
        assert img.shape == (rows, cols) == want.shape[:2]
        np.testing.assert_array_equal(img, want.astype(np.float32))
        assert 200.0 < gt[1] < 260.0 and gt[5] == -gt[1] and gt[2] == gt[4] == 0.0      # "aster_250m" (231.65 m grid), north-up
        assert proj.startswith("EPSG:326") or proj.startswith("EPSG:327")        # a UTM zone


def test_host_bicubic_vs_cv2_golden():
    g = np.load(os.path.join(GOLD, "bicubic.npz"))
    for i in range(g["lst"].shape[0]):
        up = ds.upsampling(g["lst"][i, 0], (4, 4))
        assert up.shape == (256, 256) and up.dtype == np.float32
        ref = g["up_cv2"][i, 0]
        assert np.abs(up - ref).max() <= 2e-6 * np.abs(ref).max()
    with pytest.raises(SifnnError):
        ds.upsampling(np.zeros((4, 6), np.float32), (4, 4))


def _make_dataset(tmp_path, n=11, h=16):
    import pandas as pd
    rows = []
    for i in range(n):
        lst = _rand((h, h), 100 + i)
        ndvi = (np.random.default_rng(200 + i).random((4 * h, 4 * h)) * 1.2 - 0.2).astype(np.float32)
        tag = "day" if i % 3 else "night"
        fl, fn = tmp_path / f"lst_{tag}_{i}.tif", tmp_path / f"ndvi_{i}.tif"
        ds.save_geotiff(lst, fl, "EPSG:32631", (0.0, 1000.0, 0.0, 0.0, 0.0, -1000.0))
        ds.save_geotiff(ndvi, fn, "EPSG:32631", (0.0, 250.0, 0.0, 0.0, 0.0, -250.0))
        rows.append({"LST": str(fl), "NDVI": str(fn), "split": "Train" if i < n - 2 else "Val"})
    csv = tmp_path / "ModisDatasetB.csv"
    pd.DataFrame(rows).to_csv(csv)
    stats = tmp_path / "statistics.json"
    stats.write_text(json.dumps(STATS))
    return str(csv), str(stats)


def test_dataset_semantics(tmp_path):
    csv, stats = _make_dataset(tmp_path)
    d = ds.ModisDatasetB(csv, stats_path=stats)
    assert len(d) == 9 and len(ds.ModisDatasetB(csv, split="Val", stats_path=stats)) == 2
    assert len(ds.ModisDatasetB(csv, time="night", stats_path=stats)) == 3
    lst, lst_up, ndvi = d[4]
    assert lst.shape == (1, 16, 16) and lst_up.shape == (1, 64, 64) and ndvi.shape == (1, 64, 64)
    assert all(a.dtype == np.float32 for a in (lst, lst_up, ndvi))
    raw_l = ds.read_geotiff(d.pairs.iloc[4]["LST"])[0]
    raw_n = ds.read_geotiff(d.pairs.iloc[4]["NDVI"])[0]
    np.testing.assert_allclose(lst[0], (raw_l - np.float32(307.24)) / np.float32(5.57), rtol=1e-6)
    np.testing.assert_allclose(ndvi[0], (raw_n - np.float32(0.645)) / np.float32(0.168), rtol=1e-6, atol=1e-7)
    l01 = ds.ModisDatasetB(csv, transf="0-1", stats_path=stats)[4][0]
    np.testing.assert_allclose(l01[0], raw_l / np.float32(340.0), rtol=1e-6)
    l11 = ds.ModisDatasetB(csv, transf="-1_1", stats_path=stats)[4][0]
    np.testing.assert_allclose(l11[0], 2 * (raw_l / np.float32(340.0) - 0.5), rtol=1e-5, atol=1e-6)
    with pytest.raises(SifnnError):
        ds.ModisDatasetB(csv, transf="zscore", stats_path=stats)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "dataset.py")), reason="reference not on this box")
def test_dataset_vs_reference_class(tmp_path, monkeypatch):
    """The reference's ModisDatasetB run here, its two utils calls bound to the GDAL-free reader and to real cv2."""
    cv2 = pytest.importorskip("cv2")
    csv, stats = _make_dataset(tmp_path)
    fake_utils = types.ModuleType("utils")
    fake_utils.read_GeoTiff = lambda f: ds.read_geotiff(f)
    fake_utils.upsampling = lambda img, scale: cv2.resize(img, dsize=(img.shape[0] * scale[0], img.shape[1] * scale[1]),
                                                          fx=scale[0], fy=scale[1], interpolation=cv2.INTER_CUBIC)
    monkeypatch.setitem(sys.modules, "utils", fake_utils)
    import pandas as pd
    orig_drop = pd.DataFrame.drop

    def drop(self, *a, **kw):       # the reference passes columns= together with axis=1, which its pinned pandas accepted
        if "columns" in kw:
            kw.pop("axis", None)
        return orig_drop(self, *a, **kw)
    monkeypatch.setattr(pd.DataFrame, "drop", drop)
    src = open(os.path.join(REF, "dataset.py")).read()
    mod = types.ModuleType("ref_dataset")
    exec(compile(src, "ref_dataset.py", "exec"), mod.__dict__)
    os.makedirs(tmp_path / "data", exist_ok=True)
    (tmp_path / "data" / "statistics.json").write_text(json.dumps(STATS))
    monkeypatch.chdir(tmp_path)                      # the reference opens ./data/statistics.json
    for kw in ({}, {"transf": "0-1"}, {"transf": "-1_1", "split": "Val"}, {"time": "day"}):
        ref = mod.ModisDatasetB(csv, **kw)
        ours = ds.ModisDatasetB(csv, stats_path=stats, **kw)
        assert len(ref) == len(ours) > 0
        for i in range(len(ref)):
            for a, b in zip(ref[i], ours[i]):
                assert a.shape == b.shape and a.dtype == b.dtype
                assert np.abs(a - b).max() <= 2e-6 * max(np.abs(a).max(), 1.0)


def test_loader_order_and_content(tmp_path):
    csv, stats = _make_dataset(tmp_path)
    d = ds.ModisDatasetB(csv, stats_path=stats)
    items = [d[i] for i in range(len(d))]
    ld = ds.PinnedBatchLoader(d, batch_size=4, shuffle=False, workers=3, depth=2, pin=False)
    assert len(ld) == 3
    got = [(a.clone(), b.clone(), c.clone()) for a, b, c in ld]
    assert [g[0].shape[0] for g in got] == [4, 4, 1]
    k = 0
    for lst, up, ndvi in got:
        for j in range(lst.shape[0]):
            np.testing.assert_array_equal(lst[j].numpy(), items[k][0])
            np.testing.assert_array_equal(up[j].numpy(), items[k][1])
            np.testing.assert_array_equal(ndvi[j].numpy(), items[k][2])
            k += 1
    assert k == 9
    assert len(ds.PinnedBatchLoader(d, 4, drop_last=True, pin=False)) == 2
    assert sum(1 for _ in ds.PinnedBatchLoader(d, 4, drop_last=True, shuffle=False, pin=False)) == 2
    # shuffling: a permutation, reproducible per (seed, epoch), different between epochs; lst_up left to the device
    def epoch_ids(loader):
        ids = []
        for lst, up, ndvi in loader:
            assert up is None
            for j in range(lst.shape[0]):
                ids.append(next(i for i, it in enumerate(items) if np.array_equal(it[0], lst[j].numpy())))
        return ids
    a = ds.PinnedBatchLoader(d, 2, shuffle=True, seed=7, with_upsampled=False, pin=False)
    e1, e2 = epoch_ids(a), epoch_ids(a)
    b = ds.PinnedBatchLoader(d, 2, shuffle=True, seed=7, with_upsampled=False, pin=False)
    assert sorted(e1) == sorted(e2) == list(range(9)) and e1 != e2 and epoch_ids(b) == e1
    g = torch.Generator(); g.manual_seed(7)
    assert e1 == torch.randperm(9, generator=g).tolist()


def test_loader_abandoned_iterator_and_errors(tmp_path):
    csv, stats = _make_dataset(tmp_path)
    d = ds.ModisDatasetB(csv, stats_path=stats)
    ld = ds.PinnedBatchLoader(d, 2, shuffle=False, workers=2, depth=2, pin=False)
    it = iter(ld)
    first = next(it)[0].clone()
    it.close()
    again = next(iter(ld))[0]
    np.testing.assert_array_equal(first.numpy(), again.numpy())
    os.remove(d.pairs.iloc[5]["NDVI"])
    with pytest.raises(FileNotFoundError):
        for _ in ds.PinnedBatchLoader(d, 2, shuffle=False, pin=False):
            pass
    with pytest.raises(SifnnError):
        ds.PinnedBatchLoader(d, 0)


def test_loader_process_workers(tmp_path):
    """Forked workers filling shared-memory batches: same batches as the thread flavour, errors reach the consumer, the
    loader restarts after close()."""
    csv, stats = _make_dataset(tmp_path)
    d = ds.ModisDatasetB(csv, stats_path=stats)
    a = ds.PinnedBatchLoader(d, 4, shuffle=True, seed=5, workers=2, depth=2, pin=False, processes=False, chunk=3)
    b = ds.PinnedBatchLoader(d, 4, shuffle=True, seed=5, workers=3, depth=2, pin=False, processes=True, chunk=2)
    try:
        for _ in range(2):
            n = 0
            for (l1, u1, n1), (l2, u2, n2) in zip(a, b):
                assert l2.is_shared()
                assert torch.equal(l1, l2) and torch.equal(u1, u2) and torch.equal(n1, n2)
                n += 1
            assert n == 3
        b.close()
        assert torch.equal(next(iter(b))[0], next(iter(ds.PinnedBatchLoader(d, 4, shuffle=True, seed=7, pin=False)))[0])  # epoch 2 <-> seed 5+2
        os.remove(d.pairs.iloc[2]["LST"])
        with pytest.raises(FileNotFoundError):
            for _ in b:
                pass
    finally:
        a.close()
        b.close()
    assert not b._pool


def test_loader_rank_sharding(tmp_path):
    """Data-parallel sharding: the ranks' epochs partition one common permutation (padded by wrap-around to equal lengths)."""
    csv, stats = _make_dataset(tmp_path)
    d = ds.ModisDatasetB(csv, stats_path=stats)           # 9 samples
    items = [d[i][0] for i in range(len(d))]

    def ids(loader):
        out = []
        for lst, _, _ in loader:
            for j in range(lst.shape[0]):
                out.append(next(i for i, it in enumerate(items) if np.array_equal(it, lst[j].numpy())))
        return out
    world = 2
    per_rank = [ids(ds.PinnedBatchLoader(d, 2, shuffle=True, seed=4, with_upsampled=False, pin=False, rank=r, world_size=world)) for r in range(world)]
    g = torch.Generator(); g.manual_seed(4)
    perm = torch.randperm(9, generator=g).tolist()
    padded = perm + perm[:1]
    assert per_rank[0] == padded[0::2] and per_rank[1] == padded[1::2]
    assert len(per_rank[0]) == len(per_rank[1]) == 5
    assert len(ds.PinnedBatchLoader(d, 2, seed=4, pin=False, rank=1, world_size=2)) == 3
    seq = [ids(ds.PinnedBatchLoader(d, 4, shuffle=False, with_upsampled=False, pin=False, rank=r, world_size=3)) for r in range(3)]
    assert seq == [[0, 3, 6], [1, 4, 7], [2, 5, 8]]
    with pytest.raises(SifnnError, match="seed"):
        ds.PinnedBatchLoader(d, 2, shuffle=True, rank=0, world_size=2)
    with pytest.raises(SifnnError, match="rank"):
        ds.PinnedBatchLoader(d, 2, seed=1, rank=2, world_size=2)


def test_loader_stall_timeout(tmp_path):
    """A worker that never finishes must not hang the consumer."""
    import time
    csv, stats = _make_dataset(tmp_path)
    d = ds.ModisDatasetB(csv, stats_path=stats)

    class Slow:
        def __len__(self):
            return len(d)

        def load_pair(self, idx, lst_out=None, ndvi_out=None):
            if lst_out is not None and idx == 3:
                time.sleep(6.0)
            return d.load_pair(idx, lst_out, ndvi_out)
    ld = ds.PinnedBatchLoader(Slow(), 4, shuffle=False, workers=1, pin=False, with_upsampled=False, stall_timeout=2.0)
    t0 = time.time()
    with pytest.raises(SifnnError, match="stalled"):
        for _ in ld:
            pass
    assert time.time() - t0 < 5.5
