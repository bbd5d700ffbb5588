"""Input side of the training path: the reference's ``ModisDatasetB`` (dataset.py:29-142) without GDAL, and a
batch loader that fills pinned host buffers from worker threads so that ``Trainer.step_host_async`` / ``fit`` never wait for I/O.

* ``read_geotiff`` / ``save_geotiff`` follow ``utils.read_GeoTiff`` / ``utils.save_GeoTiff`` (utils.py:508-543): band 1 as float32
  plus (cols, rows, projection, geotransform).  The decoder is a plain TIFF 6.0 reader (strips or tiles; no compression,
  Deflate or LZW; predictors 1-3) -- enough for what GDAL's GTiff driver writes by default and with COMPRESS=DEFLATE/LZW.
  ``projection`` is "EPSG:<code>" (or the GeoTIFF citation) instead of GDAL's WKT: nothing on the training path reads it.
* ``ModisDatasetB`` keeps the constructor, the csv / split / time filtering, ``statistics.json`` and the three ``transf``
  modes; ``__getitem__`` returns the reference's (lst, lst_up, ndvi) triple.  The bicubic x4 of ``utils.upsampling``
  (cv2.INTER_CUBIC, a = -0.75, replicated border) is restated as two small matrix products.
* ``PinnedBatchLoader`` replaces ``torch.utils.data.DataLoader(dataset, batch_size, shuffle=True)``
  (train_model_B_gradFTM.py:295-296, ``num_workers=0`` there): worker threads decode and normalise straight into a ring of
  pinned batches; ``lst_up`` is optional because the device front-end (``sifnn_bicubic4_cat``) recomputes it from ``lst``.
"""
from __future__ import annotations

import json
import os
import queue
import struct
import threading
import zlib
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import SifnnError

# ---------------------------------------------------------------------------------------------------------------- TIFF

_TYPE = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("I", 4), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2), 9: ("i", 4),
         10: ("i", 4), 11: ("f", 4), 12: ("d", 8), 16: ("Q", 8), 17: ("q", 8)}


def _ifd(buf: bytes, bo: str, off: int) -> Dict[int, tuple]:
    (n,) = struct.unpack_from(bo + "H", buf, off)
    tags: Dict[int, tuple] = {}
    for i in range(n):
        tag, typ, cnt, raw = struct.unpack_from(bo + "HHI4s", buf, off + 2 + 12 * i)
        if typ not in _TYPE:
            continue
        ch, sz = _TYPE[typ]
        per = 2 if typ in (5, 10) else 1
        nbytes = cnt * sz * per
        if nbytes <= 4:
            data = raw[:nbytes]
        else:
            o = struct.unpack(bo + "I", raw)[0]
            data = buf[o:o + nbytes]
        if len(data) < nbytes:
            raise SifnnError(f"read_geotiff: tag {tag} points past the end of the file")
        if typ == 2:
            tags[tag] = (data.split(b"\0")[0].decode("latin-1"),)
        else:
            vals = struct.unpack(bo + ch * (cnt * per), data)
            tags[tag] = tuple(vals[2 * j] / max(vals[2 * j + 1], 1) for j in range(cnt)) if per == 2 else vals
    return tags


def _lzw(data: bytes) -> bytes:
    """TIFF flavour of LZW: MSB-first codes, 9..12 bits, early change, 256 = clear, 257 = end."""
    out = bytearray()
    table: List[bytes] = []
    width, old = 9, -1
    acc = nbits = 0
    it = iter(data)
    while True:
        while nbits < width:
            try:
                acc = (acc << 8) | next(it)
            except StopIteration:
                return bytes(out)
            nbits += 8
        code = (acc >> (nbits - width)) & ((1 << width) - 1)
        nbits -= width
        acc &= (1 << nbits) - 1
        if code == 257:
            return bytes(out)
        if code == 256:
            table = [bytes((i,)) for i in range(256)] + [b"", b""]
            width, old = 9, -1
            continue
        if not table:
            raise SifnnError("read_geotiff: LZW strip does not start with a clear code")
        if old < 0:
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            table.append(table[old] + entry[:1])
        elif code == len(table):
            entry = table[old] + table[old][:1]
            table.append(entry)
        else:
            raise SifnnError("read_geotiff: corrupt LZW stream")
        out += entry
        old = code
        if len(table) >= (1 << width) - 1 and width < 12:
            width += 1


def _decode_chunk(raw: bytes, comp: int, pred: int, rows: int, cols: int, spp: int, dt: np.dtype, bo: str) -> np.ndarray:
    if comp == 1:
        data = raw
    elif comp in (8, 32946):
        data = zlib.decompress(raw)
    elif comp == 5:
        data = _lzw(raw)
    else:
        raise SifnnError(f"read_geotiff: TIFF compression {comp} is not supported (none, Deflate and LZW are)")
    need = rows * cols * spp * dt.itemsize
    if len(data) < need:
        raise SifnnError("read_geotiff: truncated strip / tile")
    data = data[:need]
    if pred == 1:
        return np.frombuffer(data, dtype=dt.newbyteorder(bo)).reshape(rows, cols, spp)
    if pred == 2:      # horizontal differencing on the samples
        a = np.frombuffer(data, dtype=dt.newbyteorder(bo)).reshape(rows, cols, spp)
        if dt.kind == "f":
            raise SifnnError("read_geotiff: predictor 2 on floating-point samples")
        return np.cumsum(a, axis=1, dtype=a.dtype.newbyteorder("="))
    if pred == 3:      # floating-point predictor: byte planes (most significant first), differenced bytewise
        b = np.frombuffer(data, dtype=np.uint8).reshape(rows, cols * spp * dt.itemsize)
        # the differencing runs with a stride of spp bytes over the whole row
        b = b.reshape(rows, -1, spp)
        b = np.cumsum(b, axis=1, dtype=np.uint8).reshape(rows, dt.itemsize, cols * spp)
        be = np.ascontiguousarray(b.transpose(0, 2, 1))
        return be.view(dt.newbyteorder(">")).reshape(rows, cols, spp)
    raise SifnnError(f"read_geotiff: TIFF predictor {pred} is not supported")


def _geokeys(tags: Dict[int, tuple]) -> Dict[int, int]:
    d = tags.get(34735)
    if not d or len(d) < 4:
        return {}
    return {d[4 + 4 * i]: d[7 + 4 * i] for i in range(d[3]) if 4 + 4 * i + 3 < len(d) and d[5 + 4 * i] == 0}


def read_geotiff(file, view_ok: bool = False) -> Tuple[np.ndarray, int, int, str, Tuple[float, ...]]:
    """``utils.read_GeoTiff`` (utils.py:508-525): (band 1 as float32, cols, rows, projection, geotransform).
    ``view_ok`` (loader fast path): when the band is stored as back-to-back uncompressed strips the image is returned as a
    read-only view of the file bytes in the file's sample type, without the float32 copy -- the caller converts while it
    normalises into its own buffer."""
    with open(file, "rb") as fh:
        buf = fh.read()
    if len(buf) < 8 or buf[:2] not in (b"II", b"MM"):
        raise SifnnError(f"read_geotiff: {file} is not a TIFF file")
    bo = "<" if buf[:2] == b"II" else ">"
    magic, off = struct.unpack_from(bo + "HI", buf, 2)
    if magic != 42:
        raise SifnnError(f"read_geotiff: {file}: BigTIFF / unknown magic {magic} is not supported")
    t = _ifd(buf, bo, off)
    one = lambda tag, default=None: (t[tag][0] if tag in t else default)  # noqa: E731
    cols, rows = one(256), one(257)
    if cols is None or rows is None:
        raise SifnnError(f"read_geotiff: {file} has no image size")
    bits, fmt, spp = one(258, 1), one(339, 1), one(277, 1)
    comp, pred, planar = one(259, 1), one(317, 1), one(284, 1)
    kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise SifnnError(f"read_geotiff: {file}: sample format {fmt} with {bits} bits is not supported")
    dt = np.dtype(f"{kind}{bits // 8}")
    chunk_spp = spp if planar == 1 else 1
    img = None if view_ok else np.empty((rows, cols), dtype=np.float32)
    if 322 in t:       # tiles
        if img is None:
            img = np.empty((rows, cols), dtype=np.float32)
        tw, th = one(322), one(323)
        offs, cnts = t[324], t[325]
        across = (cols + tw - 1) // tw
        down = (rows + th - 1) // th
        for ty in range(down):
            for tx in range(across):
                i = ty * across + tx
                a = _decode_chunk(buf[offs[i]:offs[i] + cnts[i]], comp, pred, th, tw, chunk_spp, dt, bo)
                r1, c1 = min(th, rows - ty * th), min(tw, cols - tx * tw)
                img[ty * th:ty * th + r1, tx * tw:tx * tw + c1] = a[:r1, :c1, 0]
    else:
        rps = min(one(278, rows), rows)
        offs = t.get(273)
        if offs is None:
            raise SifnnError(f"read_geotiff: {file} has neither strips nor tiles")
        cnts = t.get(279) or tuple(len(buf) - o for o in offs)
        nstrips = (rows + rps - 1) // rps
        total = rows * cols * chunk_spp * dt.itemsize
        if comp == 1 and pred == 1 and all(offs[i] + cnts[i] == offs[i + 1] for i in range(nstrips - 1)) \
                and offs[0] + total <= len(buf):
            # uncompressed strips laid out back to back (what GDAL and save_geotiff write): one view of the whole band
            band = np.frombuffer(buf, dtype=dt.newbyteorder(bo), count=rows * cols * chunk_spp,
                                 offset=offs[0]).reshape(rows, cols, chunk_spp)[:, :, 0]
            if img is None:
                img = band
            else:
                img[...] = band
            nstrips = 0
        elif img is None:
            img = np.empty((rows, cols), dtype=np.float32)
        for s in range(nstrips):
            r0 = s * rps
            r1 = min(rps, rows - r0)
            img[r0:r0 + r1] = _decode_chunk(buf[offs[s]:offs[s] + cnts[s]], comp, pred, r1, cols, chunk_spp, dt, bo)[:, :, 0]
    keys = _geokeys(t)
    if 34264 in t and len(t[34264]) >= 8:
        m = t[34264]
        gt = (m[3], m[0], m[1], m[7], m[4], m[5])
    elif 33550 in t and 33922 in t and len(t[33922]) >= 6:
        sx, sy = t[33550][0], t[33550][1]
        i, j, _, x, y, _ = t[33922][:6]
        gt = (x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy)
    else:
        gt = (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)    # GDAL's default for a raster without georeferencing
    if keys.get(1025) == 2 and gt[2] == 0.0 and gt[4] == 0.0:   # RasterPixelIsPoint: GDAL reports the corner of the pixel
        gt = (gt[0] - 0.5 * gt[1], gt[1], 0.0, gt[3] - 0.5 * gt[5], 0.0, gt[5])
    code = keys.get(3072) or keys.get(2048)
    projection = f"EPSG:{code}" if code and code != 32767 else (t.get(34737, ("",))[0].split("|")[0])
    return img, int(cols), int(rows), projection, (float(gt[0]), float(gt[1]), float(gt[2]), float(gt[3]), float(gt[4]), float(gt[5]))


def save_geotiff(cropped_img, out_file, projection, geotransform) -> bool:
    """``utils.save_GeoTiff`` (utils.py:528-543): one float32 band, uncompressed strips, pixel scale + tie point from the
    geotransform, and the EPSG code when ``projection`` is "EPSG:<code>"."""
    img = np.ascontiguousarray(np.asarray(cropped_img, dtype="<f4"))
    if img.ndim != 2:
        raise SifnnError("save_geotiff: expected a 2-D image")
    rows, cols = img.shape
    gt = tuple(float(v) for v in geotransform)
    if len(gt) != 6 or gt[2] != 0.0 or gt[4] != 0.0:
        raise SifnnError("save_geotiff: only north-up geotransforms (no rotation terms) are written")
    rps = max(1, min(rows, 8192 // max(cols * 4, 1)))
    nstrips = (rows + rps - 1) // rps
    entries: List[Tuple[int, int, int, bytes]] = []

    def add(tag, typ, vals):
        ch, _ = _TYPE[typ]
        entries.append((tag, typ, len(vals), struct.pack("<" + ch * len(vals), *vals)))

    add(256, 3, [cols]); add(257, 3, [rows]); add(258, 3, [32]); add(259, 3, [1]); add(262, 3, [1])
    add(273, 4, [0] * nstrips)   # patched below
    add(277, 3, [1]); add(278, 3, [rps])
    add(279, 4, [min(rps, rows - s * rps) * cols * 4 for s in range(nstrips)])
    add(284, 3, [1]); add(339, 3, [3])
    add(33550, 12, [gt[1], -gt[5], 0.0])
    add(33922, 12, [0.0, 0.0, 0.0, gt[0], gt[3], 0.0])
    code = int(projection.split(":")[1]) if isinstance(projection, str) and projection.upper().startswith("EPSG:") else 0
    geo = [1, 1, 0, 0]
    keys = [(1025, 0, 1, 1)]
    if code:
        geographic = 4000 <= code < 5000
        keys = [(1024, 0, 1, 2 if geographic else 1)] + keys + [((2048 if geographic else 3072), 0, 1, code)]
    geo[3] = len(keys)
    add(34735, 3, geo + [v for k in sorted(keys) for v in k])
    entries.sort(key=lambda e: e[0])
    head = 8 + 2 + 12 * len(entries) + 4
    extra = bytearray()
    data_off = head + sum((len(p) + 1) // 2 * 2 for _, _, _, p in entries if len(p) > 4)
    strip_offs = [data_off + s * rps * cols * 4 for s in range(nstrips)]
    out = bytearray(b"II" + struct.pack("<HI", 42, 8) + struct.pack("<H", len(entries)))
    for tag, typ, cnt, payload in entries:
        if tag == 273:
            payload = struct.pack("<" + "I" * nstrips, *strip_offs)
        if len(payload) <= 4:
            out += struct.pack("<HHI", tag, typ, cnt) + payload.ljust(4, b"\0")
        else:
            out += struct.pack("<HHII", tag, typ, cnt, head + len(extra))
            extra += payload + (b"\0" if len(payload) % 2 else b"")
    out += struct.pack("<I", 0) + extra
    assert len(out) == data_off
    with open(out_file, "wb") as fh:
        fh.write(out)
        fh.write(img.tobytes())
    return True


# ---------------------------------------------------------------------------------------------------- bicubic x4, host

_A = -0.75


def _cubic_taps(n: int, scale: int) -> Tuple[np.ndarray, np.ndarray]:
    """Source indices (scale*n, 4) and weights (scale*n, 4) of cv2.resize(..., INTER_CUBIC): half-pixel centres, a = -0.75,
    replicated border."""
    d = np.arange(scale * n, dtype=np.float64)
    s = (d + 0.5) / scale - 0.5
    f = np.floor(s)
    t = s - f
    w = np.stack([((_A * (t + 1) - 5 * _A) * (t + 1) + 8 * _A) * (t + 1) - 4 * _A,
                  ((_A + 2) * t - (_A + 3)) * t * t + 1,
                  ((_A + 2) * (1 - t) - (_A + 3)) * (1 - t) * (1 - t) + 1], axis=1)
    w = np.concatenate([w, 1.0 - w.sum(axis=1, keepdims=True)], axis=1).astype(np.float32)
    idx = np.clip(f.astype(np.int64)[:, None] - 1 + np.arange(4)[None, :], 0, n - 1)
    return idx, w


_CUBIC_CACHE: Dict[Tuple[int, int], Tuple[np.ndarray, np.ndarray]] = {}


def upsampling(img: np.ndarray, scale: Sequence[int] = (4, 4)) -> np.ndarray:
    """``utils.upsampling`` (utils.py:163-180): cv2.resize(img, fx, fy, INTER_CUBIC) as two separable 4-tap passes in float32
    (plain numpy gathers: no BLAS call, so worker threads do not fight over a BLAS thread pool)."""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape
    if h != w or int(scale[0]) != int(scale[1]):
        # the reference passes dsize = (rows * scale[0], cols * scale[1]) where cv2 expects (width, height): only square
        # images with one factor (all the path uses) have an unambiguous meaning
        raise SifnnError("upsampling: square images and equal factors only")
    key = (h, int(scale[0]))
    if key not in _CUBIC_CACHE:
        _CUBIC_CACHE[key] = _cubic_taps(*key)
    idx, wt = _CUBIC_CACHE[key]
    rows = img[:, idx[:, 0]] * wt[:, 0]
    for k in range(1, 4):
        rows += img[:, idx[:, k]] * wt[:, k]
    out = rows[idx[:, 0]] * wt[:, 0, None]
    for k in range(1, 4):
        out += rows[idx[:, k]] * wt[:, k, None]
    return out


# ------------------------------------------------------------------------------------------------------------ dataset

class ModisDatasetB(torch.utils.data.Dataset):
    """The reference's training dataset (dataset.py:29-142): csv with columns (index, LST, NDVI, split), one GeoTiff per
    image, statistics from ``./data/statistics.json`` (``stats_path`` overrides the reference's hard-wired location)."""

    def __init__(self, csv_path, transf="norm", split="Train", time="Both", stats_path="./data/statistics.json"):
        import pandas as pd
        self.path = csv_path
        df = pd.read_csv(self.path, sep=",")
        df = df.drop(columns=df.columns[0])
        self.transf = transf
        self.split = split
        df = df.loc[df["split"] == self.split]
        self.pairs = df if time == "Both" else df.loc[df["LST"].str.contains(time)]
        with open(stats_path) as fh:
            self.stats = json.load(fh)
        if transf not in ("-1_1", "0-1", "norm"):
            raise SifnnError(f"ModisDatasetB: unknown transf {transf!r} ('-1_1', '0-1' or 'norm')")
        self._lst_files = self.pairs["LST"].tolist()
        self._ndvi_files = self.pairs["NDVI"].tolist()

    def __len__(self):
        return len(self._lst_files)

    def load_pair(self, idx: int, lst_out: Optional[np.ndarray] = None, ndvi_out: Optional[np.ndarray] = None):
        """The read + transform half of ``__getitem__`` (dataset.py:121-139) -> (lst (h,w), ndvi (4h,4w)); with ``*_out`` the
        result is written into those arrays (pinned batch slices) instead of fresh ones."""
        s = self.stats
        direct = self.transf == "norm" and lst_out is not None and ndvi_out is not None
        lst = read_geotiff(self._lst_files[idx], view_ok=direct)[0]
        ndvi = read_geotiff(self._ndvi_files[idx], view_ok=direct)[0]
        if direct:      # the reference's in-place ``x -= mean; x /= std`` in float32, written straight into the batch slices
            np.subtract(lst, np.float32(s["mean_lst"]), out=lst_out, dtype=np.float32, casting="unsafe")
            np.divide(lst_out, np.float32(s["std_lst"]), out=lst_out)
            np.subtract(ndvi, np.float32(s["mean_ndvi"]), out=ndvi_out, dtype=np.float32, casting="unsafe")
            np.divide(ndvi_out, np.float32(s["std_ndvi"]), out=ndvi_out)
            return lst_out, ndvi_out
        if self.transf == "-1_1":
            lst = lst / np.float32(s["maxi"])
            lst = 2 * (lst - 0.5)
        elif self.transf == "0-1":
            lst = lst / np.float32(s["maxi"])
        else:
            lst -= np.float32(s["mean_lst"]); lst /= np.float32(s["std_lst"])
            ndvi -= np.float32(s["mean_ndvi"]); ndvi /= np.float32(s["std_ndvi"])
        if lst_out is not None:
            lst_out[...] = lst
            lst = lst_out
        if ndvi_out is not None:
            ndvi_out[...] = ndvi
            ndvi = ndvi_out
        return lst, ndvi

    def __getitem__(self, idx):
        lst, ndvi = self.load_pair(idx)
        lst_up = upsampling(lst, (4, 4))
        return np.expand_dims(lst, axis=0), np.expand_dims(lst_up, axis=0), np.expand_dims(ndvi, axis=0)


# ------------------------------------------------------------------------------------------------------- batch loader

def _fill(dataset, arrays, task):
    """Worker body, shared by the thread and the process flavour: decode + normalise ``task``'s samples into slot ``s``."""
    s, j0, idxs = task
    lst_a, ndvi_a, up_a = arrays[s]
    for k, idx in enumerate(idxs):
        lst, _ = dataset.load_pair(idx, lst_a[j0 + k, 0], ndvi_a[j0 + k, 0])
        if up_a is not None:
            up_a[j0 + k, 0][...] = upsampling(lst, (4, 4))


def _worker(dataset, arrays, tasks, done):
    while True:
        item = tasks.get()
        if item is None:
            return
        gen, b, task = item
        try:
            _fill(dataset, arrays, task)
            done.put((gen, b, len(task[2]), None))
        except BaseException as e:  # noqa: BLE001 -- handed to the consumer
            try:
                import pickle
                pickle.loads(pickle.dumps(e))
            except Exception:  # noqa: BLE001
                e = SifnnError(f"{type(e).__name__}: {e}")
            done.put((gen, b, len(task[2]), e))


class PinnedBatchLoader:
    """Iterable of (lst, lst_up, ndvi) batches in pinned host memory -- what ``fit`` / ``train_epoch`` and
    ``Trainer.step_host_async`` consume.  ``workers`` persistent workers fill a ring of ``depth`` batch slots ahead of the
    consumer, ``chunk`` samples per work item.  ``processes=False``: threads (file reads and the big numpy operations
    release the GIL, the TIFF header parsing does not -- good for one or two workers).  ``processes=True``: forked worker
    processes that decode straight into shared-memory batches which the parent registers with CUDA as pinned memory
    (``cudaHostRegister``), so the decode scales with the host cores and the copies out of the slots stay asynchronous; the
    workers never touch CUDA.  ``with_upsampled=False`` yields ``lst_up = None``: the device front-end computes it, which
    saves 4/5 of the host->device bytes of a step and the host bicubic.

    Slot reuse: when the consumer asks for the next batch, an event is recorded on its current CUDA stream for the batch it
    just had, and the slot is handed back to the workers only once that event has completed (polled, never waited for unless no batch is in flight), so asynchronous copies out of
    the slot (``.to(device, non_blocking=True)``, ``step_host_async``) issued before the next ``next()`` are safe.
    Iteration order follows ``torch.randperm`` with ``seed + epoch`` (or the global generator when seed is None), like
    DataLoader(shuffle=True); the last short batch is kept unless ``drop_last``.  ``rank`` / ``world_size`` shard every epoch's
    permutation over the data-parallel ranks like DistributedSampler (same seed on every rank).  ``close()`` stops the workers;
    a consumer that sees no finished work item for ``stall_timeout`` seconds raises instead of waiting forever."""

    def __init__(self, dataset, batch_size: int, shuffle: bool = True, seed: Optional[int] = None, drop_last: bool = False,
                 workers: int = 1, depth: int = 3, with_upsampled: bool = True, pin: Optional[bool] = None,
                 processes: bool = False, chunk: int = 8, rank: int = 0, world_size: int = 1, stall_timeout: float = 300.0):
        if batch_size < 1 or depth < 2 or workers < 1 or chunk < 1:
            raise SifnnError("PinnedBatchLoader: batch_size >= 1, depth >= 2, workers >= 1, chunk >= 1")
        if world_size < 1 or not 0 <= rank < world_size:
            raise SifnnError("PinnedBatchLoader: need 0 <= rank < world_size")
        if world_size > 1 and shuffle and seed is None:
            raise SifnnError("PinnedBatchLoader: data-parallel shuffling needs a seed (every rank must draw the same permutation)")
        self.rank, self.world_size = rank, world_size
        self.stall_timeout = stall_timeout    # seconds without any finished work item before the consumer gives up (a hung worker)
        self.dataset, self.batch_size, self.shuffle, self.seed, self.drop_last = dataset, batch_size, shuffle, seed, drop_last
        self.workers, self.depth, self.with_upsampled = workers, depth, with_upsampled
        self.processes, self.chunk = processes, chunk
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.epoch = 0
        self._gen = 0
        self._slots: Optional[List[dict]] = None
        self._pool: List = []
        self._registered: List[int] = []
        self._active: Optional["_LoaderIter"] = None

    def _shard_len(self):
        return (len(self.dataset) + self.world_size - 1) // self.world_size

    def __len__(self):
        n = self._shard_len()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    # -- buffers and workers, created at the first iteration
    def _alloc(self, *shape):
        if not self.processes:
            return torch.empty(shape, dtype=torch.float32, pin_memory=self.pin)
        return torch.empty(shape, dtype=torch.float32).share_memory_()     # registered as pinned after the workers have forked

    def _start(self):
        lst0, ndvi0 = self.dataset.load_pair(0)
        (h, w), (H, W) = lst0.shape, ndvi0.shape
        B = self.batch_size
        self._slots = [{"lst": self._alloc(B, 1, h, w), "ndvi": self._alloc(B, 1, H, W),
                        "up": self._alloc(B, 1, 4 * h, 4 * w) if self.with_upsampled else None, "event": None}
                       for _ in range(self.depth)]
        arrays = [(sl["lst"].numpy(), sl["ndvi"].numpy(), sl["up"].numpy() if sl["up"] is not None else None) for sl in self._slots]
        if self.processes:
            import multiprocessing as mp
            ctx = mp.get_context("fork")
            self._tasks, self._done = ctx.Queue(), ctx.Queue()
            self._pool = [ctx.Process(target=_worker, args=(self.dataset, arrays, self._tasks, self._done), daemon=True)
                          for _ in range(self.workers)]
        else:
            self._tasks, self._done = queue.Queue(), queue.Queue()
            self._pool = [threading.Thread(target=_worker, args=(self.dataset, arrays, self._tasks, self._done), daemon=True)
                          for _ in range(self.workers)]
        for p in self._pool:
            p.start()
        if self.processes and self.pin:
            # after the fork: the children keep plain mappings of the shared pages, only the parent's are registered with CUDA
            for sl in self._slots:
                for t in (sl["lst"], sl["ndvi"], sl["up"]):
                    if t is not None:
                        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
                        if int(rc) != 0:
                            raise SifnnError(f"PinnedBatchLoader: cudaHostRegister failed with {rc}")
                        self._registered.append(t.data_ptr())

    def close(self):
        """Stop the workers and release the pinned registration; the loader can be iterated again afterwards (it restarts)."""
        if self._active is not None:
            self._active.close()
        if self._pool:
            for _ in self._pool:
                self._tasks.put(None)
            for p in self._pool:
                p.join(timeout=10.0)
                if self.processes and p.is_alive():
                    p.terminate()
            self._pool = []
        if self._registered:
            for ptr in self._registered:
                torch.cuda.cudart().cudaHostUnregister(ptr)
            self._registered = []
        self._slots = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]]:
        if self._active is not None:
            self._active.close()        # one iterator at a time: the slots belong to the loader
        if self._slots is None:
            self._start()
        n = len(self.dataset)
        if self.shuffle:
            g = None
            if self.seed is not None:
                g = torch.Generator()
                g.manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = list(range(n))
        self.epoch += 1
        if self.world_size > 1:
            # torch.utils.data.DistributedSampler semantics: the permutation is padded by wrapping around to a multiple of the
            # world size and dealt out round-robin, so every rank gets the same number of samples (equal step counts)
            pad = self._shard_len() * self.world_size - n
            order = (order + order[:pad])[self.rank::self.world_size] if n else []
            n = len(order)
        batches = [order[i:i + self.batch_size] for i in range(0, n, self.batch_size)]
        if self.drop_last and batches and len(batches[-1]) < self.batch_size:
            batches.pop()
        self._gen += 1
        self._active = _LoaderIter(self, batches, self._gen)
        return self._active


class _LoaderIter:
    def __init__(self, loader: PinnedBatchLoader, batches: List[List[int]], gen: int):
        self.l, self.batches, self.gen = loader, batches, gen
        self.free = list(range(loader.depth))
        self.left: Dict[int, int] = {}      # batch -> samples still to be filled
        self.slot_of: Dict[int, int] = {}
        self.outstanding = 0                # work items handed out and not yet reported back
        self.next_fill = 0                  # next batch to hand to the workers
        self.next_out = 0                   # next batch the consumer gets
        self.held: Optional[int] = None
        self.closed = False
        self._dispatch()

    def _dispatch(self, block: bool = False):
        """Hand free slots to the workers.  A slot whose consumer-side copies may still be running (its event has not completed) is skipped, not
        waited for: the consumer thread never blocks on the GPU here (round 1 synchronised on the slot it had just released, i.e. on the step
        it had just launched).  ``block=True`` -- only when the consumer would otherwise have no batch to wait for -- waits for the oldest one."""
        l = self.l
        while self.next_fill < len(self.batches) and self.free:
            pick = None
            for i, cand in enumerate(self.free):
                ev = l._slots[cand]["event"]
                if ev is None or ev.query():
                    pick = i
                    break
            if pick is None:
                if not block:
                    return
                pick = 0
                l._slots[self.free[0]]["event"].synchronize()
            block = False
            s = self.free.pop(pick)
            l._slots[s]["event"] = None
            b = self.next_fill
            self.next_fill += 1
            self.slot_of[b] = s
            idxs = self.batches[b]
            self.left[b] = len(idxs)
            for j0 in range(0, len(idxs), l.chunk):
                l._tasks.put((self.gen, b, (s, j0, idxs[j0:j0 + l.chunk])))
                self.outstanding += 1

    def _collect_one(self):
        waited = 0.0
        while True:
            try:
                gen, b, n, err = self.l._done.get(timeout=1.0)
            except queue.Empty:
                if not all(p.is_alive() for p in self.l._pool):
                    raise SifnnError("PinnedBatchLoader: a worker died")
                waited += 1.0
                if waited >= self.l.stall_timeout:
                    if self.l.processes:
                        for p in self.l._pool:
                            p.terminate()
                    raise SifnnError(f"PinnedBatchLoader: no work item finished for {self.l.stall_timeout:.0f} s (worker stalled)")
                continue
            if gen != self.gen:
                continue
            self.outstanding -= 1
            if err is not None:
                return err
            self.left[b] -= n
            return None

    def _release_held(self):
        if self.held is None:
            return
        slot = self.l._slots[self.held]
        if self.l.pin and torch.cuda.is_available():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            slot["event"] = ev
        self.free.append(self.held)
        self.held = None

    def __iter__(self):
        return self

    def __next__(self):
        if self.closed:
            raise StopIteration
        self._release_held()
        self._dispatch()
        if self.next_out >= len(self.batches):
            self.close()
            raise StopIteration
        b = self.next_out
        if b not in self.slot_of:        # every free slot is still being read by the GPU: now (and only now) wait for the oldest
            self._dispatch(block=True)
        while self.left[b] > 0:
            err = self._collect_one()
            if err is not None:
                self.close()
                raise err
        self.next_out += 1
        s = self.slot_of.pop(b)
        self.left.pop(b)
        self.held = s
        slot = self.l._slots[s]
        k = len(self.batches[b])
        return slot["lst"][:k], (slot["up"][:k] if slot["up"] is not None else None), slot["ndvi"][:k]

    def close(self):
        """Wait for the work items in flight (the slots are written by the next iterator again) and detach from the loader."""
        if self.closed:
            return
        self.closed = True
        self._release_held()
        try:
            while self.outstanding > 0:
                self._collect_one()
        except SifnnError:
            pass
        if self.l._active is self:
            self.l._active = None
