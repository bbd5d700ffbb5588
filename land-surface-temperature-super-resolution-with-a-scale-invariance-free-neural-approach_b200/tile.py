"""Whole-tile inference (reference predict.py:81-105): a MODIS tile (LST 1200x1200 K, NDVI 4800x4800) is cut into
64x64 / 256x256 windows, every window goes through ModelB and the de-normalised result is written into the 4x tile.

The reference runs one batch-1 forward per window in a Python double loop; here the window list is gathered on the
device in batches (clip + z-score fused into the gather kernel), pushed through the fused bicubic front-end and the
network, and scattered back (de-normalisation fused).  Windows that are not full 64x64 are skipped and stay 0 --
like the reference (predict.py:95), which leaves the trailing 48-pixel strip of a 1200-pixel tile untouched.
Across ranks the window list is block-partitioned (no collective); each rank fills its own part of the output."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import SifnnError
from .model import ModelB_2, _stream
from .parallel import block_partition


def window_list(ht: int, wt: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Window coordinates in the reference's loop order (predict.py:84-85), full 64x64 windows only."""
    ny, nx = ht // 64, wt // 64
    wy = torch.arange(ny, dtype=torch.int32).repeat_interleave(nx)
    wx = torch.arange(nx, dtype=torch.int32).repeat(ny)
    return wy, wx


@torch.inference_mode()
def super_resolve_tile(model: ModelB_2, lst_tile: torch.Tensor, ndvi_tile: torch.Tensor, stats: Dict[str, float], batch: int = 64,
                       rank: int = 0, world_size: int = 1, out: Optional[torch.Tensor] = None,
                       windows: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """LST_SR tile (4Ht, 4Wt) in Kelvin.  ``stats`` has mean_lst/std_lst/mean_ndvi/std_ndvi like data/statistics.json.
    With world_size > 1 only this rank's block of windows is computed (the rest of ``out`` is left as passed in).
    ``windows`` = explicit (wy, wx) int32 lists instead of every full window of the tile (used by the row-band form below)."""
    if not (lst_tile.is_cuda and ndvi_tile.is_cuda) or lst_tile.dtype != torch.float32 or ndvi_tile.dtype != torch.float32:
        raise SifnnError("super_resolve_tile needs fp32 CUDA tensors")
    ht, wt = lst_tile.shape
    if tuple(ndvi_tile.shape) != (4 * ht, 4 * wt) or wt % 4:
        raise SifnnError(f"expected NDVI (4Ht,4Wt) for LST {(ht, wt)}, got {tuple(ndvi_tile.shape)}")
    if model.training:
        raise SifnnError("super_resolve_tile needs model.eval()")
    lst_tile, ndvi_tile = lst_tile.contiguous(), ndvi_tile.contiguous()
    dev = lst_tile.device
    if out is None:
        out = torch.zeros((4 * ht, 4 * wt), dtype=torch.float32, device=dev)
    if windows is None:
        wy, wx = window_list(ht, wt)
        start, cnt = block_partition(wy.numel(), world_size)[rank]
    else:
        (wy, wx), start = windows, 0
        cnt = wy.numel()
    wy, wx = wy[start:start + cnt].to(dev), wx[start:start + cnt].to(dev)
    ml, sl, mn, sn = (float(stats[k]) for k in ("mean_lst", "std_lst", "mean_ndvi", "std_ndvi"))
    for i in range(0, cnt, batch):
        p = min(batch, cnt - i)
        lst = torch.empty((p, 1, 64, 64), dtype=torch.float32, device=dev)
        ndvi = torch.empty((p, 1, 256, 256), dtype=torch.float32, device=dev)
        _lib.call("sifnn_tile_gather", lst_tile.data_ptr(), ndvi_tile.data_ptr(), wy[i:].data_ptr(), wx[i:].data_ptr(), lst.data_ptr(),
                  ndvi.data_ptr(), p, ht, wt, ml, sl, mn, sn, _stream())
        sr = model.forward_from_lowres(lst, ndvi)
        _lib.call("sifnn_tile_scatter", sr.data_ptr(), wy[i:].data_ptr(), wx[i:].data_ptr(), out.data_ptr(), p, ht, wt, ml, sl, _stream())
    return out


def owned_rows(ht: int, wt: int, rank: int, world_size: int) -> Tuple[int, int, torch.Tensor, torch.Tensor]:
    """This rank's block of the window list (reference order, predict.py:84-85) and the band of window rows it touches:
    (first window row y0, one past the last y1, wy - y0, wx).  An empty share gives y0 == y1."""
    wy, wx = window_list(ht, wt)
    start, cnt = block_partition(wy.numel(), world_size)[rank]
    if cnt == 0:
        return 0, 0, wy[:0], wx[:0]
    wy, wx = wy[start:start + cnt], wx[start:start + cnt]
    y0, y1 = int(wy[0]), int(wy[-1]) + 1
    return y0, y1, wy - y0, wx


@torch.inference_mode()
def super_resolve_tile_host(model: ModelB_2, lst_host: torch.Tensor, ndvi_host: torch.Tensor, stats: Dict[str, float], batch: int = 64,
                            rank: int = 0, world_size: int = 1, out_host: Optional[torch.Tensor] = None,
                            device: Optional[torch.device] = None) -> Tuple[torch.Tensor, int, int]:
    """Host-to-host form for a tile sharded over ranks: only the ROWS this rank's windows touch travel.  LST rows [64 y0, 64 y1) and NDVI
    rows [256 y0, 256 y1) go up; of the result, exactly this rank's windows come back into ``out_host`` (full 256-row bands where the rank
    owns the whole window row, the owned columns of the first / last window row otherwise), so ranks that share one host buffer (or merge
    their buffers) never overwrite each other.  Returns (out_host, first output row, one past the last).  Pinned host tensors make the
    copies asynchronous; the function synchronises the stream before returning.  (Round 1 uploaded the whole tile on every rank.)"""
    if lst_host.is_cuda or ndvi_host.is_cuda:
        raise SifnnError("super_resolve_tile_host takes host tensors (use super_resolve_tile for device tensors)")
    ht, wt = lst_host.shape
    if tuple(ndvi_host.shape) != (4 * ht, 4 * wt) or wt % 4:
        raise SifnnError(f"expected NDVI (4Ht,4Wt) for LST {(ht, wt)}, got {tuple(ndvi_host.shape)}")
    dev = device or next(model.parameters()).device
    if out_host is None:
        out_host = torch.zeros((4 * ht, 4 * wt), dtype=torch.float32)
    y0, y1, wy, wx = owned_rows(ht, wt, rank, world_size)
    if y1 == y0:
        return out_host, 0, 0
    lst_d = lst_host[64 * y0:64 * y1].to(dev, non_blocking=True)
    ndvi_d = ndvi_host[256 * y0:256 * y1].to(dev, non_blocking=True)
    band = torch.zeros((256 * (y1 - y0), 4 * wt), dtype=torch.float32, device=dev)
    super_resolve_tile(model, lst_d, ndvi_d, stats, batch=batch, out=band, windows=(wy, wx))
    # first and last window row of the band may be shared with a neighbouring rank: copy back only this rank's windows there
    nx = wt // 64
    x_first, x_last = int(wx[0]), int(wx[-1]) + 1
    rows = y1 - y0
    for r in range(rows):
        xa = x_first if r == 0 else 0
        xb = x_last if r == rows - 1 else nx
        full = (xa == 0 and xb == nx)
        src = band[256 * r:256 * (r + 1)] if full else band[256 * r:256 * (r + 1), 256 * xa:256 * xb]
        dst = out_host[256 * (y0 + r):256 * (y0 + r + 1)] if full else out_host[256 * (y0 + r):256 * (y0 + r + 1), 256 * xa:256 * xb]
        dst.copy_(src, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return out_host, 256 * y0, 256 * y1


def super_resolve_geotiff(model: ModelB_2, lst_file, ndvi_file, out_file, stats: Dict[str, float], batch: int = 64,
                          device: Optional[torch.device] = None) -> torch.Tensor:
    """File-to-file form of predict.py:70-125 for rasters that are already GeoTiffs: LST in Kelvin at (Ht, Wt), NDVI at (4Ht, 4Wt).
    Reads both with the GDAL-free reader, runs ``super_resolve_tile`` and writes ``out_file`` on the NDVI grid (the reference saves
    its prediction with the CRS and transform of the 250 m product, predict.py:104-125).  Returns the LST_SR tile (device tensor).
    The HDF side of predict.py (us.read_LST / read_NIRRED / compute_NDVI) is outside this library: no HDF4 reader here."""
    from .dataset import read_geotiff, save_geotiff
    lst, _, _, _, _ = read_geotiff(lst_file)
    ndvi, _, _, projection, geotransform = read_geotiff(ndvi_file)
    dev = device or next(model.parameters()).device
    out = super_resolve_tile(model, torch.from_numpy(lst).to(dev), torch.from_numpy(ndvi).to(dev), stats, batch=batch)
    save_geotiff(out.cpu().numpy(), out_file, projection, geotransform)
    return out
