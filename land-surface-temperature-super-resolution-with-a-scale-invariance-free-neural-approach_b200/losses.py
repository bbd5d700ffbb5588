"""Fused SIF-NN-SR losses (reference train_model_B_predef_filters.py:111-133 for SR1,
train_model_B_gradFTM.py:99-117 for SR2) on top of ``sifnn_loss_fwd_bwd``.

The CUDA kernel evaluates the three scalars the reference's loop logs (``ds_loss``,
``percep_loss``, ``loss``) and dLoss/dSR in one pass.  What stays on the host is the
construction of four tiny per-axis tables (done once per image size, in float64):

* ``g9(mtf)``   -- the 1-D factor of the reference's 9x9 PSF (utils.py:1615-1639; the
  kernel is exactly rank-1);
* ``h12``       -- blur(mtf .1) followed by the stride-4 bicubic decimation taps
  [-3,19,19,-3]/32 that ``interpolate(scale_factor=1/4,'bicubic')`` reduces to;
* ``tab_ds``, ``tab_lp`` -- the adjoints of those reflect-padded operators in gather form.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import SifnnError

KIND = {"sr1": 1, "sr2": 2}


def gauss9(mtf: float, factor: float = 4.0) -> np.ndarray:
    """1-D factor g with outer(g,g) == generate_psf_kernel(1, factor, mtf) (utils.py:1615-1639)."""
    fc = 0.5 / factor
    sigma = math.sqrt(-math.log(mtf) / 2.0) / (math.pi * fc)
    h = int(math.ceil(factor))
    ax = np.arange(-h, h + 1, dtype=np.float64)
    g = np.exp(-ax * ax / (2 * sigma * sigma))
    return g / g.sum()


def _reflect(p: int, n: int) -> int:
    return -p if p < 0 else (2 * (n - 1) - p if p >= n else p)


@lru_cache(maxsize=16)
def _tables_np(n: int) -> Dict[str, np.ndarray]:
    g01, g025 = gauss9(0.1), gauss9(0.25)
    dec = np.array([-3.0, 19.0, 19.0, -3.0]) / 32.0
    # D[I] = sum_a dec[a] * blur[4I + a],  blur[p] = sum_m g[m+4] * u[reflect(p + m)]  =>  h12[t], u index 4I - 4 + t
    h12 = np.zeros(12)
    for a in range(4):
        for m in range(9):
            h12[a + m] += dec[a] * g01[m]
    tab_ds = np.zeros((n, 3))
    for i_lo in range(n // 4):
        for t in range(12):
            r = _reflect(4 * i_lo - 4 + t, n)
            i = i_lo - (r // 4 - 1)
            assert 0 <= i < 3
            tab_ds[r, i] += h12[t]
    tab_lp = np.zeros((n, 9))
    for rp in range(n):
        for m in range(9):
            r = _reflect(rp + m - 4, n)
            j = rp - r + 4
            assert 0 <= j < 9
            tab_lp[r, j] += g025[m]
    return {"h12": h12.astype(np.float32), "tab_ds": tab_ds.astype(np.float32),
            "tab_lp": tab_lp.astype(np.float32), "g9": g025.astype(np.float32)}


_dev_tables: Dict[Tuple[int, str], Dict[str, torch.Tensor]] = {}


def tables(n: int, device) -> Dict[str, torch.Tensor]:
    key = (n, str(device))
    t = _dev_tables.get(key)
    if t is None:
        t = {k: torch.from_numpy(v).to(device).contiguous() for k, v in _tables_np(n).items()}
        _dev_tables[key] = t
    return t


def loss_fwd_bwd(kind: str, sr: torch.Tensor, lst: torch.Tensor, ndvi: torch.Tensor, alpha: float, gamma: float,
                 want_grad: bool = True, losses: torch.Tensor = None, dsr: torch.Tensor = None):
    """Raw launch: returns (losses (3,) float64 device tensor, dsr or None).  No host sync."""
    if kind not in KIND:
        raise SifnnError(f"unknown loss kind {kind!r} (expected 'sr1' or 'sr2')")
    for t in (sr, lst, ndvi):
        if not t.is_cuda or t.dtype != torch.float32:
            raise SifnnError("loss: fp32 CUDA tensors required (no CPU fallback)")
    B, c, H, W = sr.shape
    if c != 1 or tuple(ndvi.shape) != (B, 1, H, W) or tuple(lst.shape) != (B, 1, H // 4, W // 4):
        raise SifnnError(f"loss: shapes sr {tuple(sr.shape)}, ndvi {tuple(ndvi.shape)}, lst {tuple(lst.shape)} do not match (B,1,H,W)/(B,1,H/4,W/4)")
    sr, lst, ndvi = sr.contiguous(), lst.contiguous(), ndvi.contiguous()
    tb = tables(H, sr.device)
    if losses is None:
        losses = torch.zeros(3, dtype=torch.float64, device=sr.device)
    else:
        losses.zero_()
    if want_grad and dsr is None:
        dsr = torch.empty_like(sr)
    k = KIND[kind]
    _lib.call("sifnn_loss_fwd_bwd", k, sr.data_ptr(), ndvi.data_ptr(), lst.data_ptr(), tb["tab_ds"].data_ptr(), tb["h12"].data_ptr(),
              tb["tab_lp"].data_ptr() if k == 2 else None, tb["g9"].data_ptr() if k == 2 else None, float(alpha), float(gamma),
              losses.data_ptr(), dsr.data_ptr() if want_grad else None, B, H, W, torch.cuda.current_stream().cuda_stream)
    return losses, (dsr if want_grad else None)


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, lst, ndvi, kind, alpha, gamma):
        losses, dsr = loss_fwd_bwd(kind, sr.detach(), lst, ndvi, alpha, gamma, want_grad=True)
        ctx.save_for_backward(dsr)
        out = losses.to(torch.float32)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_ds, g_pl, g_loss):
        (dsr,) = ctx.saved_tensors
        # Only the total is meant to be back-propagated (the reference calls loss.backward()).
        return dsr * g_loss, None, None, None, None, None


def sr_losses(kind: str, lst_SR: torch.Tensor, lst: torch.Tensor, ndvi: torch.Tensor, alpha: float, gamma: float):
    """(ds_loss, percep_loss, loss) as 0-dim tensors; ``loss.backward()`` works.

    Drop-in for reference lines train_model_B_gradFTM.py:99-117 (kind 'sr2') and
    train_model_B_predef_filters.py:111-133 (kind 'sr1').  ``ds_loss`` and ``percep_loss``
    are for logging: gradients flow through ``loss`` only."""
    return _LossFn.apply(lst_SR, lst, ndvi, kind, float(alpha), float(gamma))


def sr1_losses(lst_SR, lst, ndvi, alpha, gamma):
    return sr_losses("sr1", lst_SR, lst, ndvi, alpha, gamma)


def sr2_losses(lst_SR, lst, ndvi, alpha, gamma):
    return sr_losses("sr2", lst_SR, lst, ndvi, alpha, gamma)
