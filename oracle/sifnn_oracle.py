"""CPU oracle for the SIF-NN-SR ModelB hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *functional* restatement (plain torch CPU ops, fp32 or
fp64) of the reference algorithm.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` -- never by the product path, which must fail loudly without its CUDA
library.

Parity status: **pinned**.  The reference ships no tests/golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference code
itself, produced in the dev container by ``tests/golden/make_golden.py`` (which
imports ``/root/reference/model.py`` and ``/root/reference/utils.py`` unmodified) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every
function below against those vectors.

Every function cites the reference file:line it restates (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------
# Network description.  model.py:596-605 builds inbloc, db1-3, ub1-3, outlay; the
# state_dict key prefixes below are what those module attribute names generate.
# Each entry: (conv key prefix, bn key prefix or None, C_in, C_out)
# ----------------------------------------------------------------------------------


def conv_table(in_channels: int = 2, down: Sequence[int] = (16, 32, 64, 128)) -> List[Tuple[str, Optional[str], int, int]]:
    """The 18 convolutions of ModelB_2 in execution order (model.py:596-605, 627-643)."""
    d0, d1, d2, d3 = down
    half = d3 // 2  # bilinear=True -> upfactor 2 (model.py:591, 599)
    t: List[Tuple[str, Optional[str], int, int]] = []

    def dconv(prefix: str, cin: int, cout: int, mid: Optional[int] = None) -> None:
        mid = mid or cout  # model.py:129-130
        t.append((f"{prefix}.bloc.0", f"{prefix}.bloc.1", cin, mid))
        t.append((f"{prefix}.bloc.3", f"{prefix}.bloc.4", mid, cout))

    dconv("inbloc", in_channels, d0)
    for name, cin, cout in (("db1", d0, d1), ("db2", d1, d2), ("db3", d2, half)):
        dconv(f"{name}.resblock.doubleconv", cin, cin)
        t.append((f"{name}.lastconv.0", f"{name}.lastconv.1", cin, cout))
    # UpBlock(in, out) with bilinear: DoubleConvolution(in, out, mid=in//2) (model.py:207-208)
    for name, cin, cout in (("ub1", d3, d2 // 2), ("ub2", d2, d1 // 2), ("ub3", d1, d0)):
        dconv(f"{name}.convbloc", cin, cout, cin // 2)
    t.append(("outlay", None, d0, 1))
    return t


def init_state_dict(seed: int = 0, in_channels: int = 2, down: Sequence[int] = (16, 32, 64, 128),
                    dtype: torch.dtype = torch.float32) -> Dict[str, Tensor]:
    """A random state_dict with the reference's key set / shapes.

    Not the reference's default initialiser -- parity tests copy the *same* tensors
    to both sides, so only shapes and key order matter."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for conv, bn, cin, cout in conv_table(in_channels, down):
        bound = 1.0 / math.sqrt(cin * 9)
        sd[f"{conv}.weight"] = ((torch.rand(cout, cin, 3, 3, generator=g, dtype=torch.float64) * 2 - 1) * bound * math.sqrt(3.0)).to(dtype)
        if bn is None:
            sd[f"{conv}.bias"] = ((torch.rand(cout, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        else:
            sd[f"{bn}.weight"] = (1.0 + 0.1 * torch.randn(cout, generator=g, dtype=torch.float64)).to(dtype)
            sd[f"{bn}.bias"] = (0.1 * torch.randn(cout, generator=g, dtype=torch.float64)).to(dtype)
            sd[f"{bn}.running_mean"] = (0.1 * torch.randn(cout, generator=g, dtype=torch.float64)).to(dtype)
            sd[f"{bn}.running_var"] = (1.0 + 0.2 * torch.rand(cout, generator=g, dtype=torch.float64)).to(dtype)
            sd[f"{bn}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return sd


def trainable_keys(sd: Dict[str, Tensor]) -> List[str]:
    """The 53 parameter tensors in ``module.parameters()`` order (registration order)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


# ----------------------------------------------------------------------------------
# Forward pass
# ----------------------------------------------------------------------------------


def _conv3x3_replicate(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    # nn.Conv2d(kernel 3, padding 1, padding_mode='replicate'): model.py:135,138,507,605
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w, b)


def _bn_relu(sd: Dict[str, Tensor], bn: str, x: Tensor, train: bool, new_stats: Optional[Dict[str, Tensor]]) -> Tensor:
    # nn.BatchNorm2d defaults (eps 1e-5, momentum 0.1) then ReLU: model.py:136-137,139-140,508-509
    rm, rv = sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"]
    if train:
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, sd[f"{bn}.weight"], sd[f"{bn}.bias"], True, 0.1, 1e-5)
        if new_stats is not None:
            new_stats[f"{bn}.running_mean"] = rm
            new_stats[f"{bn}.running_var"] = rv
            new_stats[f"{bn}.num_batches_tracked"] = sd[f"{bn}.num_batches_tracked"] + 1
    else:
        y = F.batch_norm(x, rm, rv, sd[f"{bn}.weight"], sd[f"{bn}.bias"], False, 0.1, 1e-5)
    return F.relu(y)


def _cbr(sd, conv: str, bn: str, x: Tensor, train: bool, ns) -> Tensor:
    return _bn_relu(sd, bn, _conv3x3_replicate(x, sd[f"{conv}.weight"]), train, ns)


def _double(sd, prefix: str, x: Tensor, train: bool, ns) -> Tensor:
    # DoubleConvolution.forward: model.py:134-141,159
    x = _cbr(sd, f"{prefix}.bloc.0", f"{prefix}.bloc.1", x, train, ns)
    return _cbr(sd, f"{prefix}.bloc.3", f"{prefix}.bloc.4", x, train, ns)


def _down(sd, name: str, x: Tensor, train: bool, ns) -> Tensor:
    # DownBlock_pool.forward: AvgPool2d(2) -> x + DoubleConv(x) -> Conv.BN.ReLU (model.py:504-509,528-531,311-312)
    x = F.avg_pool2d(x, 2, 2)
    x = x + _double(sd, f"{name}.resblock.doubleconv", x, train, ns)
    return _cbr(sd, f"{name}.lastconv.0", f"{name}.lastconv.1", x, train, ns)


def _up(sd, name: str, x: Tensor, skip: Tensor, train: bool, ns) -> Tensor:
    # UpBlock.forward: bilinear x2 align_corners=True, cat([up, skip]), DoubleConv (model.py:207,235-248)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return _double(sd, f"{name}.convbloc", torch.cat([x, skip], dim=1), train, ns)


def forward(sd: Dict[str, Tensor], x: Tensor, train: bool = False,
            new_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ModelB_2.forward (model.py:608-645).  ``x`` is (B,2,H,W); returns (B,1,H,W).

    ``train=True`` uses batch statistics and, if ``new_stats`` is given, fills it with
    the updated BatchNorm buffers (the state_dict itself is not mutated)."""
    s0 = _double(sd, "inbloc", x, train, new_stats)
    s1 = _down(sd, "db1", s0, train, new_stats)
    s2 = _down(sd, "db2", s1, train, new_stats)
    s3 = _down(sd, "db3", s2, train, new_stats)
    y = _up(sd, "ub1", s3, s2, train, new_stats)
    y = _up(sd, "ub2", y, s1, train, new_stats)
    y = _up(sd, "ub3", y, s0, train, new_stats)
    return _conv3x3_replicate(y, sd["outlay.weight"], sd["outlay.bias"])


# ----------------------------------------------------------------------------------
# Host-side input preparation
# ----------------------------------------------------------------------------------


def bicubic_up4(lst: Tensor) -> Tensor:
    """us.upsampling(img,(4,4)) = cv2.resize(INTER_CUBIC) (utils.py:163-180); OpenCV's
    kernel is Keys a=-0.75 with half-pixel centres and clamped borders, which is what
    torch's bicubic with align_corners=False computes (checked to 5e-7 in make_golden.py)."""
    return F.interpolate(lst, scale_factor=4, mode="bicubic", align_corners=False)


# ----------------------------------------------------------------------------------
# Loss helpers
# ----------------------------------------------------------------------------------


def psf_kernel(mtf: float, factor: float = 4.0, res: float = 1.0) -> Tensor:
    """generate_psf_kernel(res, factor, mtf, None) (utils.py:1615-1639): (2h+1)^2
    Gaussian with sigma = sqrt(-ln(mtf)/2) / (pi * 0.5/factor), normalised, fp32."""
    fc = 0.5 / factor
    sigma = math.sqrt(-math.log(mtf) / 2.0) / (math.pi * fc)
    h = int(math.ceil(factor / res))
    ax = torch.arange(-h, h + 1, dtype=torch.float64) * res
    d2 = ax[:, None] ** 2 + ax[None, :] ** 2
    k = torch.exp(-d2 / (2 * sigma * sigma)) / (sigma * math.sqrt(2 * math.pi))
    return (k / k.sum()).to(torch.float32)


def _psf_blur_padded(x: Tensor, mtf: float) -> Tuple[Tensor, int]:
    # shared head of downscale_LST_SR_to_LR / get_output_ftm (utils.py:1683-1696, 1845-1858):
    # reflect-pad by the half width, then a 'same' (zero padded) depthwise conv.
    k = psf_kernel(mtf).to(device=x.device, dtype=x.dtype)
    hw = (k.shape[-1] - 1) // 2
    c = x.shape[1]
    xp = F.pad(x, (hw, hw, hw, hw), mode="reflect")
    y = F.conv2d(xp, k[None, None].expand(c, -1, -1, -1), groups=c, padding="same")
    return y, hw


def downscale_to_lr(x: Tensor, factor: int = 4, mtf: float = 0.1) -> Tensor:
    """downscale_LST_SR_to_LR(data) 'bic' branch (utils.py:1671-1706)."""
    y, hw = _psf_blur_padded(x, mtf)
    y = F.interpolate(y, scale_factor=1.0 / factor, mode="bicubic")
    s = int(hw / factor)
    return y[:, :, s:y.shape[-2] - s, s:y.shape[-1] - s]


def lowpass(x: Tensor, mtf: float) -> Tensor:
    """get_output_ftm(data, mtf=...) (utils.py:1833-1860)."""
    y, hw = _psf_blur_padded(x, mtf)
    return y[:, :, hw:y.shape[-2] - hw, hw:y.shape[-1] - hw]


SOBEL4 = torch.tensor([
    [[1, 2, 1], [0, 0, 0], [-1, -2, -1]],
    [[1, 0, -1], [2, 0, -2], [1, 0, -1]],
    [[2, 1, 0], [1, 0, -1], [0, -1, -2]],
    [[0, 1, 2], [-1, 0, 1], [-2, -1, 0]],
], dtype=torch.float32)  # train_model_B_predef_filters.py:38-42


def huber_mean(a: Tensor, b: Tensor) -> Tensor:
    # nn.HuberLoss(reduction='mean', delta=1.0): train_model_B_gradFTM.py:454
    return F.huber_loss(a, b, reduction="mean", delta=1.0)


def ds_loss(sr: Tensor, lst: Tensor, mean: float, std: float) -> Tensor:
    """Down-sampling consistency term (train_model_B_gradFTM.py:99-106,
    train_model_B_predef_filters.py:111-118): un-normalise, downscale, re-normalise."""
    down = downscale_to_lr(sr * std + mean)
    return huber_mean((down - mean) / std, lst)


def sr1_losses(sr: Tensor, lst: Tensor, ndvi: Tensor, alpha: float, gamma: float,
               mean: float, std: float) -> Tuple[Tensor, Tensor, Tensor]:
    """SR1 loss (train_model_B_predef_filters.py:111-133): returns (ds, percep, total)."""
    ds = ds_loss(sr, lst, mean, std)
    f = SOBEL4.to(device=sr.device, dtype=sr.dtype)[:, None]
    g_l = F.conv2d(sr, f, padding="same")
    g_n = F.conv2d(ndvi, f, padding="same")
    pl = huber_mean(g_l, gamma * g_n)
    return ds, pl, alpha * ds + (1 - alpha) * pl


def sr2_losses(sr: Tensor, lst: Tensor, ndvi: Tensor, alpha: float, gamma: float,
               mean: float, std: float) -> Tuple[Tensor, Tensor, Tensor]:
    """SR2 / gradFTM loss (train_model_B_gradFTM.py:99-117): returns (ds, percep, total)."""
    ds = ds_loss(sr, lst, mean, std)
    hp_l = sr - lowpass(sr, 0.25)
    hp_n = ndvi - lowpass(ndvi, 0.25)
    pl = huber_mean(hp_l, gamma * hp_n)
    return ds, pl, alpha * ds + (1 - alpha) * pl


LOSSES = {"sr1": sr1_losses, "sr2": sr2_losses}

# Stand-ins for the un-shipped data/statistics.json.
MEAN_LST, STD_LST, MEAN_NDVI, STD_NDVI = 307.24, 5.57, 0.645, 0.168


# ----------------------------------------------------------------------------------
# Training step
# ----------------------------------------------------------------------------------


class Trainer:
    """Body of ``train_step`` (train_model_B_gradFTM.py:89-121 /
    train_model_B_predef_filters.py:101-137) driven with explicit tensors, with
    ``torch.optim.Adam(params, lr)`` exactly as the reference constructs it
    (train_model_B_gradFTM.py:453).  PSNR/SSIM logging is excluded (not part of the
    loss or the gradient)."""

    def __init__(self, sd: Dict[str, Tensor], kind: str, alpha: float, gamma: float, lr: float,
                 mean: float = MEAN_LST, std: float = STD_LST, dtype: torch.dtype = torch.float32):
        self.kind, self.alpha, self.gamma, self.mean, self.std = kind, alpha, gamma, mean, std
        self.sd = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()) for k, v in sd.items()}
        self.keys = trainable_keys(self.sd)
        for k in self.keys:
            self.sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[k] for k in self.keys], lr=lr)
        self.dtype = dtype

    def loss_and_grads(self, lst: Tensor, lst_up: Tensor, ndvi: Tensor, update_stats: bool = True):
        lst, lst_up, ndvi = (t.to(self.dtype) for t in (lst, lst_up, ndvi))
        self.opt.zero_grad(set_to_none=True)
        ns: Dict[str, Tensor] = {}
        sr = forward(self.sd, torch.cat((lst_up, ndvi), dim=1), train=True, new_stats=ns)
        sr.retain_grad()
        ds, pl, loss = LOSSES[self.kind](sr, lst, ndvi, self.alpha, self.gamma, self.mean, self.std)
        loss.backward()
        if update_stats:
            with torch.no_grad():
                for k, v in ns.items():
                    self.sd[k].copy_(v)
        return sr.detach(), (float(ds.detach()), float(pl.detach()), float(loss.detach())), sr.grad.detach()

    def step(self, lst: Tensor, lst_up: Tensor, ndvi: Tensor) -> Tuple[float, float, float]:
        _, scalars, _ = self.loss_and_grads(lst, lst_up, ndvi)
        self.opt.step()
        return scalars

    def flat_grads(self) -> Tensor:
        return torch.cat([self.sd[k].grad.reshape(-1) for k in self.keys])

    def flat_params(self) -> Tensor:
        return torch.cat([self.sd[k].detach().reshape(-1) for k in self.keys])

    def state_dict(self) -> Dict[str, Tensor]:
        return {k: v.detach().clone() for k, v in self.sd.items()}


def synthetic_batch(batch: int, seed: int = 1234, hr: int = 256) -> Tuple[Tensor, Tensor, Tensor]:
    """Seeded z-scored synthetic patches of the paper's shape:
    lst (B,1,hr/4,hr/4), lst_up = bicubic x4, ndvi (B,1,hr,hr)."""
    g = torch.Generator().manual_seed(seed)
    lst = torch.randn(batch, 1, hr // 4, hr // 4, generator=g)
    ndvi = torch.randn(batch, 1, hr, hr, generator=g)
    return lst, bicubic_up4(lst), ndvi


def smooth_batch(batch: int, seed: int = 7, hr: int = 256) -> Tuple[Tensor, Tensor, Tensor]:
    """A smoother, more image-like synthetic distribution (low-pass filtered noise)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch, 1, hr // 16, hr // 16, generator=g)
    ndvi = F.interpolate(base, size=(hr, hr), mode="bicubic", align_corners=False)
    ndvi = ndvi + 0.1 * torch.randn(batch, 1, hr, hr, generator=g)
    lst = F.avg_pool2d(-0.6 * ndvi, 4) + 0.05 * torch.randn(batch, 1, hr // 4, hr // 4, generator=g)
    return lst, bicubic_up4(lst), ndvi
