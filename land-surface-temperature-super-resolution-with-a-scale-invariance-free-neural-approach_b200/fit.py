"""Epoch loop, early stopping and checkpoint writer around the fused training step.

Host-side mirror of the reference's ``train()`` / ``train_step`` / ``test_step`` epoch bookkeeping
(train_model_B_gradFTM.py:76-137, 181-237, 240-354 -- the SR1 script is identical), of its
``model_checkpoint`` early stopping (utils.py:667-714) and of ``save_model`` / ``load_model``
(utils.py:791-826), with the per-batch arithmetic running through ``Trainer.step`` /
``Trainer.evaluate`` (the hand-written CUDA path).  The artefacts have the reference's names
and formats -- ``<name>_state_dict.pt`` (the 104-key state_dict), ``<name>.pt`` (the pickled
module, class path ``model.ModelB_2``), ``<name>_lossdata.pkl`` (the metrics dict) -- so the
reference's own ``load_model`` / ``read_losses`` / ``predict.py`` read them.

Differences, all deliberate: the three loss scalars are accumulated on the device and read
back once per epoch instead of three ``.item()`` calls per batch; the PSNR / SSIM columns of
the metrics dict (skimage on the host in the reference, after two device->host copies per
batch) come from ``quality``: ``"device"`` = the fused CUDA kernel of row N2
(``ops.psnr_ssim``, accumulated on the device as well), a callable ``(sr, lst_up) -> (psnr,
ssim)``, or ``None`` (columns stay NaN).
"""
from __future__ import annotations

import copy
import os
import pickle
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .trainer import Trainer
from .model import bicubic4_cat
from . import ops

Batch = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]   # (lst, lst_up, ndvi) as ModisDatasetB.__getitem__ collates them

METRIC_KEYS = ("train_loss", "train_dsloss", "train_perceploss", "train_psnr", "train_ssim",
               "val_dsloss", "val_perceploss", "val_loss", "val_psnr", "val_ssim")


class model_checkpoint:
    """Early stopping exactly as the reference's ``us.model_checkpoint`` (utils.py:667-714): keeps a deep copy of the best
    state_dict in host memory, ``train_state`` becomes 'break' when the monitored value has not improved for ``patience``
    epochs, or when the last epoch is reached with a non-zero patience counter."""

    def __init__(self, n_epochs: int, patience: int = 5):
        self.patience = patience
        self.curr_patience = 0
        self.saved_state = None
        self.saved_best_value = None
        self.curr_epoch = None
        self.best_epoch = None
        self.max_epochs = n_epochs
        self.train_state = None

    def test_update(self, model, metrics: Dict[str, List[float]], val_monitored: str, epoch: int) -> None:
        self.curr_epoch = epoch
        value = metrics[val_monitored][-1]
        if epoch == 1:                      # first epoch: take it, train_state stays None (as in the reference)
            self.best_epoch = epoch
            self.saved_state = copy.deepcopy(model.state_dict())
            self.saved_best_value = value
            return
        if value >= self.saved_best_value:  # no improvement (ties count as none)
            self.curr_patience += 1
            if self.curr_patience >= self.patience:
                self.train_state = "break"
            elif self.curr_patience > 0 and epoch == self.max_epochs:
                self.train_state = "break"
            else:
                self.train_state = "continue"
        else:
            self.best_epoch = epoch
            self.curr_patience = 0
            self.saved_best_value = value
            self.saved_state = copy.deepcopy(model.state_dict())
            self.train_state = "continue"


def _quality_fn(quality):
    """None | "device" | callable(sr, lst_up) -> (psnr, ssim) (floats or a 2-element tensor)."""
    if quality is None:
        return None
    if quality == "device":
        return lambda sr, lst_up: ops.psnr_ssim(sr.detach().contiguous(), lst_up.contiguous())   # utils.py:548-578 argument order: (predictions, targets)
    return quality


def _dp_active() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def reduce_epoch_sums(sums: torch.Tensor, qsum: torch.Tensor, n_batches: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Under data parallelism every rank sees its own shard: sum the per-batch loss / quality sums and the batch counts over the ranks
    (one small all-reduce per epoch), so that every rank computes the SAME epoch means -- and therefore takes the same early-stopping
    decision; ranks that disagree on `break` would leave the others hanging in the next gradient all-reduce."""
    if not _dp_active():
        return sums, qsum, n_batches
    import torch.distributed as dist
    t = torch.cat([sums.reshape(-1).double(), qsum.reshape(-1).double(), torch.tensor([float(n_batches)], dtype=torch.float64, device=sums.device)])
    dist.all_reduce(t)
    return t[:sums.numel()].reshape(sums.shape), t[sums.numel():-1].reshape(qsum.shape), int(round(float(t[-1])))


def sync_batchnorm_buffers(model, src: int = 0) -> None:
    """BatchNorm running statistics are per rank (local statistics, like PyTorch DDP, which broadcasts rank 0's buffers before every
    forward): before evaluating / checkpointing make every rank use rank ``src``'s running_mean / running_var / num_batches_tracked."""
    if not _dp_active():
        return
    import torch.distributed as dist
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            for b in (mod.running_mean, mod.running_var, mod.num_batches_tracked):
                dist.broadcast(b, src)


def _epoch_mean(sums: torch.Tensor, n_batches: int) -> Tuple[float, float, float]:
    ds, pl, loss = (sums / max(n_batches, 1)).cpu().tolist()   # the single device -> host read of the epoch
    return ds, pl, loss


def _to_device(dev, lst, lst_up, ndvi):
    """Batches come as tensors or numpy arrays, on the host (pinned: the copies are asynchronous) or already on the device;
    ``lst_up`` may be None (``PinnedBatchLoader(with_upsampled=False)``): the device front-end recomputes it."""
    mv = lambda t: None if t is None else torch.as_tensor(t).to(dev, non_blocking=True)  # noqa: E731
    return mv(lst), mv(lst_up), mv(ndvi)


def train_epoch(trainer: Trainer, batches: Iterable[Batch], device=None,
                quality=None):
    """``train_step`` of the reference: one pass over the loader in train mode; returns the epoch means
    (ds_loss, percep_loss, loss, psnr, ssim) -- psnr / ssim are NaN without a ``quality`` callable."""
    m = trainer.model
    m.train()
    dev = device or next(m.parameters()).device
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    qf = _quality_fn(quality)
    qsum = torch.zeros(2, dtype=torch.float64, device=dev)
    n = 0
    for lst, lst_up, ndvi in batches:
        lst, lst_up, ndvi = _to_device(dev, lst, lst_up, ndvi)
        losses, y = trainer._step_impl(lst, ndvi, lst_up)
        sums += losses
        if qf is not None:
            if lst_up is None:
                lst_up = bicubic4_cat(lst, ndvi)[:, :1]
            qsum += torch.as_tensor(qf(y, lst_up), dtype=torch.float64, device=dev)
        n += 1
    sums, qsum, n = reduce_epoch_sums(sums, qsum, n)
    ds, pl, loss = _epoch_mean(sums, n)
    psnr, ssim = ((qsum / max(n, 1)).cpu().tolist() if qf is not None else (float("nan"), float("nan")))
    return ds, pl, loss, psnr, ssim


@torch.no_grad()
def eval_epoch(trainer: Trainer, batches: Iterable[Batch], device=None,
               quality=None):
    """``test_step`` of the reference: eval-mode forward + the same losses, no update."""
    m = trainer.model
    dev = device or next(m.parameters()).device
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    qf = _quality_fn(quality)
    qsum = torch.zeros(2, dtype=torch.float64, device=dev)
    n = 0
    for lst, lst_up, ndvi in batches:
        lst, lst_up, ndvi = _to_device(dev, lst, lst_up, ndvi)
        sums += trainer.evaluate(lst, ndvi, lst_up)
        if qf is not None:
            if lst_up is None:
                lst_up = bicubic4_cat(lst, ndvi)[:, :1]
            was = m.training
            m.eval()
            y = m(torch.cat((lst_up, ndvi), dim=1))
            m.train(was)
            qsum += torch.as_tensor(qf(y, lst_up), dtype=torch.float64, device=dev)
        n += 1
    sums, qsum, n = reduce_epoch_sums(sums, qsum, n)
    ds, pl, loss = _epoch_mean(sums, n)
    psnr, ssim = ((qsum / max(n, 1)).cpu().tolist() if qf is not None else (float("nan"), float("nan")))
    return ds, pl, loss, psnr, ssim


def fit(trainer: Trainer, train_batches: Callable[[], Iterable[Batch]], val_batches: Callable[[], Iterable[Batch]], n_epochs: int,
        checkpoint: Optional[model_checkpoint] = None, quality=None, on_epoch: Optional[Callable[[int, Dict], None]] = None):
    """The reference's ``train()`` (train_model_B_gradFTM.py:240-354).  ``train_batches`` / ``val_batches`` are callables returning
    a fresh iterable of (lst, lst_up, ndvi) batches per epoch (a DataLoader with shuffle=True is re-iterated the same way).
    Returns (model, metrics) with the reference's metric names; on early stopping the best state_dict is loaded back.

    Data parallel (one process per GPU, each loader yielding its own shard): the epoch means are reduced over the ranks and the BatchNorm
    running buffers of rank 0 are broadcast before every evaluation, so all ranks log the same numbers, make the same early-stopping
    decision and hold the same state_dict when they checkpoint (the weights themselves are identical by construction)."""
    model = trainer.model
    checkpoint = checkpoint or model_checkpoint(n_epochs, patience=5)
    metrics: Dict[str, object] = {k: [] for k in METRIC_KEYS}
    for i in range(1, n_epochs + 1):
        dl, pl, tl, tp, ts = train_epoch(trainer, train_batches(), quality=quality)
        for k, v in zip(("train_dsloss", "train_perceploss", "train_loss", "train_psnr", "train_ssim"), (dl, pl, tl, tp, ts)):
            metrics[k].append(v)
        sync_batchnorm_buffers(model, 0)
        dl, pl, tl, tp, ts = eval_epoch(trainer, val_batches(), quality=quality)
        for k, v in zip(("val_dsloss", "val_perceploss", "val_loss", "val_psnr", "val_ssim"), (dl, pl, tl, tp, ts)):
            metrics[k].append(v)
        checkpoint.test_update(model, metrics, "val_loss", i)
        if on_epoch is not None:
            on_epoch(i, metrics)
        if checkpoint.train_state == "continue" and i == n_epochs:
            metrics["best_epoch"] = n_epochs
        if checkpoint.train_state == "break":
            metrics["best_epoch"] = checkpoint.best_epoch
            model.load_state_dict(checkpoint.saved_state)
            break
    return model, metrics


def save_model(model, path: str, model_name: str) -> Tuple[str, str]:
    """``us.save_model`` (utils.py:802-826): ``<name>_state_dict.pt`` and the pickled module ``<name>.pt``.  Tensors are written
    from host copies so the files load on a CPU-only machine exactly like the shipped ``models/modelB_*`` artefacts."""
    os.makedirs(path, exist_ok=True)
    sd_name = os.path.join(path, model_name + "_state_dict.pt")
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, sd_name)
    md_name = os.path.join(path, model_name + ".pt")
    dev = next(model.parameters()).device
    host = copy.deepcopy(model).to("cpu") if dev.type != "cpu" else model
    torch.save(host, md_name)
    return sd_name, md_name


def load_model(model, state_dict_file: str, device: str = "cpu") -> None:
    """``us.load_model`` (utils.py:791-800)."""
    if device == "cpu":
        model.load_state_dict(torch.load(state_dict_file, map_location=torch.device("cpu")))
    else:
        model.load_state_dict(torch.load(state_dict_file))


def save_metrics(metrics: Dict, path: str, model_name: str) -> str:
    """``<name>_lossdata.pkl`` as written at train_model_B_gradFTM.py:490-491 (read back by ``us.read_losses``)."""
    os.makedirs(path, exist_ok=True)
    name = os.path.join(path, model_name + "_lossdata.pkl")
    with open(name, "wb") as f:
        pickle.dump(metrics, f)
    return name
