"""Pipelined batched inference from host memory (the serving shape of predict.py:86-103): while batch i runs through the network, batch
i+1 is already crossing PCIe on a copy stream and the result of batch i-1 is on its way back on a third one.  Inputs and outputs are the
caller's PINNED host tensors; two sets of device buffers alternate, ordered by CUDA events only (no host synchronisation until flush())."""
from __future__ import annotations

from typing import List, Optional

import torch

from ._lib import SifnnError
from .model import ModelB_2


class PipelinedInference:
    def __init__(self, model: ModelB_2, depth: int = 2):
        if model.training:
            raise SifnnError("PipelinedInference needs model.eval()")
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise SifnnError("PipelinedInference needs a CUDA model (no CPU fallback)")
        self.depth = depth
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self._bufs: List[Optional[dict]] = [None] * depth
        self._n = 0
        self._checked = set()

    def _slot(self, lst: torch.Tensor, ndvi: torch.Tensor) -> dict:
        i = self._n % self.depth
        b = self._bufs[i]
        if b is None or b["lst"].shape != lst.shape or b["ndvi"].shape != ndvi.shape:
            b = {"lst": torch.empty(lst.shape, dtype=torch.float32, device=self.dev),
                 "ndvi": torch.empty(ndvi.shape, dtype=torch.float32, device=self.dev),
                 "in_done": torch.cuda.Event(), "compute_done": torch.cuda.Event(), "out_done": torch.cuda.Event(), "used": False, "y": None}
            self._bufs[i] = b
        return b

    @torch.inference_mode()
    def submit(self, lst_host: torch.Tensor, ndvi_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Queue one batch: lst (B,1,h,w), ndvi (B,1,4h,4w) z-scored fp32 in pinned host memory; the (B,1,4h,4w) result lands in out_host
        (pinned) once flush() -- or a later submit() that reuses the slot -- has returned."""
        for t in (lst_host, ndvi_host, out_host):
            key = (t.data_ptr(), t.numel())
            if key not in self._checked:   # is_pinned() asks the driver: tens of microseconds per call, so each buffer is vetted once
                if t.is_cuda or not t.is_pinned() or t.dtype != torch.float32 or not t.is_contiguous():
                    raise SifnnError("PipelinedInference.submit takes contiguous pinned fp32 host tensors")
                self._checked.add(key)
        b = self._slot(lst_host, ndvi_host)
        cur = torch.cuda.current_stream(self.dev)
        if b["used"]:
            self.s_in.wait_event(b["compute_done"])      # the device input buffers of this slot are free once its last forward has run
        with torch.cuda.stream(self.s_in):
            b["lst"].copy_(lst_host, non_blocking=True)
            b["ndvi"].copy_(ndvi_host, non_blocking=True)
            b["in_done"].record(self.s_in)
        cur.wait_event(b["in_done"])
        y = self.model.forward_from_lowres(b["lst"], b["ndvi"])
        b["compute_done"].record(cur)
        self.s_out.wait_event(b["compute_done"])
        y.record_stream(self.s_out)
        with torch.cuda.stream(self.s_out):
            out_host.copy_(y, non_blocking=True)
            b["out_done"].record(self.s_out)
        b["y"], b["used"] = y, True
        self._n += 1

    def flush(self) -> None:
        """Block until every submitted batch has been written to its host buffer."""
        for b in self._bufs:
            if b is not None and b["used"]:
                b["out_done"].synchronize()
