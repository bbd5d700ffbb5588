"""Fused training step: the body of the reference's ``train_step`` loop
(train_model_B_gradFTM.py:89-121 for SR2, train_model_B_predef_filters.py:101-137 for
SR1) as five native launches plans -- input stage, forward, loss (+dLoss/dSR), backward,
Adam -- with no host synchronisation and no autograd bookkeeping.

Data parallelism (one process per GPU, ``torch.distributed`` / NCCL): the batch is
sharded across ranks; BatchNorm statistics are local to a rank (PyTorch-DDP semantics:
same result as the reference run on that rank's shard); the flat 282 705-float gradient
buffer is all-reduced in two buckets -- the decoder half as soon as the decoder
backward has been queued, overlapping the encoder backward, then the encoder half --
and every rank applies the identical Adam update with grad_scale = 1/world_size.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import SifnnError
from .losses import loss_fwd_bwd
from .model import ModelB_2, bicubic4_cat, _stream
from .parallel import BucketedAllReduce


class Trainer:
    def __init__(self, model: ModelB_2, kind: str = "sr2", alpha: float = 0.5, gamma: float = -0.25, lr: float = 1e-4,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, data_parallel: Optional[bool] = None,
                 overlap_allreduce: bool = True):
        if kind not in ("sr1", "sr2"):
            raise SifnnError("kind must be 'sr1' or 'sr2'")
        self.model, self.kind, self.alpha, self.gamma = model, kind, float(alpha), float(gamma)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if data_parallel else 1
        self.overlap = overlap_allreduce
        self._opt = None
        self._graph = None

    # ------------------------------------------------------------------ state
    def _opt_state(self, device):
        st = self.model._ensure_flat(device)
        if self._opt is None or self._opt["m"].device != device or self._opt["m"].numel() != st["n"]:
            self._opt = {"m": torch.zeros(st["n"], dtype=torch.float32, device=device),
                         "v": torch.zeros(st["n"], dtype=torch.float32, device=device),
                         "t": torch.zeros((), dtype=torch.int64, device=device)}
        return st, self._opt

    def broadcast_parameters(self, src: int = 0) -> None:
        """Make every rank start from rank ``src``'s weights and BatchNorm buffers."""
        if self.world == 1:
            return
        dev = next(self.model.parameters()).device
        st = self.model._ensure_flat(dev)
        for t in (st["flat"], st["rm"], st["rv"]):
            dist.broadcast(t, src)

    # ------------------------------------------------------------------ one step
    def _backward_and_update(self, x, dsr, ws):
        m = self.model
        st, opt = self._opt_state(x.device)
        fgrad, dec = st["fgrad"], st["dec_off"]
        if self.world > 1 and self.overlap:
            ar = BucketedAllReduce(fgrad, dec)
            m._run_backward(x, dsr, ws, phase=1)
            ar.start(0)                                            # decoder bucket, overlaps the encoder backward
            m._run_backward(x, dsr, ws, phase=2)
            ar.start(1)
            ar.finish()
        else:
            m._run_backward(x, dsr, ws, phase=0)
            if self.world > 1:
                dist.all_reduce(fgrad)
        _lib.call("sifnn_adam_step", st["flat"].data_ptr(), fgrad.data_ptr(), opt["m"].data_ptr(), opt["v"].data_ptr(),
                  opt["t"].data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world, st["n"], _stream())

    def _step_impl(self, lst, ndvi, lst_up):
        m = self.model
        if not m.training:
            raise SifnnError("Trainer.step needs model.train()")
        x = bicubic4_cat(lst, ndvi) if lst_up is None else torch.cat((lst_up, ndvi), dim=1)
        x = m._check_input(x)
        y, ws, key = m._run_forward(x, train=True, keep=True)
        losses, dsr = loss_fwd_bwd(self.kind, y, lst, ndvi, self.alpha, self.gamma, want_grad=True)
        self._backward_and_update(x, dsr, ws)
        m._ws.give(key, ws)
        return losses, y

    def step(self, lst: torch.Tensor, ndvi: torch.Tensor, lst_up: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on this rank's shard.  lst (B,1,h,w), ndvi (B,1,4h,4w) fp32 CUDA
        (already z-scored, like the reference Dataset delivers them); ``lst_up`` optional -- when
        omitted the bicubic x4 is done on the device.  Returns a (3,) float64 device tensor
        (ds_loss, percep_loss, loss) of the local shard WITHOUT synchronising."""
        return self._step_impl(lst, ndvi, lst_up)[0]

    def step_host(self, lst_pinned: torch.Tensor, ndvi_pinned: torch.Tensor) -> Tuple[float, float, float]:
        """End-to-end step from (pinned) HOST buffers: H2D copies, step, D2H of the three scalars."""
        dev = next(self.model.parameters()).device
        if self._graph is not None and tuple(lst_pinned.shape) == tuple(self._static_lst.shape):
            losses = self.step_graph(lst_pinned, ndvi_pinned)      # H2D straight into the graph's static buffers
        else:
            lst = lst_pinned.to(dev, non_blocking=True)
            ndvi = ndvi_pinned.to(dev, non_blocking=True)
            losses = self.step(lst, ndvi)
        ds, pl, loss = losses.cpu().tolist()
        return ds, pl, loss

    def step_host_async(self, lst_pinned: torch.Tensor, ndvi_pinned: torch.Tensor, losses_pinned: torch.Tensor) -> None:
        """End-to-end step from pinned HOST buffers without a host synchronisation: the H2D copy of this step's inputs runs on a copy
        stream (so it overlaps the previous step, which is still executing when the host gets here), the step replays the captured
        graphs, and the three loss scalars are copied into ``losses_pinned`` ((3,) float64, pinned) asynchronously -- valid after the
        caller's next stream / device synchronisation.  Needs capture() first."""
        if self._graph is None:
            raise SifnnError("call capture() first")
        for t in (lst_pinned, ndvi_pinned, losses_pinned):
            if t.is_cuda or not t.is_pinned():
                raise SifnnError("step_host_async takes pinned host tensors")
        dev = self._static_lst.device
        hp = getattr(self, "_host_pipe", None)
        if hp is None:
            hp = self._host_pipe = {"stream": torch.cuda.Stream(dev), "n": 0,
                                    "slots": [{"lst": torch.empty_like(self._static_lst), "ndvi": torch.empty_like(self._static_ndvi),
                                               "in_done": torch.cuda.Event(), "consumed": torch.cuda.Event(), "used": False} for _ in range(2)]}
        b = hp["slots"][hp["n"] % 2]
        cur = torch.cuda.current_stream(dev)
        if b["used"]:
            hp["stream"].wait_event(b["consumed"])
        with torch.cuda.stream(hp["stream"]):
            b["lst"].copy_(lst_pinned, non_blocking=True)
            b["ndvi"].copy_(ndvi_pinned, non_blocking=True)
            b["in_done"].record(hp["stream"])
        cur.wait_event(b["in_done"])
        self._static_lst.copy_(b["lst"], non_blocking=True)
        self._static_ndvi.copy_(b["ndvi"], non_blocking=True)
        b["consumed"].record(cur)
        b["used"] = True
        hp["n"] += 1
        self._replay_or_run(self._segs, eager=False, fgrad=self._ar[0], dec=self._ar[1])
        losses_pinned.copy_(self._static_losses, non_blocking=True)

    @torch.no_grad()
    def evaluate(self, lst: torch.Tensor, ndvi: torch.Tensor, lst_up: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The reference's ``test_step`` body (train_model_B_gradFTM.py:181-215): eval-mode forward + losses."""
        m = self.model
        was = m.training
        m.eval()
        try:
            x = bicubic4_cat(lst, ndvi) if lst_up is None else torch.cat((lst_up, ndvi), dim=1)
            y = m(x)
            losses, _ = loss_fwd_bwd(self.kind, y, lst, ndvi, self.alpha, self.gamma, want_grad=False)
        finally:
            m.train(was)
        return losses

    # ------------------------------------------------------------------ CUDA graph replay
    def capture(self, lst: torch.Tensor, ndvi: torch.Tensor, single_graph: Optional[bool] = None) -> None:
        """Capture the step for inputs shaped like (lst, ndvi); afterwards ``step_graph`` copies new data into the
        static buffers and replays.  ONE graph for the whole step, also under data parallelism: the two NCCL bucket
        all-reduces are captured with the kernels (the decoder bucket on NCCL's stream, overlapping the encoder
        backward), so the host issues one replay per step.  ``single_graph=False`` (or SIFNN_DP_GRAPHS=3, or a failed
        capture of the collectives) falls back to three graphs around two eager all-reduces.

        Capturing does NOT change the training state: the two eager warm-up steps it needs (lazy CUDA attributes,
        allocator, NCCL channels) run on a snapshot of the weights, Adam moments, step counter, BatchNorm running
        statistics and ``num_batches_tracked``, which is restored before returning."""
        m = self.model
        if not m.training:
            raise SifnnError("Trainer.capture needs model.train()")
        dev = lst.device
        st, opt = self._opt_state(dev)
        B, _, h, w = lst.shape
        H, W = 4 * h, 4 * w
        sb = self._static = {
            "lst": lst.clone(), "ndvi": ndvi.clone(),
            "x": torch.empty((B, 2, H, W), dtype=torch.float32, device=dev),
            "y": torch.empty((B, 1, H, W), dtype=torch.float32, device=dev),
            "dsr": torch.empty((B, 1, H, W), dtype=torch.float32, device=dev),
            "losses": torch.zeros(3, dtype=torch.float64, device=dev),
            "ws": torch.empty(m._workspace_bytes(B, H, W, True), dtype=torch.uint8, device=dev),
        }
        self._static_lst, self._static_ndvi, self._static_losses = sb["lst"], sb["ndvi"], sb["losses"]
        fgrad, dec = st["fgrad"], st["dec_off"]
        lib = _lib.load()

        def seg_front():
            bicubic4_cat(sb["lst"], sb["ndvi"], out=sb["x"])
            m._run_forward(sb["x"], train=True, keep=True, ws=sb["ws"], y=sb["y"])
            loss_fwd_bwd(self.kind, sb["y"], sb["lst"], sb["ndvi"], self.alpha, self.gamma, want_grad=True, losses=sb["losses"], dsr=sb["dsr"])
            m._run_backward(sb["x"], sb["dsr"], sb["ws"], phase=0 if self.world == 1 else 1)

        def seg_encoder():
            m._run_backward(sb["x"], sb["dsr"], sb["ws"], phase=2)

        def seg_adam():
            _lib.call("sifnn_adam_step", st["flat"].data_ptr(), fgrad.data_ptr(), opt["m"].data_ptr(), opt["v"].data_ptr(),
                      opt["t"].data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world, st["n"], _stream())

        segs = [seg_front, seg_adam] if self.world == 1 else [seg_front, seg_encoder, seg_adam]
        if single_graph is None:
            import os
            single_graph = os.environ.get("SIFNN_DP_GRAPHS", "1") != "3"
        # warm-up outside capture (lazy attribute setup, cudaFuncSetAttribute, allocator, NCCL channels) -- two eager steps on a snapshot of the state
        snap = [t.clone() for t in (st["flat"], st["rm"], st["rv"], opt["m"], opt["v"], opt["t"])] + [c.clone() for c in st["counters"]]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._replay_or_run(segs, eager=True, fgrad=fgrad, dec=dec)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def restore():
            for t, v in zip((st["flat"], st["rm"], st["rv"], opt["m"], opt["v"], opt["t"]), snap):
                t.copy_(v)
            for c, v in zip(st["counters"], snap[6:]):
                c.copy_(v)

        n0 = lib.sifnn_launch_count()
        graphs = None
        if self.world > 1 and single_graph:
            try:   # everything, collectives included, in one graph
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ar = BucketedAllReduce(fgrad, dec)
                    seg_front()
                    ar.start(0)      # decoder bucket: NCCL's stream forks off the capture stream here and overlaps the encoder backward
                    seg_encoder()
                    ar.start(1)
                    ar.finish()
                    seg_adam()
                graphs, self._dp_single = [g], True
            except Exception as e:   # this NCCL / driver pair cannot capture the collective: three graphs around eager all-reduces
                import warnings
                warnings.warn(f"sifnn Trainer.capture: single-graph data-parallel capture failed ({e!r}); using three graphs")
                torch.cuda.synchronize()
                n0 = lib.sifnn_launch_count()
                graphs = None
        if graphs is None:
            graphs, self._dp_single = [], False
            for fn in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                graphs.append(g)
        self.graph_launches = int(lib.sifnn_launch_count() - n0)  # kernels of this library inside one replayed step
        self._graph = graphs
        self._segs = segs
        self._ar = (fgrad, dec)
        restore()

    def release_graphs(self) -> None:
        """Drop the captured CUDA graphs (and the host-copy pipeline).  REQUIRED before ``torch.distributed.destroy_process_group()`` when the
        data-parallel step was captured as one graph: NCCL counts every captured collective as a persistent reference on its communicator and
        ``ncclCommDestroy`` waits until the graphs that hold them are gone -- a live graph makes the teardown hang."""
        import gc
        torch.cuda.synchronize()
        self._graph = None
        self._segs = None
        self._dp_single = False
        self._host_pipe = None
        gc.collect()
        torch.cuda.synchronize()

    def _replay_or_run(self, segs, eager: bool, fgrad, dec):
        if not eager and getattr(self, "_dp_single", False):
            self._graph[0].replay()   # data parallel, collectives captured: one replay per step
            return
        run = (lambda i: segs[i]()) if eager else (lambda i: self._graph[i].replay())
        if self.world == 1:
            run(0)
            run(1)
            return
        ar = BucketedAllReduce(fgrad, dec)
        run(0)
        ar.start(0)      # decoder bucket overlaps the encoder backward
        run(1)
        ar.start(1)
        ar.finish()
        run(2)

    def step_graph(self, lst: torch.Tensor, ndvi: torch.Tensor) -> torch.Tensor:
        if self._graph is None:
            raise SifnnError("call capture() first")
        self._static_lst.copy_(lst, non_blocking=True)
        self._static_ndvi.copy_(ndvi, non_blocking=True)
        self._replay_or_run(self._segs, eager=False, fgrad=self._ar[0], dec=self._ar[1])
        return self._static_losses
