"""Drop-in ``ModelB_2`` (reference model.py:533-645) executed by libsifnn_b200.so.

The class names, constructor signatures, sub-module attribute names and therefore the
104 ``state_dict`` keys are those of the reference, so ``models/modelB_*/
modelB_state_dict.pt`` loads unchanged and ``modelB.pt`` (a pickled module that refers
to ``model.ModelB_2``, ``model.DoubleConvolution``, ...) un-pickles when this module is
importable as ``model``.  The sub-modules are *parameter containers only*: the
arithmetic of ``ModelB_2.forward`` and of its backward runs in hand-written sm_100a
kernels over one flat parameter buffer that the ``nn.Parameter`` objects are views of.
There is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import ModelBCfg, SifnnError

__all__ = ["ModelB_2", "DoubleConvolution", "UpBlock", "ResidualConnection", "DownBlock_pool", "DownBlock",
           "ResBridgeBlock", "Serf", "activation_functions"]


class Serf(nn.Module):
    """x * erf(softplus(x)) -- accepted by the constructors for signature parity
    (reference model.py:26-76); no shipped checkpoint uses it and the native forward
    rejects it."""

    def forward(self, x):  # pragma: no cover - not on the hot path
        raise SifnnError("activation 'Serf' is not implemented by the sm_100a path (no shipped checkpoint uses it)")


# one shared ReLU instance, like the reference (model.py:79-82)
activation_functions = {"ReLU": nn.ReLU(), "Serf": Serf()}


def _container_forward(self, *a, **k):
    raise SifnnError(f"{type(self).__name__} is a parameter container; only ModelB_2.forward is executed natively")


def _cbr(cin: int, cout: int, padding_mode: str, activation: str) -> List[nn.Module]:
    return [nn.Conv2d(cin, cout, 3, 1, 1, bias=False, padding_mode=padding_mode), nn.BatchNorm2d(cout),
            activation_functions[activation]]


class DoubleConvolution(nn.Module):
    """(Conv3x3 -> BatchNorm -> act) x 2; keys ``bloc.{0,1,3,4}`` (reference model.py:85-159)."""

    def __init__(self, in_channels, out_channels, mid_channels=None, padding_mode="zeros", activation="ReLU"):
        super().__init__()
        mid = mid_channels or out_channels
        self.bloc = nn.Sequential(*_cbr(in_channels, mid, padding_mode, activation), *_cbr(mid, out_channels, padding_mode, activation))

    forward = _container_forward


class UpBlock(nn.Module):
    """up x2 -> cat([up, skip]) -> DoubleConvolution(in, out, mid=in//2) (reference model.py:161-248)."""

    def __init__(self, in_channels, out_channels, bilinear=False, padding_mode="zeros", activation="ReLU"):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.convbloc = DoubleConvolution(in_channels, out_channels, in_channels // 2, padding_mode, activation=activation)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.convbloc = DoubleConvolution(in_channels, out_channels, padding_mode=padding_mode, activation=activation)

    forward = _container_forward


class ResidualConnection(nn.Module):
    """x + DoubleConvolution(x) (reference model.py:251-312)."""

    def __init__(self, in_channels, out_channels, padding_mode="zeros", activation="ReLU"):
        super().__init__()
        self.doubleconv = DoubleConvolution(in_channels, out_channels, padding_mode=padding_mode, activation=activation)

    forward = _container_forward


class ResBridgeBlock(nn.Module):
    """Unused by ModelB_2 (reference model.py:315-379); kept so ``from model import *`` users find it."""

    def __init__(self, in_channels, padding_mode="zeros", activation="ReLU"):
        super().__init__()
        self.block = nn.Sequential(*_cbr(in_channels, in_channels, padding_mode, activation),
                                   nn.Conv2d(in_channels, in_channels, 3, bias=False, padding=1, padding_mode=padding_mode),
                                   nn.BatchNorm2d(in_channels))

    forward = _container_forward


class _DownBase(nn.Module):
    def __init__(self, in_channels, out_channels, padding_mode, activation, downsampling):
        super().__init__()
        self.downsampling = downsampling
        self.resblock = ResidualConnection(in_channels, in_channels, padding_mode=padding_mode, activation=activation)
        self.lastconv = nn.Sequential(*_cbr(in_channels, out_channels, padding_mode, activation))

    forward = _container_forward


class DownBlock(_DownBase):
    """Strided-conv variant, unused by ModelB_2 (reference model.py:382-455)."""

    def __init__(self, in_channels, out_channels, padding_mode="zeros", activation="ReLU"):
        super().__init__(in_channels, out_channels, padding_mode, activation, nn.Conv2d(in_channels, in_channels, kernel_size=2, stride=2))


class DownBlock_pool(_DownBase):
    """AvgPool2 -> residual DoubleConvolution -> Conv.BN.act (reference model.py:458-531)."""

    def __init__(self, in_channels, out_channels, padding_mode="zeros", activation="ReLU"):
        super().__init__(in_channels, out_channels, padding_mode, activation, nn.AvgPool2d(kernel_size=2, stride=2))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Workspaces:
    """Per-model pool of workspace tensors keyed by (B, H, W, train, device)."""

    def __init__(self):
        self.free: Dict[Tuple, List[torch.Tensor]] = {}

    def take(self, key, nbytes: int, device) -> torch.Tensor:
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        return torch.empty(nbytes, dtype=torch.uint8, device=device)

    def give(self, key, t: torch.Tensor) -> None:
        lst = self.free.setdefault(key, [])
        if len(lst) < 2:
            lst.append(t)


class _ModelBFn(torch.autograd.Function):
    """Autograd boundary: one node for the whole network (forward + hand-written backward)."""

    @staticmethod
    def forward(ctx, model, x, *params):
        y, ws, key = model._run_forward(x, train=True, keep=True)
        ctx.model, ctx.ws, ctx.key = model, ws, key
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        model = ctx.model
        (x,) = ctx.saved_tensors
        if ctx.ws is None:
            raise SifnnError("backward called twice on the same ModelB_2 forward (workspace already released)")
        grads = model._run_backward(x, gy.contiguous(), ctx.ws)
        model._ws.give(ctx.key, ctx.ws)
        ctx.ws = None
        grads = grads.clone()  # fgrad is reused by the next backward; autograd may keep what we return
        views = [grads[o:o + n].view(s) for (o, n, s) in model._pviews]
        return (None, None, *views)


class ModelB_2(nn.Module):
    """SIF-NN-SR ModelB: 4-level residual U-Net, (B,2,H,W) -> (B,1,H,W)
    (reference model.py:533-645; constructor signature model.py:563, forward model.py:608)."""

    #: patches per launch plan in eval mode (bounds the workspace for very large batches)
    eval_chunk = 32

    def __init__(self, in_channels, downchannels=[16, 32, 64, 128], padding_mode="replicate", activation="ReLU",
                 bilinear=True, n_bridge_blocks=1):
        super().__init__()
        d = downchannels
        self.in_channels, self.downchannels, self.padding, self.activation = in_channels, downchannels, padding_mode, activation
        self.upfactor = 2 if bilinear else 1
        self.bridge = n_bridge_blocks  # stored and ignored, like the reference (model.py:592)
        self.inbloc = DoubleConvolution(in_channels, d[0], padding_mode=padding_mode, activation=activation)
        self.db1 = DownBlock_pool(d[0], d[1], padding_mode, activation=activation)
        self.db2 = DownBlock_pool(d[1], d[2], padding_mode, activation=activation)
        self.db3 = DownBlock_pool(d[2], d[3] // self.upfactor, padding_mode, activation=activation)
        self.ub1 = UpBlock(d[3], d[2] // self.upfactor, bilinear, padding_mode, activation=activation)
        self.ub2 = UpBlock(d[2], d[1] // self.upfactor, bilinear, padding_mode, activation=activation)
        self.ub3 = UpBlock(d[1], d[0], bilinear, padding_mode, activation=activation)
        self.outlay = nn.Conv2d(d[0], 1, kernel_size=3, stride=1, padding=1, padding_mode=padding_mode)

    def __getstate__(self):  # (graphs and native state are rebuilt on demand)
        d = self.__dict__.copy()
        d.pop("_sifnn_state", None)  # flat buffers / workspaces are rebuilt lazily
        d.pop("_eval_graphs", None)  # CUDA graphs are not picklable; re-enable with enable_eval_graphs()
        d.pop("_eval_graph_max", None)
        return d

    # ------------------------------------------------------------------ native state
    def _native_state(self):
        st = self.__dict__.get("_sifnn_state")
        if st is None:  # also reached for modules restored by torch.load of a pickled reference model
            st = {"flat": None, "ws": _Workspaces()}
            self.__dict__["_sifnn_state"] = st
        return st

    @property
    def _ws(self) -> _Workspaces:
        return self._native_state()["ws"]

    @property
    def _pviews(self):
        return self._native_state()["pviews"]

    def _check_supported(self):
        if self.padding != "replicate" or self.activation != "ReLU" or self.upfactor != 2:
            raise SifnnError("the sm_100a path implements the shipped configuration only: padding_mode='replicate', "
                             f"activation='ReLU', bilinear=True (got {self.padding!r}, {self.activation!r}, bilinear={self.upfactor == 2})")

    def _cfg(self) -> ModelBCfg:
        c = ModelBCfg()
        c.in_channels = int(self.in_channels)
        for i in range(4):
            c.down[i] = int(self.downchannels[i])
        return c

    def _bn_layers(self) -> List[nn.BatchNorm2d]:
        return [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]

    def _flatten(self, device) -> None:
        """Re-home all parameters / BN buffers as views of flat device buffers (layout =
        module.parameters() order = sifnn_modelb_param_layout)."""
        self._check_supported()
        st = self._native_state()
        lib = _lib.load()
        cfg = self._cfg()
        A17, A18 = ctypes.c_int64 * 17, ctypes.c_int64 * 18
        w_off, g_off, b_off, bn_off = A18(), A17(), A17(), A17()
        bias_off, bn_total = ctypes.c_int64(), ctypes.c_int64()
        n = lib.sifnn_modelb_param_layout(ctypes.byref(cfg), w_off, g_off, b_off, ctypes.byref(bias_off), bn_off, ctypes.byref(bn_total))
        if n < 0:
            _lib.check(1, "sifnn_modelb_param_layout")
        params = list(self.parameters())
        expect = []
        for i in range(17):
            expect += [w_off[i], g_off[i], b_off[i]]
        expect += [w_off[17], bias_off.value]
        if len(params) != len(expect) or sum(p.numel() for p in params) != n:
            raise SifnnError(f"parameter list does not match the native layout ({len(params)} tensors / {sum(p.numel() for p in params)} floats vs {len(expect)} / {n})")
        for p in params:
            if p.dtype != torch.float32:
                raise SifnnError("the sm_100a path is fp32; got parameter dtype %s" % p.dtype)
        flat = torch.empty(n, dtype=torch.float32, device=device)
        fgrad = torch.zeros(n, dtype=torch.float32, device=device)
        pviews = []
        with torch.no_grad():
            off = 0
            for p, e in zip(params, expect):
                k = p.numel()
                if off != e:
                    raise SifnnError("parameter order does not match the native layout")
                flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + k].view(p.shape)
                pviews.append((off, k, tuple(p.shape)))
                off += k
            bns = self._bn_layers()
            rm = torch.empty(bn_total.value, dtype=torch.float32, device=device)
            rv = torch.empty(bn_total.value, dtype=torch.float32, device=device)
            for i, bn in enumerate(bns):
                o, c = bn_off[i], bn.num_features
                rm[o:o + c].copy_(bn.running_mean)
                rv[o:o + c].copy_(bn.running_var)
                bn.running_mean.data = rm[o:o + c]
                bn.running_var.data = rv[o:o + c]
                if bn.num_batches_tracked.device != torch.device(device):
                    bn.num_batches_tracked.data = bn.num_batches_tracked.to(device)
        st.update(flat=flat, fgrad=fgrad, rm=rm, rv=rv, pviews=pviews, cfg=cfg, n=n, params=params,
                  ptrs=[p.data_ptr() for p in params], bns=bns, bn_ptrs=[b.running_mean.data_ptr() for b in bns],
                  counters=[b.num_batches_tracked for b in bns], dec_off=lib.sifnn_modelb_decoder_offset(ctypes.byref(cfg)))

    def _ensure_flat(self, device) -> dict:
        st = self._native_state()
        ok = st["flat"] is not None and st["flat"].device == device
        if ok:
            ok = all(p.data_ptr() == q for p, q in zip(st["params"], st["ptrs"])) and \
                all(b.running_mean.data_ptr() == q for b, q in zip(st["bns"], st["bn_ptrs"]))
        if not ok:
            self._flatten(device)
        return st

    # ------------------------------------------------------------------ native launches
    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise SifnnError("ModelB_2 (sm_100a) needs a CUDA tensor: there is no CPU fallback")
        if x.dtype != torch.float32:
            raise SifnnError(f"ModelB_2 (sm_100a) is fp32; got {x.dtype}")
        if x.dim() != 4 or x.shape[1] != self.in_channels or x.shape[2] % 8 or x.shape[3] % 8:
            raise SifnnError(f"expected (B,{self.in_channels},H,W) with H, W multiples of 8; got {tuple(x.shape)}")
        return x.contiguous()

    def _workspace_bytes(self, B: int, H: int, W: int, keep: bool) -> int:
        st = self._native_state()
        nbytes = _lib.load().sifnn_modelb_workspace_bytes(ctypes.byref(st["cfg"]), B, H, W, 1 if keep else 0)
        if nbytes == 0:
            raise SifnnError(f"unsupported shape (B={B}, H={H}, W={W})")
        return nbytes

    def _run_forward(self, x: torch.Tensor, train: bool, keep: bool, ws: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None):
        st = self._ensure_flat(x.device)
        B, _, H, W = x.shape
        key = (B, H, W, bool(keep), x.device)
        own_ws = ws is None
        if own_ws:
            ws = self._ws.take(key, self._workspace_bytes(B, H, W, keep), x.device)
        if y is None:
            y = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        _lib.call("sifnn_modelb_forward", ctypes.byref(st["cfg"]), st["flat"].data_ptr(), st["rm"].data_ptr(), st["rv"].data_ptr(),
                  x.data_ptr(), y.data_ptr(), ws.data_ptr(), B, H, W, 1 if train else 0, _stream())
        if train:
            torch._foreach_add_(st["counters"], 1)
        if keep or not own_ws:
            return y, ws, key
        self._ws.give(key, ws)
        return y, None, key

    def _run_backward(self, x, gy, ws, phase: int = 0) -> torch.Tensor:
        st = self._native_state()
        B, _, H, W = x.shape
        _lib.call("sifnn_modelb_backward", ctypes.byref(st["cfg"]), st["flat"].data_ptr(), x.data_ptr(), gy.data_ptr(),
                  st["fgrad"].data_ptr(), ws.data_ptr(), B, H, W, phase, _stream())
        return st["fgrad"]

    # ------------------------------------------------------------------ public API
    def forward(self, x_lst_ndvi):
        """y = Net(x): (B,2,H,W) fp32 CUDA -> (B,1,H,W)  (reference model.py:608-645)."""
        x = self._check_input(x_lst_ndvi)
        st = self._ensure_flat(x.device)
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in st["params"]))
        if need_grad:
            if x.requires_grad:
                raise SifnnError("gradient w.r.t. the network input is not part of the hot path (the reference never asks for it)")
            if self.training:
                return _ModelBFn.apply(self, x, *st["params"])
            # eval mode outside no_grad(): the reference never back-propagates through eval-mode BatchNorm;
            # the result is returned without an autograd graph.
        if self.training or x.shape[0] <= self.eval_chunk:
            return self._run_forward(x, train=self.training, keep=False)[0]
        outs = [self._run_forward(x[i:i + self.eval_chunk], train=False, keep=False)[0] for i in range(0, x.shape[0], self.eval_chunk)]
        return torch.cat(outs, dim=0)

    def forward_from_lowres(self, lst, ndvi):
        """Fused input stage: bicubic x4 of the (B,1,h,w) LST patch (cv2.INTER_CUBIC,
        reference utils.py:163-180) + concat with the (B,1,4h,4w) NDVI patch, then forward."""
        g = getattr(self, "_eval_graphs", None)
        if g is not None and not self.training and lst.is_cuda and lst.shape[0] <= self._eval_graph_max and not torch.cuda.is_current_stream_capturing():
            return self._graph_eval(lst, ndvi)
        return self.forward(bicubic4_cat(lst, ndvi))

    # ------------------------------------------------------------------ CUDA-graph replay for small eval batches
    def enable_eval_graphs(self, on: bool = True, max_batch: int = 8) -> None:
        """Replay a captured CUDA graph for eval-mode ``forward_from_lowres`` calls with at most ``max_batch`` patches.  The per-window
        loop of predict.py:84-103 runs batch 1, where the ~35 kernel launches of a forward cost more than the kernels themselves."""
        object.__setattr__(self, "_eval_graphs", {} if on else None)
        object.__setattr__(self, "_eval_graph_max", int(max_batch))

    def _graph_eval(self, lst: torch.Tensor, ndvi: torch.Tensor) -> torch.Tensor:
        if lst.dtype != torch.float32 or ndvi.dtype != torch.float32 or not ndvi.is_cuda:
            raise SifnnError("forward_from_lowres needs fp32 CUDA tensors")
        st = self._ensure_flat(lst.device)
        B, _, h, w = lst.shape
        key = (B, h, w, lst.device, st["flat"].data_ptr(), st["rm"].data_ptr())
        ent = self._eval_graphs.get(key)
        if ent is None:
            dev = lst.device
            ent = {"lst": torch.empty_like(lst, memory_format=torch.contiguous_format), "ndvi": torch.empty_like(ndvi, memory_format=torch.contiguous_format),
                   "x": torch.empty((B, 2, 4 * h, 4 * w), dtype=torch.float32, device=dev),
                   "y": torch.empty((B, 1, 4 * h, 4 * w), dtype=torch.float32, device=dev)}
            self._check_input(ent["x"])
            ent["ws"] = torch.empty(self._workspace_bytes(B, 4 * h, 4 * w, False), dtype=torch.uint8, device=dev)

            def run():
                bicubic4_cat(ent["lst"], ent["ndvi"], out=ent["x"])
                self._run_forward(ent["x"], train=False, keep=False, ws=ent["ws"], y=ent["y"])
            ent["lst"].copy_(lst)
            ent["ndvi"].copy_(ndvi)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                run()                      # warm-up outside capture (function attributes, lazy state)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            ent["graph"] = graph
            self._eval_graphs[key] = ent
        ent["lst"].copy_(lst, non_blocking=True)
        ent["ndvi"].copy_(ndvi, non_blocking=True)
        ent["graph"].replay()
        return ent["y"].clone()


def bicubic4_cat(lst: torch.Tensor, ndvi: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x = cat(bicubic_x4(lst), ndvi) on the device (reference utils.py:180 + train_model_B_gradFTM.py:94)."""
    if not (lst.is_cuda and ndvi.is_cuda) or lst.dtype != torch.float32 or ndvi.dtype != torch.float32:
        raise SifnnError("bicubic4_cat needs fp32 CUDA tensors")
    B, c, h, w = lst.shape
    if c != 1 or tuple(ndvi.shape) != (B, 1, 4 * h, 4 * w):
        raise SifnnError(f"bicubic4_cat: expected lst (B,1,h,w) and ndvi (B,1,4h,4w); got {tuple(lst.shape)} / {tuple(ndvi.shape)}")
    lst, ndvi = lst.contiguous(), ndvi.contiguous()
    x = out if out is not None else torch.empty((B, 2, 4 * h, 4 * w), dtype=torch.float32, device=lst.device)
    _lib.call("sifnn_bicubic4_cat", lst.data_ptr(), ndvi.data_ptr(), x.data_ptr(), B, h, w, _stream())
    return x
