"""Thin tensor-level wrappers over the per-op C-ABI entry points (include/sifnn.h).

They allocate outputs with torch, pass raw device pointers + the current stream, and raise
on error.  Used by the parity tests and available to callers that want a single fused op;
``ModelB_2`` / ``Trainer`` call the whole-network entry points instead.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import SifnnError


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise SifnnError("CUDA tensors required (no CPU fallback)")
        if not t.is_contiguous():
            raise SifnnError("contiguous tensors required")


def _p(t):
    return None if t is None else t.data_ptr()


def conv3x3_fwd(x, w, bias=None, in_scale=None, in_shift=None, stats: Optional[torch.Tensor] = None):
    _chk(x, w, bias, in_scale, in_shift, stats)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    _lib.call("sifnn_conv3x3_fwd", _p(x), _p(in_scale), _p(in_shift), _p(w), _p(bias), _p(out), _p(stats), B, Cin, Cout, H, W, _s())
    return out


def conv3x3_dgrad(dy, w, dx: Optional[torch.Tensor] = None, accumulate: bool = False):
    _chk(dy, w, dx)
    B, Cout, H, W = dy.shape
    Cin = w.shape[1]
    if dx is None:
        dx = torch.empty((B, Cin, H, W), dtype=torch.float32, device=dy.device)
        accumulate = False
    _lib.call("sifnn_conv3x3_dgrad", _p(dy), _p(w), _p(dx), 1 if accumulate else 0, B, Cin, Cout, H, W, _s())
    return dx


def conv3x3_fwd_tc(x, w, bias=None, in_scale=None, in_shift=None, stats: Optional[torch.Tensor] = None):
    """tcgen05 / TMEM implicit-GEMM forward (3-term TF32 split)."""
    _chk(x, w, bias, in_scale, in_shift, stats)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    lib = _lib.load()
    if not lib.sifnn_conv3x3_tc_supported(Cin, Cout, H, W):
        raise SifnnError(f"conv3x3_fwd_tc: unsupported shape {tuple(x.shape)} -> {Cout}")
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(Cin, Cout), dtype=torch.uint8, device=x.device)
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    _lib.call("sifnn_conv3x3_fwd_tc", _p(x), _p(in_scale), _p(in_shift), _p(w), _p(bias), _p(out), _p(stats), _p(wprep), B, Cin, Cout, H, W, _s())
    return out


def conv3x3_dgrad_tc(dy, w, dx: Optional[torch.Tensor] = None, accumulate: bool = False):
    _chk(dy, w, dx)
    B, Cout, H, W = dy.shape
    Cin = w.shape[1]
    lib = _lib.load()
    if not lib.sifnn_conv3x3_tc_supported(Cout, Cin, H, W):
        raise SifnnError(f"conv3x3_dgrad_tc: unsupported shape {tuple(dy.shape)} -> {Cin}")
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(Cout, Cin), dtype=torch.uint8, device=dy.device)
    if dx is None:
        dx = torch.empty((B, Cin, H, W), dtype=torch.float32, device=dy.device)
        accumulate = False
    _lib.call("sifnn_conv3x3_dgrad_tc", _p(dy), _p(w), _p(dx), 1 if accumulate else 0, _p(wprep), B, Cin, Cout, H, W, _s())
    return dx


def conv3x3_fwd_ff(x, w, in_scale=None, in_shift=None, stats: Optional[torch.Tensor] = None):
    """Full-fold tcgen05 forward (csrc/conv3x3_ff.cu): nine taps per MMA, shift-and-add epilogue."""
    _chk(x, w, in_scale, in_shift, stats)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    lib = _lib.load()
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(Cin, Cout), dtype=torch.uint8, device=x.device)
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    _lib.call("sifnn_conv3x3_fwd_ff", _p(x), _p(in_scale), _p(in_shift), _p(w), _p(out), _p(stats), _p(wprep), B, Cin, Cout, H, W, _s())
    return out


def conv3x3_dgrad_ff(dy, w, dx: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Full-fold tcgen05 data gradient, padding adjoint included (one launch)."""
    _chk(dy, w, dx)
    B, Cout, H, W = dy.shape
    Cin = w.shape[1]
    lib = _lib.load()
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(Cout, Cin), dtype=torch.uint8, device=dy.device)
    if dx is None:
        dx = torch.empty((B, Cin, H, W), dtype=torch.float32, device=dy.device)
        accumulate = False
    _lib.call("sifnn_conv3x3_dgrad_ff", _p(dy), _p(w), _p(dx), 1 if accumulate else 0, _p(wprep), B, Cin, Cout, H, W, _s())
    return dx


def conv3x3_fwd_fs(x, w, in_scale=None, in_shift=None, stats: Optional[torch.Tensor] = None):
    """Fold + shift tcgen05 forward (csrc/conv3x3_fs.cu), widths that are multiples of 128."""
    _chk(x, w, in_scale, in_shift, stats)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    lib = _lib.load()
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(Cin, Cout), dtype=torch.uint8, device=x.device)
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    _lib.call("sifnn_conv3x3_fwd_fs", _p(x), _p(in_scale), _p(in_shift), _p(w), _p(out), _p(stats), _p(wprep), B, Cin, Cout, H, W, _s())
    return out


def conv3x3_dgrad_fs(dy, w, dx: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Fold + shift tcgen05 data gradient, padding adjoint included (one launch)."""
    _chk(dy, w, dx)
    B, Cout, H, W = dy.shape
    Cin = w.shape[1]
    lib = _lib.load()
    wprep = torch.empty(lib.sifnn_conv3x3_tc_wprep_bytes(Cout, Cin) + 2 * Cout * 3 * Cin * 4, dtype=torch.uint8, device=dy.device)
    if dx is None:
        dx = torch.empty((B, Cin, H, W), dtype=torch.float32, device=dy.device)
        accumulate = False
    _lib.call("sifnn_conv3x3_dgrad_fs", _p(dy), _p(w), _p(dx), 1 if accumulate else 0, _p(wprep), B, Cin, Cout, H, W, _s())
    return dx


def conv3x3_wgrad(x, dy, in_scale=None, in_shift=None, want_bias: bool = False):
    _chk(x, dy, in_scale, in_shift)
    B, Cin, H, W = x.shape
    Cout = dy.shape[1]
    nbytes = _lib.load().sifnn_conv3x3_wgrad_workspace(B, Cin, Cout, H, W)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=x.device)
    db = torch.empty((Cout,), dtype=torch.float32, device=x.device) if want_bias else None
    _lib.call("sifnn_conv3x3_wgrad", _p(x), _p(in_scale), _p(in_shift), _p(dy), _p(dw), _p(db), _p(ws), B, Cin, Cout, H, W, _s())
    return (dw, db) if want_bias else dw


def conv3x3_wgrad_tc(x, dy, in_scale=None, in_shift=None):
    """Weight gradient on the tcgen05 tensor cores (csrc/wgrad_tc.cu); same result contract as conv3x3_wgrad."""
    _chk(x, dy, in_scale, in_shift)
    B, Cin, H, W = x.shape
    Cout = dy.shape[1]
    lib = _lib.load()
    if not lib.sifnn_conv3x3_wgrad_tc_supported(Cin, Cout, H, W):
        raise _lib.SifnnError(f"conv3x3_wgrad_tc: unsupported shape Cin={Cin} Cout={Cout} H={H} W={W}")
    nbytes = lib.sifnn_conv3x3_wgrad_tc_workspace(B, Cin, Cout, H, W)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=x.device)
    _lib.call("sifnn_conv3x3_wgrad_tc", _p(x), _p(in_scale), _p(in_shift), _p(dy), _p(dw), _p(ws), B, Cin, Cout, H, W, _s())
    return dw


def conv3x3_wgrad_km(x, dy, in_scale=None, in_shift=None):
    """Weight gradient with 16-bit K-major operands (csrc/wgrad_km.cu); same result contract as conv3x3_wgrad."""
    _chk(x, dy, in_scale, in_shift)
    B, Cin, H, W = x.shape
    Cout = dy.shape[1]
    lib = _lib.load()
    if not lib.sifnn_conv3x3_wgrad_km_supported(Cin, Cout, H, W):
        raise _lib.SifnnError(f"conv3x3_wgrad_km: unsupported shape Cin={Cin} Cout={Cout} H={H} W={W}")
    ws = torch.empty(max(lib.sifnn_conv3x3_wgrad_km_workspace(B, Cin, Cout, H, W), 16), dtype=torch.uint8, device=x.device)
    dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=x.device)
    _lib.call("sifnn_conv3x3_wgrad_km", _p(x), _p(in_scale), _p(in_shift), _p(dy), _p(dw), _p(ws), B, Cin, Cout, H, W, _s())
    return dw


def bn_train_finalize(stats, gamma, beta, n: float, running_mean=None, running_var=None):
    _chk(stats, gamma, beta, running_mean, running_var)
    C = gamma.numel()
    outs = [torch.empty(C, dtype=torch.float32, device=gamma.device) for _ in range(4)]
    _lib.call("sifnn_bn_train_finalize", _p(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
              *[_p(o) for o in outs], C, float(n), _s())
    return tuple(outs)  # scale, shift, mean, invstd


def bn_eval_affine(gamma, beta, running_mean, running_var):
    _chk(gamma, beta, running_mean, running_var)
    C = gamma.numel()
    sc, sh = (torch.empty(C, dtype=torch.float32, device=gamma.device) for _ in range(2))
    _lib.call("sifnn_bn_eval_affine", _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(sc), _p(sh), C, _s())
    return sc, sh


def bn_relu_bwd(dY, raw, scale, shift, mean, invstd, gamma):
    _chk(dY, raw, scale, shift, mean, invstd, gamma)
    B, C, H, W = raw.shape
    sums = torch.zeros(2 * C, dtype=torch.float64, device=raw.device)
    _lib.call("sifnn_bn_relu_bwd_reduce", _p(dY), _p(raw), _p(scale), _p(shift), _p(mean), _p(invstd), _p(sums), B, C, H * W, _s())
    dx = torch.empty_like(raw)
    dg, db = (torch.empty(C, dtype=torch.float32, device=raw.device) for _ in range(2))
    _lib.call("sifnn_bn_relu_bwd_apply", _p(dY), _p(raw), _p(scale), _p(shift), _p(mean), _p(invstd), _p(gamma), _p(sums),
              _p(dx), _p(dg), _p(db), B, C, H * W, _s())
    return dx, dg, db


def act_avgpool2_fwd(raw, scale, shift):
    _chk(raw, scale, shift)
    B, C, H, W = raw.shape
    out = torch.empty((B, C, H // 2, W // 2), dtype=torch.float32, device=raw.device)
    _lib.call("sifnn_act_avgpool2_fwd", _p(raw), _p(scale), _p(shift), _p(out), B, C, H, W, _s())
    return out


def avgpool2_bwd(dout, din: Optional[torch.Tensor] = None, accumulate: bool = False):
    _chk(dout, din)
    B, C, Ho, Wo = dout.shape
    if din is None:
        din = torch.empty((B, C, 2 * Ho, 2 * Wo), dtype=torch.float32, device=dout.device)
        accumulate = False
    _lib.call("sifnn_avgpool2_bwd", _p(dout), _p(din), 1 if accumulate else 0, B, C, 2 * Ho, 2 * Wo, _s())
    return din


def act_residual_fwd(x, raw, scale, shift):
    _chk(x, raw, scale, shift)
    B, C, H, W = raw.shape
    out = torch.empty_like(raw)
    _lib.call("sifnn_act_residual_fwd", _p(x), _p(raw), _p(scale), _p(shift), _p(out), B, C, H * W, _s())
    return out


def act_upcat_fwd(low, low_scale, low_shift, skip, skip_scale, skip_shift):
    _chk(low, low_scale, low_shift, skip, skip_scale, skip_shift)
    B, C1, H, W = low.shape
    C2 = skip.shape[1]
    out = torch.empty((B, C1 + C2, 2 * H, 2 * W), dtype=torch.float32, device=low.device)
    _lib.call("sifnn_act_upcat_fwd", _p(low), _p(low_scale), _p(low_shift), _p(skip), _p(skip_scale), _p(skip_shift), _p(out),
              B, C1, C2, H, W, _s())
    return out


def upcat_bwd(dout, C1: int) -> Tuple[torch.Tensor, torch.Tensor]:
    _chk(dout)
    B, Ct, Ho, Wo = dout.shape
    C2 = Ct - C1
    dlow = torch.empty((B, C1, Ho // 2, Wo // 2), dtype=torch.float32, device=dout.device)
    dskip = torch.empty((B, C2, Ho, Wo), dtype=torch.float32, device=dout.device)
    _lib.call("sifnn_upcat_bwd", _p(dout), _p(dlow), _p(dskip), B, C1, C2, Ho // 2, Wo // 2, _s())
    return dlow, dskip


def adam_step(p, g, m, v, step_count, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    _chk(p, g, m, v, step_count)
    _lib.call("sifnn_adam_step", _p(p), _p(g), _p(m), _p(v), _p(step_count), float(lr), float(beta1), float(beta2), float(eps),
              float(grad_scale), p.numel(), _s())


def fp32_peak_tflops(iters: int = 4000, reps: int = 5) -> float:
    """Measured fp32 FFMA throughput of this GPU (TFLOP/s): the roof for the SIMT kernels."""
    import ctypes
    sink = torch.zeros(4, dtype=torch.float32, device="cuda")
    flops = ctypes.c_double()
    best = 0.0
    _lib.call("sifnn_fp32_peak_kernel", _p(sink), 100, ctypes.byref(flops), _s())
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("sifnn_fp32_peak_kernel", _p(sink), iters, ctypes.byref(flops), _s())
        e1.record()
        e1.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def psnr_ssim(pred, target):
    """Mean PSNR and SSIM of a batch (B,1,H,W) on the device, as ``us.psnr_skimage`` / ``us.ssim_skimage`` (utils.py:548-578) compute
    them on the host.  Returns a (2,) float32 device tensor (psnr, ssim) without synchronising."""
    _chk(pred, target)
    if pred.shape != target.shape or pred.dim() != 4 or pred.shape[1] != 1:
        raise _lib.SifnnError("psnr_ssim: pred and target must both be (B,1,H,W)")
    B, _, H, W = pred.shape
    ws = torch.empty(_lib.load().sifnn_quality_workspace_bytes(B), dtype=torch.uint8, device=pred.device)
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    _lib.call("sifnn_quality_psnr_ssim", _p(pred), _p(target), _p(out), _p(ws), B, H, W, _s())
    return out
