"""ctypes binding of libsifnn_b200.so (the C-ABI declared in include/sifnn.h).

The product path has NO fallback: if the shared library is missing or a call fails the
caller gets an exception, never a silent PyTorch/CPU substitute.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libsifnn_b200.so")
SOURCES = ["core.cu", "conv3x3.cu", "conv3x3_tc.cu", "conv3x3_ff.cu", "conv3x3_fs.cu", "wgrad.cu", "wgrad_tc.cu", "wgrad_km.cu", "elementwise.cu", "loss.cu", "quality.cu", "adam.cu", "modelb.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared"]


class SifnnError(RuntimeError):
    pass


class ModelBCfg(ctypes.Structure):
    _fields_ = [("in_channels", c_int), ("down", c_int * 4)]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"),
                                                        os.path.join(_HERE, "..", "include", "sifnn.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu of the library for sm_100a into one in-tree shared object."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    extra = os.environ.get("SIFNN_NVCC_EXTRA", "").split()   # e.g. -DSIFNN_MBAR_TEST_WAIT for A/B experiments
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise SifnnError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


_f = POINTER(c_float)
_d = POINTER(c_double)

# name -> (restype, argtypes); keep in sync with include/sifnn.h (tests/test_abi.py checks the symbols)
SIGNATURES = {
    "sifnn_version": (c_int, []),
    "sifnn_last_error": (c_char_p, []),
    "sifnn_launch_count": (ctypes.c_ulonglong, []),
    "sifnn_conv3x3_fwd": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_dgrad": (c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    "sifnn_conv3x3_tc_supported": (c_int, [c_int] * 4),
    "sifnn_conv3x3_tc_wprep_bytes": (c_size_t, [c_int] * 2),
    "sifnn_conv3x3_fwd_tc": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_dgrad_tc_main": (c_int, [c_void_p] * 3 + [c_int] + [c_void_p] + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_dgrad_tc": (c_int, [c_void_p] * 3 + [c_int] + [c_void_p] + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_ff_supported": (c_int, [c_int] * 4),
    "sifnn_conv3x3_ff_config": (None, [c_int, c_int]),
    "sifnn_conv3x3_ff_debug": (None, [c_int]),
    "sifnn_conv3x3_ff_trace": (None, [c_void_p]),
    "sifnn_conv3x3_fwd_ff": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_dgrad_ff": (c_int, [c_void_p] * 3 + [c_int] + [c_void_p] + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_fs_supported": (c_int, [c_int] * 4),
    "sifnn_conv3x3_fs_config": (None, [c_int, c_int]),
    "sifnn_conv3x3_fs_narrow": (None, [c_int]),
    "sifnn_conv3x3_fs_trace": (None, [c_void_p]),
    "sifnn_conv3x3_fwd_fs": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_dgrad_fs": (c_int, [c_void_p] * 3 + [c_int] + [c_void_p] + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_dgrad_border": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_wgrad_workspace": (c_size_t, [c_int] * 5),
    "sifnn_conv3x3_wgrad": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_wgrad_tc_supported": (c_int, [c_int] * 4),
    "sifnn_conv3x3_wgrad_tc_workspace": (c_size_t, [c_int] * 5),
    "sifnn_conv3x3_wgrad_tc": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_wgrad_km_supported": (c_int, [c_int] * 4),
    "sifnn_conv3x3_wgrad_km_workspace": (c_size_t, [c_int] * 5),
    "sifnn_conv3x3_wgrad_km_config": (None, [c_int, c_int]),
    "sifnn_conv3x3_wgrad_km": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_void_p]),
    "sifnn_conv3x3_wgrad_km_partials": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_void_p, c_void_p]),
    "sifnn_wgrad_reduce": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "sifnn_bn_train_finalize": (c_int, [c_void_p] * 9 + [c_int, c_double, c_void_p]),
    "sifnn_bn_eval_affine": (c_int, [c_void_p] * 6 + [c_int, c_void_p]),
    "sifnn_bn_relu_bwd_reduce": (c_int, [c_void_p] * 7 + [c_int] * 3 + [c_void_p]),
    "sifnn_bn_relu_bwd_apply": (c_int, [c_void_p] * 11 + [c_int] * 3 + [c_void_p]),
    "sifnn_act_avgpool2_fwd": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p]),
    "sifnn_avgpool2_bwd": (c_int, [c_void_p] * 2 + [c_int] * 5 + [c_void_p]),
    "sifnn_act_residual_fwd": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "sifnn_act_upcat_fwd": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "sifnn_upcat_bwd": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
    "sifnn_bicubic4_cat": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    "sifnn_tile_gather": (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_float] * 4 + [c_void_p]),
    "sifnn_tile_scatter": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_float] * 2 + [c_void_p]),
    "sifnn_loss_fwd_bwd": (c_int, [c_int] + [c_void_p] * 7 + [c_float, c_float] + [c_void_p] * 2 + [c_int] * 3 + [c_void_p]),
    "sifnn_quality_workspace_bytes": (c_size_t, [c_int]),
    "sifnn_quality_psnr_ssim": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "sifnn_adam_step": (c_int, [c_void_p] * 5 + [c_double] * 4 + [c_float, c_int64, c_void_p]),
    "sifnn_fp32_peak_kernel": (c_int, [c_void_p, c_int, _d, c_void_p]),
    "sifnn_set_tensor_cores": (None, [c_int]),
    "sifnn_get_tensor_cores": (c_int, []),
    "sifnn_modelb_param_layout": (c_int64, [POINTER(ModelBCfg)] + [POINTER(c_int64)] * 6),
    "sifnn_modelb_workspace_bytes": (c_size_t, [POINTER(ModelBCfg), c_int, c_int, c_int, c_int]),
    "sifnn_modelb_forward": (c_int, [POINTER(ModelBCfg)] + [c_void_p] * 6 + [c_int] * 4 + [c_void_p]),
    "sifnn_modelb_backward": (c_int, [POINTER(ModelBCfg)] + [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "sifnn_modelb_decoder_offset": (c_int64, [POINTER(ModelBCfg)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (must have been built; see __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SifnnError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                         "There is no CPU / PyTorch fallback for the SIF-NN hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sifnn_last_error()
        raise SifnnError(f"{what or 'sifnn call'} failed (code {rc}): {msg.decode() if msg else '?'}")


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise on a non-zero code."""
    check(getattr(load(), name)(*args), name)
