"""Drive the reference's OWN code (oracle/_ref/model.py, utils.py -- byte copies made by oracle/make_ref.py) on the CPU: the `kind: "reference"`
baseline of bench.py and a second anchor for the oracle port.

TEST / BENCH INFRASTRUCTURE ONLY: nothing in the product package imports this module.

The reference's training scripts cannot run as they are (they need GDAL and the absent training set), so the step below calls the reference's
functions in the order its ``train_step`` does -- train_model_B_gradFTM.py:94-121 (SR2) / train_model_B_predef_filters.py:106-137 (SR1) --
on synthetic tensors: ``model(x)``, ``us.downscale_LST_SR_to_LR``, ``us.get_output_ftm`` / the four Sobel filters of
train_model_B_predef_filters.py:38-42, ``nn.HuberLoss``, ``loss.backward()``, ``torch.optim.Adam.step()``.
"""
import os
import sys
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
MEAN_LST, STD_LST = 307.24, 5.57
_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "model.py")) and os.path.isfile(os.path.join(REF_DIR, "utils.py"))


def load():
    """Import the reference's model.py and utils.py (utils behind empty stubs for plotting / GDAL packages that are absent here)."""
    global _mods
    if _mods is not None:
        return _mods

    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            m = _Stub(self.__name__ + "." + name)
            sys.modules[m.__name__] = m
            return m

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "skimage", "skimage.metrics", "skimage.measure", "skimage.transform",
                 "skimage.filters", "osgeo", "osgeo.gdal", "osgeo.osr", "osgeo.gdalconst", "pymp", "pymodis", "rasterio", "shapely", "affine",
                 "pyproj", "torchinfo"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    import importlib.util
    out = []
    for name in ("model", "utils"):
        spec = importlib.util.spec_from_file_location(f"sifnn_reference_{name}", os.path.join(REF_DIR, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        out.append(mod)
    _mods = tuple(out)
    return _mods


def build_model(state_dict=None, seed: int = 0):
    ref_model, _ = load()
    torch.manual_seed(seed)
    m = ref_model.ModelB_2(in_channels=2, downchannels=[16, 32, 64, 128], padding_mode="replicate", activation="ReLU", bilinear=1, n_bridge_blocks=1)
    if state_dict is not None:
        m.load_state_dict(state_dict)
    return m


SOBEL4 = [[[1, 2, 1], [0, 0, 0], [-1, -2, -1]], [[1, 0, -1], [2, 0, -2], [1, 0, -1]],
          [[2, 1, 0], [1, 0, -1], [0, -1, -2]], [[0, 1, 2], [-1, 0, 1], [-2, -1, 0]]]   # train_model_B_predef_filters.py:38-42


class RefTrainer:
    """The reference's train_step body around the reference's own nn.Module and loss helpers."""

    def __init__(self, kind: str, alpha: float, gamma: float, lr: float, state_dict=None, seed: int = 0):
        _, self.us = load()
        self.kind, self.alpha, self.gamma = kind, alpha, gamma
        self.model = build_model(state_dict, seed).train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr)          # train_model_B_gradFTM.py:453
        self.loss_fn = torch.nn.HuberLoss()                                   # train_model_B_gradFTM.py:454
        self.filters = torch.tensor(SOBEL4, dtype=torch.float32)[:, None]

    def step(self, lst, lst_up, ndvi):
        us, model = self.us, self.model
        self.opt.zero_grad()
        sr = model(torch.cat((lst_up, ndvi), dim=1))                          # :94-96
        down = us.downscale_LST_SR_to_LR(sr * STD_LST + MEAN_LST)              # :99-103
        down = (down - MEAN_LST) / STD_LST
        ds = self.loss_fn(down, lst)
        if self.kind == "sr2":                                                # train_model_B_gradFTM.py:108-114
            g_sr = sr - us.get_output_ftm(sr, mtf=0.25)
            g_nd = ndvi - us.get_output_ftm(ndvi, mtf=0.25)
        else:                                                                 # train_model_B_predef_filters.py:127-130
            g_sr = F.conv2d(sr, self.filters, padding="same")
            g_nd = F.conv2d(ndvi, self.filters, padding="same")
        pl = self.loss_fn(g_sr, self.gamma * g_nd)
        loss = self.alpha * ds + (1 - self.alpha) * pl
        loss.backward()                                                       # :119
        self.opt.step()                                                       # :121
        return float(ds.detach()), float(pl.detach()), float(loss.detach())
