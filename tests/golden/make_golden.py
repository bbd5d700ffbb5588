#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Only runs in the dev container (needs /root/reference, which does not exist on the
GPU box).  It imports the reference's ``model.py`` and ``utils.py`` unmodified
(``utils`` behind empty stubs for the absent matplotlib / skimage / osgeo packages),
drives them with seeded inputs and writes small .npz fixtures.  Nothing of the
reference's source is copied; the train-step driver below calls the reference's
functions in the order its ``train_step`` does
(train_model_B_gradFTM.py:94-121, train_model_B_predef_filters.py:106-137).

    python tests/golden/make_golden.py
"""
import os
import pickle
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
MEAN_LST, STD_LST, MEAN_NDVI, STD_NDVI = 307.24, 5.57, 0.645, 0.168


def import_reference():
    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            m = _Stub(self.__name__ + "." + name)
            sys.modules[m.__name__] = m
            return m

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "skimage", "skimage.metrics", "skimage.measure",
                 "skimage.transform", "skimage.filters", "osgeo", "osgeo.gdal", "osgeo.osr", "osgeo.gdalconst",
                 "pymp", "pymodis", "rasterio", "shapely", "affine", "pyproj", "torchinfo"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    import model as ref_model  # noqa
    import utils as ref_utils  # noqa
    return ref_model, ref_utils


def build(ref_model):
    return ref_model.ModelB_2(in_channels=2, downchannels=[16, 32, 64, 128], padding_mode="replicate",
                              activation="ReLU", bilinear=1, n_bridge_blocks=1)


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


class _AnyStub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.state = s


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("rasterio", "affine", "shapely", "pyproj"):
            return _AnyStub
        return super().find_class(module, name)


def load_real_pairs(n):
    ddir = os.path.join(REF, "test_data_formatted", "data")
    ids = sorted(int(f.split("_")[0]) for f in os.listdir(ddir) if f.endswith("_data_dict.pkl"))[:n]
    lst, ndvi = [], []
    for i in ids:
        with open(os.path.join(ddir, f"{i}_data_dict.pkl"), "rb") as fh:
            d = _StubUnpickler(fh).load()
        arrs = {k: v for k, v in d.items() if isinstance(v, np.ndarray)}
        l = next(v for v in arrs.values() if v.shape == (64, 64))
        nd = next(v for v in arrs.values() if v.shape == (256, 256))
        lst.append(l.astype(np.float32))
        ndvi.append(nd.astype(np.float32))
    return np.array(ids), np.stack(lst), np.stack(ndvi)


def write_all_real_pairs():
    """Every real MODIS pair the reference ships (test_data_formatted/data/*_data_dict.pkl, 83 of them) as raw inputs: LST (64,64) in Kelvin,
    NDVI (256,256).  The oracle is run live on them by tests/test_gpu_fullsize.py, so no outputs are stored.  ~17 MB (fp32 NDVI mantissas do
    not compress)."""
    ids, lst, ndvi = load_real_pairs(10 ** 6)
    np.savez_compressed(os.path.join(OUT, "real_pairs_all.npz"), ids=ids, lst=lst, ndvi=ndvi)
    print("all real pairs:", ids.shape, lst.shape, ndvi.shape)


def ref_step_losses(us, model, kind, lst, lst_up, ndvi, alpha, gamma, loss_fn):
    """The reference train_step body, reference functions only."""
    lst_ndvi = torch.cat((lst_up, ndvi), dim=1)
    lst_SR = model(lst_ndvi)
    lst_SR.retain_grad()
    mean, std = MEAN_LST, STD_LST
    lst_SR_down = us.downscale_LST_SR_to_LR(lst_SR * std + mean)
    lst_SR_down = (lst_SR_down - mean) / std
    ds_loss = loss_fn(lst_SR_down, lst)
    if kind == "sr2":
        grads_lst = lst_SR - us.get_output_ftm(lst_SR, mtf=0.25)
        grads_ndvi = ndvi - us.get_output_ftm(ndvi, mtf=0.25)
    else:
        sys.path.insert(0, REF)
        filt = [[[1, 2, 1], [0, 0, 0], [-1, -2, -1]], [[1, 0, -1], [2, 0, -2], [1, 0, -1]],
                [[2, 1, 0], [1, 0, -1], [0, -1, -2]], [[0, 1, 2], [-1, 0, 1], [-2, -1, 0]]]
        ft = torch.zeros((4, 1, 3, 3), dtype=lst_SR.dtype)
        for i in range(4):
            ft[i, 0] = torch.tensor(filt[i], dtype=lst_SR.dtype)
        grads_lst = F.conv2d(lst_SR, ft, padding="same")
        grads_ndvi = F.conv2d(ndvi, ft, padding="same")
    percep = loss_fn(grads_lst, gamma * grads_ndvi)
    loss = alpha * ds_loss + (1 - alpha) * percep
    return lst_SR, ds_loss, percep, loss


def synthetic(batch, seed=1234, hr=256):
    g = torch.Generator().manual_seed(seed)
    lst = torch.randn(batch, 1, hr // 4, hr // 4, generator=g)
    ndvi = torch.randn(batch, 1, hr, hr, generator=g)
    return lst, ndvi


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref_model, us = import_reference()
    import cv2

    # ---- checkpoints ------------------------------------------------------------
    ckpts = {}
    for tag in ("1009", "2609", "2011"):
        sd = torch.load(f"{REF}/models/modelB_{tag}/modelB_state_dict.pt", map_location="cpu")
        ckpts[tag] = sd
        np.savez(os.path.join(OUT, f"ckpt_modelB_{tag}.npz"), **sd_to_np(sd))
    print("checkpoints:", {k: len(v) for k, v in ckpts.items()})

    # ---- real pairs ---------------------------------------------------------------
    ids, rl, rn = load_real_pairs(2)
    np.savez_compressed(os.path.join(OUT, "real_pairs.npz"), ids=ids, lst=rl, ndvi=rn)
    print("real pairs", ids, rl.shape, rl.min(), rl.max(), rn.min(), rn.max())

    write_all_real_pairs()

    # ---- bicubic: cv2 (reference, utils.py:180) vs torch -------------------------------
    lst, ndvi = synthetic(2)
    up_cv2 = np.stack([us.upsampling(lst[i, 0].numpy(), (4, 4)) for i in range(2)])[:, None]
    up_t = F.interpolate(lst, scale_factor=4, mode="bicubic", align_corners=False).numpy()
    print("bicubic cv2 vs torch max abs", np.abs(up_cv2 - up_t).max())
    real_norm = (rl - MEAN_LST) / STD_LST
    up_real = np.stack([us.upsampling(real_norm[i], (4, 4)) for i in range(len(rl))])[:, None]
    np.savez_compressed(os.path.join(OUT, "bicubic.npz"), lst=lst.numpy(), up_cv2=up_cv2.astype(np.float32),
                        real_up_cv2=up_real.astype(np.float32))

    # ---- eval forward -------------------------------------------------------------
    fw = {}
    x_syn = torch.cat((torch.from_numpy(up_cv2.astype(np.float32)), ndvi), dim=1)
    x_real = torch.cat((torch.from_numpy(up_real.astype(np.float32)),
                        torch.from_numpy((np.clip(rn, -1, 1)[:, None] - MEAN_NDVI) / STD_NDVI).float()), dim=1)
    g = torch.Generator().manual_seed(99)
    x_small = torch.randn(3, 2, 64, 64, generator=g)
    fw["x_small"] = x_small.numpy()
    for tag, sd in ckpts.items():
        m = build(ref_model)
        print(tag, m.load_state_dict(sd))
        m.eval()
        with torch.inference_mode():
            fw[f"y_syn_{tag}"] = m(x_syn).numpy()
            fw[f"y_real_{tag}"] = m(x_real).numpy()
            fw[f"y_small_{tag}"] = m(x_small).numpy()
            if tag == "1009":
                fw[f"y_syn_{tag}_f64"] = m.double()(x_syn.double()).numpy()
    np.savez_compressed(os.path.join(OUT, "fwd_eval.npz"), **fw)

    # ---- train-mode forward + BN buffer update -----------------------------------------
    m = build(ref_model)
    m.load_state_dict(ckpts["1009"])
    m.train()
    with torch.no_grad():
        y_tr = m(x_syn)
    sd_after = m.state_dict()
    bn_after = {k: v.numpy() for k, v in sd_after.items() if "running" in k or "num_batches" in k}
    np.savez_compressed(os.path.join(OUT, "fwd_train.npz"), y=y_tr.numpy(), **bn_after)

    # ---- loss helpers -----------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    t = torch.randn(2, 1, 256, 256, generator=g)
    np.savez_compressed(
        os.path.join(OUT, "loss_helpers.npz"), x=t.numpy(),
        psf_01=us.generate_psf_kernel(1.0, 4, 0.1, None), psf_025=us.generate_psf_kernel(1.0, 4, 0.25, None),
        down=us.downscale_LST_SR_to_LR(t * STD_LST + MEAN_LST).numpy(),
        ftm_025=us.get_output_ftm(t, mtf=0.25).numpy())

    # ---- one full train step per loss, B=2, from the 1009 weights ---------------------------
    lst_up = torch.from_numpy(up_cv2.astype(np.float32))
    for kind, alpha, gamma, lr in (("sr1", 0.99, -0.5, 1e-3), ("sr2", 0.5, -0.25, 1e-4)):
        m = build(ref_model)
        m.load_state_dict(ckpts["1009"])
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=lr)
        loss_fn = torch.nn.HuberLoss()
        opt.zero_grad()
        sr, ds, pl, loss = ref_step_losses(us, m, kind, lst, lst_up, ndvi, alpha, gamma, loss_fn)
        loss.backward()
        grads = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        opt.step()
        params = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        extra = {"params_after": params.numpy()} if kind == "sr1" else {}
        np.savez_compressed(os.path.join(OUT, f"step_{kind}.npz"), sr=sr.detach().numpy(), dsr=sr.grad.numpy(),
                            losses=np.array([ds.item(), pl.item(), loss.item()]), grads=grads.numpy(),
                            hyper=np.array([alpha, gamma, lr]), **extra)
        print(kind, ds.item(), pl.item(), loss.item(), grads.abs().max().item())

    # ---- 100-step SR2 loss curve, B=4, seed-0 default init, fp32 and fp64 --------------
    torch.manual_seed(0)
    m0 = build(ref_model)
    init_sd = {k: v.clone() for k, v in m0.state_dict().items()}
    np.savez(os.path.join(OUT, "curve_init.npz"), **sd_to_np(init_sd))
    lst4, ndvi4 = synthetic(4)
    up4 = F.interpolate(lst4, scale_factor=4, mode="bicubic", align_corners=False)
    curves = {}
    for kind, alpha, gamma, lr in (("sr2", 0.5, -0.25, 1e-3), ("sr1", 0.99, -0.5, 1e-3)):
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            if kind == "sr1" and name == "f32":
                continue
            m = build(ref_model)
            m.load_state_dict(init_sd)
            m = m.to(dt).train()
            opt = torch.optim.Adam(m.parameters(), lr=lr)
            loss_fn = torch.nn.HuberLoss()
            rec = []
            for it in range(100):
                opt.zero_grad()
                _, ds, pl, loss = ref_step_losses(us, m, kind, lst4.to(dt), up4.to(dt), ndvi4.to(dt), alpha, gamma, loss_fn)
                loss.backward()
                opt.step()
                rec.append([ds.item(), pl.item(), loss.item()])
            curves[f"{kind}_{name}"] = np.array(rec)
            print(kind, name, rec[0], rec[-1])
    np.savez_compressed(os.path.join(OUT, "curve_100.npz"), **curves)


if __name__ == "__main__":
    main()
