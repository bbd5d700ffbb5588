"""Pin the CPU oracle (oracle/sifnn_oracle.py) against outputs of the REFERENCE ITSELF
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

import sifnn_oracle as O
from conftest import load_ckpt, load_golden, rel_err

TOL = 1e-5  # oracle and reference run the same torch CPU kernels; differences are thread-order noise


def _x_syn():
    b = load_golden("bicubic.npz")
    lst = torch.from_numpy(b["lst"])
    g = torch.Generator().manual_seed(1234)
    lst2 = torch.randn(2, 1, 64, 64, generator=g)
    ndvi = torch.randn(2, 1, 256, 256, generator=g)
    assert torch.equal(lst, lst2)
    return lst, torch.from_numpy(b["up_cv2"]), ndvi


def test_state_dict_contract():
    sd = load_ckpt("1009")
    table = O.conv_table()
    assert len(sd) == 104 and len(table) == 18
    mine = O.init_state_dict(0)
    assert list(mine.keys()) == list(sd.keys())
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape) and mine[k].dtype == sd[k].dtype, k
    assert len(O.trainable_keys(sd)) == 53
    assert sum(sd[k].numel() for k in O.trainable_keys(sd)) == 282705


def test_bicubic_matches_cv2():
    lst, up_cv2, _ = _x_syn()
    assert rel_err(O.bicubic_up4(lst), up_cv2) < 2e-6


@pytest.mark.parametrize("tag", ["1009", "2609", "2011"])
def test_forward_eval(tag):
    sd = load_ckpt(tag)
    fw = load_golden("fwd_eval.npz")
    _, up, ndvi = _x_syn()
    with torch.no_grad():
        y = O.forward(sd, torch.cat((up, ndvi), 1), train=False)
        ys = O.forward(sd, torch.from_numpy(fw["x_small"]), train=False)
    assert rel_err(y, fw[f"y_syn_{tag}"]) < TOL
    assert rel_err(ys, fw[f"y_small_{tag}"]) < TOL


def test_forward_eval_real_pairs():
    sd = load_ckpt("1009")
    fw, rp, bc = load_golden("fwd_eval.npz"), load_golden("real_pairs.npz"), load_golden("bicubic.npz")
    ndvi = (np.clip(rp["ndvi"], -1, 1)[:, None] - O.MEAN_NDVI) / O.STD_NDVI
    x = torch.cat((torch.from_numpy(bc["real_up_cv2"]), torch.from_numpy(ndvi).float()), 1)
    with torch.no_grad():
        y = O.forward(sd, x, train=False)
    assert rel_err(y, fw["y_real_1009"]) < TOL


def test_forward_fp32_noise_floor_vs_fp64():
    fw = load_golden("fwd_eval.npz")
    assert rel_err(fw["y_syn_1009"], fw["y_syn_1009_f64"]) < 5e-6


def test_forward_train_and_bn_buffers():
    sd = load_ckpt("1009")
    g = load_golden("fwd_train.npz")
    _, up, ndvi = _x_syn()
    ns = {}
    with torch.no_grad():
        y = O.forward(sd, torch.cat((up, ndvi), 1), train=True, new_stats=ns)
    assert rel_err(y, g["y"]) < TOL
    assert len(ns) == 17 * 3
    for k, v in ns.items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(g[k])
        else:
            assert rel_err(v, g[k]) < TOL, k


def test_loss_helpers():
    g = load_golden("loss_helpers.npz")
    x = torch.from_numpy(g["x"])
    assert np.abs(O.psf_kernel(0.1).numpy() - g["psf_01"]).max() < 1e-9
    assert np.abs(O.psf_kernel(0.25).numpy() - g["psf_025"]).max() < 1e-9
    assert rel_err(O.downscale_to_lr(x * O.STD_LST + O.MEAN_LST), g["down"]) < 1e-6
    assert rel_err(O.lowpass(x, 0.25), g["ftm_025"]) < 1e-6


@pytest.mark.parametrize("kind", ["sr1", "sr2"])
def test_train_step(kind):
    g = load_golden(f"step_{kind}.npz")
    alpha, gamma, lr = g["hyper"]
    lst, up, ndvi = _x_syn()
    tr = O.Trainer(load_ckpt("1009"), kind, alpha, gamma, lr)
    sr, scalars, dsr = tr.loss_and_grads(lst, up, ndvi)
    assert rel_err(sr, g["sr"]) < TOL
    assert np.allclose(scalars, g["losses"], rtol=1e-5)
    assert rel_err(dsr, g["dsr"]) < 1e-4
    assert rel_err(tr.flat_grads(), g["grads"]) < 1e-4
    if "params_after" in g:
        tr.opt.step()
        assert rel_err(tr.flat_params(), g["params_after"]) < 1e-6


def test_loss_curve_head():
    """First steps of the 100-step SR2 curve (the GPU test replays all 100)."""
    c = load_golden("curve_100.npz")
    init = {k: torch.from_numpy(v) for k, v in load_golden("curve_init.npz").items()}
    lst, up, ndvi = O.synthetic_batch(4)
    tr = O.Trainer(init, "sr2", 0.5, -0.25, 1e-3)
    for i in range(3):
        s = tr.step(lst, up, ndvi)
        assert np.allclose(s, c["sr2_f32"][i], rtol=2e-4), (i, s, c["sr2_f32"][i])
