"""Two epochs of the epoch loop (fit.py) on the GPU against the oracle doing the same epochs on the CPU (-m gpu)."""
import numpy as np
import pytest
import torch

import sifnn_b200
import sifnn_oracle as O
import model as model_mod

pytestmark = pytest.mark.gpu


def test_fit_two_epochs_matches_oracle(tmp_path):
    sd = O.init_state_dict(3)
    tb = [O.synthetic_batch(2, seed=100 + i) for i in range(2)]
    vb = [O.synthetic_batch(2, seed=200)]
    m = model_mod.ModelB_2(in_channels=2).cuda()
    m.load_state_dict(sd)
    tr = sifnn_b200.Trainer(m, "sr2", 0.5, -0.25, 1e-4)   # the SR2 checkpoint's own learning rate
    model, metrics = sifnn_b200.fit(tr, lambda: tb, lambda: vb, n_epochs=2, quality="device")
    ref = O.Trainer(sd, "sr2", 0.5, -0.25, 1e-4)
    for ep in range(2):
        acc = np.zeros(3)
        for lst, up, ndvi in tb:
            acc += np.array(ref.step(lst, up, ndvi))
        acc /= len(tb)
        got = np.array([metrics["train_dsloss"][ep], metrics["train_perceploss"][ep], metrics["train_loss"][ep]])
        # epoch 1 is two plain steps (1e-4 class); from the second epoch on Adam has amplified the fp32 rounding differences of the first
        # updates (test_gpu_model.py measures that noise against the reference's own fp32-vs-fp64 drift), hence the wider bar
        assert np.allclose(got, acc, rtol=2e-4 if ep == 0 else 1e-3), (ep, got, acc)
        with torch.no_grad():
            lst, up, ndvi = vb[0]
            sr = O.forward(ref.state_dict(), torch.cat((up, ndvi), 1), train=False)
            want = np.array([float(v) for v in O.sr2_losses(sr, lst, ndvi, 0.5, -0.25, O.MEAN_LST, O.STD_LST)])
        got = np.array([metrics["val_dsloss"][ep], metrics["val_perceploss"][ep], metrics["val_loss"][ep]])
        assert np.allclose(got, want, rtol=1e-3), (ep, got, want)
    assert metrics["best_epoch"] == 2 or "best_epoch" in metrics
    assert all(np.isfinite(v) and 0 < v for v in metrics["train_psnr"] + metrics["val_psnr"]) and all(-1 <= v <= 1 for v in metrics["train_ssim"] + metrics["val_ssim"])
    sd_file, md_file = sifnn_b200.save_model(model, str(tmp_path), "modelB")
    back = torch.load(sd_file, map_location="cpu")
    assert len(back) == 104 and all(not v.is_cuda for v in back.values())
    whole = torch.load(md_file, map_location="cpu", weights_only=False)
    assert type(whole).__module__ == "model"
