"""CPU checks of the *formulas* the CUDA kernels implement (no GPU needed): the host-built
loss tables, the gather-form adjoints and the replicate-padding border correction are
emulated with plain loops / torch ops and compared with autograd through the oracle.
These guard the algebra; the -m gpu tests check the kernels themselves."""
import numpy as np
import torch
import torch.nn.functional as F

import sifnn_b200
import sifnn_oracle as O
from sifnn_b200.losses import _tables_np, _reflect, gauss9


def test_psf_is_rank1_and_matches_oracle():
    for mtf in (0.1, 0.25):
        g = gauss9(mtf)
        k = O.psf_kernel(mtf).double().numpy()
        assert np.abs(np.outer(g, g) - k).max() < 1e-8


def _emulate_ds(sr, tb):
    """D = R sr R^T with R[I, reflect(4I-4+t)] += h12[t]  (what loss.cu P1/P2 compute)."""
    n = sr.shape[-1]
    R = np.zeros((n // 4, n))
    for i in range(n // 4):
        for t in range(12):
            R[i, _reflect(4 * i - 4 + t, n)] += tb["h12"][t].astype(np.float64)
    R = torch.from_numpy(R)
    return R @ sr.double() @ R.T, R


def test_ds_forward_and_adjoint_tables():
    n = 64
    tb = _tables_np(n)
    g = torch.Generator().manual_seed(3)
    sr = torch.randn(1, 1, n, n, generator=g, dtype=torch.float64, requires_grad=True)
    ref = O.downscale_to_lr(sr)  # weights sum to 1: un/re-normalisation cancels
    emu, R = _emulate_ds(sr.detach()[0, 0], tb)
    assert (emu - ref[0, 0]).abs().max() < 1e-6
    psi = torch.randn(n // 4, n // 4, generator=g, dtype=torch.float64)
    (ref * psi).sum().backward()
    # gather form used by loss.cu P4: dsr[r][c] = sum_ij tab[r][i] tab[c][j] psi[r/4-1+i][c/4-1+j]
    tab = tb["tab_ds"].astype(np.float64)
    out = np.zeros((n, n))
    P = psi.numpy()
    for r in range(n):
        for c in range(n):
            acc = 0.0
            for i in range(3):
                I = r // 4 - 1 + i
                if not 0 <= I < n // 4:
                    continue
                for j in range(3):
                    J = c // 4 - 1 + j
                    if 0 <= J < n // 4:
                        acc += tab[r, i] * tab[c, j] * P[I, J]
            out[r, c] = acc
    assert np.abs(out - sr.grad[0, 0].numpy()).max() < 1e-6
    # weights that fall outside the low-res image must be zero in the table
    assert tab[0, 0] == 0 and tab[n - 1, 2] == 0


def test_lowpass_adjoint_table():
    n = 64
    tb = _tables_np(n)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 1, n, n, generator=g, dtype=torch.float64, requires_grad=True)
    lp = O.lowpass(x, 0.25)
    g9 = tb["g9"].astype(np.float64)
    # forward emulation: separable reflect blur
    G = np.zeros((n, n))
    for r in range(n):
        for m in range(9):
            G[r, _reflect(r + m - 4, n)] += g9[m]
    G = torch.from_numpy(G)
    assert (G @ x.detach()[0, 0] @ G.T - lp[0, 0]).abs().max() < 1e-6
    psi = torch.randn(n, n, generator=g, dtype=torch.float64)
    (lp[0, 0] * psi).sum().backward()
    A = tb["tab_lp"].astype(np.float64)
    P = np.zeros((n + 8, n + 8))
    P[4:-4, 4:-4] = psi.numpy()
    tmp2 = np.zeros((n + 8, n))
    for rr in range(n + 8):
        for c in range(n):
            tmp2[rr, c] = sum(A[c, j] * P[rr, c + j] for j in range(9))
    out = np.zeros((n, n))
    for r in range(n):
        for c in range(n):
            out[r, c] = sum(A[r, i] * tmp2[r + i, c] for i in range(9))
    assert np.abs(out - x.grad[0, 0].numpy()).max() < 1e-6


def test_sobel_adjoint_gather_form():
    n = 16
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 1, n, n, generator=g, dtype=torch.float64, requires_grad=True)
    f = O.SOBEL4.double()[:, None]
    e = F.conv2d(x, f, padding="same")
    psi = torch.randn(1, 4, n, n, generator=g, dtype=torch.float64)
    (e * psi).sum().backward()
    P = F.pad(psi, (1, 1, 1, 1))[0].numpy()
    fs = O.SOBEL4.double().numpy()
    out = np.zeros((n, n))
    for r in range(n):
        for c in range(n):
            out[r, c] = sum(fs[k, ky, kx] * P[k, r - ky + 1 + 1, c - kx + 1 + 1] for k in range(4) for ky in range(3) for kx in range(3))
    assert np.abs(out - x.grad[0, 0].numpy()).max() < 1e-9


def test_dgrad_replicate_border_formula():
    """zero-padded transposed conv + the border pass of conv3x3.cu == autograd of replicate conv."""
    B, K, Oc, H, W = 1, 2, 3, 6, 5
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, K, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Oc, K, 3, 3, generator=g, dtype=torch.float64)
    dy = torch.randn(B, Oc, H, W, generator=g, dtype=torch.float64)
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w)
    (y * dy).sum().backward()
    dx = F.conv_transpose2d(dy, w, padding=1).numpy().copy()  # main kernel: zero padding, flipped/transposed weights
    wn, dyn = w.numpy(), dy.numpy()
    for k in range(K):
        for p in range(H):
            for q in range(W):
                if not (p in (0, H - 1) or q in (0, W - 1)):
                    continue
                ky_e = 0 if p == 0 else (2 if p == H - 1 else -1)
                r_e = 0 if p == 0 else H - 1
                kx_e = 0 if q == 0 else (2 if q == W - 1 else -1)
                c_e = 0 if q == 0 else W - 1
                s = 0.0
                for o in range(Oc):
                    if ky_e >= 0:
                        for kx in range(3):
                            c = q - kx + 1
                            if 0 <= c < W:
                                s += wn[o, k, ky_e, kx] * dyn[0, o, r_e, c]
                        if kx_e >= 0:
                            s += wn[o, k, ky_e, kx_e] * dyn[0, o, r_e, c_e]
                    if kx_e >= 0:
                        for ky in range(3):
                            r = p - ky + 1
                            if 0 <= r < H:
                                s += wn[o, k, ky, kx_e] * dyn[0, o, r, c_e]
                dx[0, k, p, q] += s
    assert np.abs(dx - x.grad.numpy()).max() < 1e-9


def test_bilinear_up2_adjoint_gather_form():
    n = 8
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 1, n, n, generator=g, dtype=torch.float32, requires_grad=True)
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    gy = torch.randn(1, 1, 2 * n, 2 * n, generator=g)
    (up * gy).sum().backward()
    rs = np.float32(n - 1) / np.float32(2 * n - 1)

    def coord(d):
        s = np.float32(rs * np.float32(d))
        i0 = int(s)
        i1 = i0 + (1 if i0 < n - 1 else 0)
        w1 = np.float32(s - np.float32(i0))
        return i0, i1, np.float32(1) - w1, w1

    # forward check of the coordinate rule
    U = np.zeros((2 * n, n), dtype=np.float64)
    for d in range(2 * n):
        i0, i1, w0, w1 = coord(d)
        U[d, i0] += w0
        U[d, i1] += w1
    assert np.abs(U @ x.detach()[0, 0].double().numpy() @ U.T - up[0, 0].detach().double().numpy()).max() < 1e-5
    G = gy[0, 0].double().numpy()
    out = np.zeros((n, n))
    for i in range(n):
        for k in range(n):
            acc = 0.0
            for t in range(6):
                y = 2 * i - 2 + t
                if not 0 <= y < 2 * n:
                    continue
                i0, i1, w0, w1 = coord(y)
                wy = (w0 if i0 == i else 0) + (w1 if i1 == i else 0)
                for u in range(6):
                    xx = 2 * k - 2 + u
                    if not 0 <= xx < 2 * n:
                        continue
                    j0, j1, v0, v1 = coord(xx)
                    wx = (v0 if j0 == k else 0) + (v1 if j1 == k else 0)
                    acc += wy * wx * G[y, xx]
            out[i, k] = acc
    assert np.abs(out - x.grad[0, 0].double().numpy()).max() < 1e-5


def test_loss_linearity_identities():
    """The two algebraic shortcuts of loss.cu: F4(sr) - gamma F4(ndvi) == F4(sr - gamma ndvi) and
    (downscale(sr*std+mean) - mean)/std == downscale(sr)."""
    g = torch.Generator().manual_seed(8)
    sr = torch.randn(1, 1, 64, 64, generator=g, dtype=torch.float64)
    nd = torch.randn(1, 1, 64, 64, generator=g, dtype=torch.float64)
    a = (O.downscale_to_lr(sr * O.STD_LST + O.MEAN_LST) - O.MEAN_LST) / O.STD_LST
    # the fp32-rounded PSF sums to 1 + 3e-9, i.e. a constant 1.5e-7 offset after the mean/std round trip:
    # two orders below fp32 resolution of the values themselves
    assert (a - O.downscale_to_lr(sr)).abs().max() < 1e-6
    hp = lambda t: t - O.lowpass(t, 0.25)
    assert (hp(sr) - (-0.25) * hp(nd) - hp(sr + 0.25 * nd)).abs().max() < 1e-12
