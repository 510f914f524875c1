"""CPU model of the one-pass LayerNorm backward (csrc/otb_norm.cu ln_bwd_fused_kernel + ln_bwd_finalize_wide_kernel):
the thread -> column-vector ownership (v = tid + j*256), row batches of kLnR, per-CTA partial rows in ws[2][grid][D] and
the finalize, restated in numpy and checked against torch autograd."""
import numpy as np
import torch

K_LN_R = 4


def fused_model(dy, x, mean, rstd, gamma, add, grid):
    rows, D = x.shape
    nvec = D // 8
    vpt = 1 if D <= 2048 else 2
    dx = np.zeros_like(x)
    ws = np.zeros((2, grid, D), np.float64)
    nbatch = (rows + K_LN_R - 1) // K_LN_R
    for cta in range(grid):
        pg = np.zeros(D)
        pb = np.zeros(D)
        for b in range(cta, nbatch, grid):
            for r in range(K_LN_R):
                row = b * K_LN_R + r
                if row >= rows:
                    continue
                owned = [v for tid in range(256) for j in range(vpt) for v in [tid + j * 256] if v < nvec]
                assert sorted(owned) == list(range(nvec))            # every column vector has exactly one owner
                xh = (x[row] - mean[row]) * rstd[row]
                g = dy[row] * gamma
                s1, s2 = g.sum() / D, (g * xh).sum() / D
                pg += dy[row] * xh
                pb += dy[row]
                dx[row] = rstd[row] * (g - s1 - xh * s2) + add[row]
        ws[0, cta], ws[1, cta] = pg, pb
    return dx, ws[0].sum(0), ws[1].sum(0)


def test_fused_ln_backward_model():
    rng = np.random.RandomState(1)
    for rows, D, grid in ((37, 256, 10), (101, 3072, 26), (6, 4096, 2)):
        x = rng.randn(rows, D) * 2
        dy, add = rng.randn(rows, D), rng.randn(rows, D)
        gamma = 1 + 0.1 * rng.randn(D)
        mean = x.mean(1)
        rstd = 1.0 / np.sqrt(x.var(1) + 1e-5)
        dx, dg, db = fused_model(dy, x, mean, rstd, gamma, add, grid)
        xt = torch.tensor(x, requires_grad=True)
        gt = torch.tensor(gamma, requires_grad=True)
        bt = torch.zeros(D, dtype=torch.float64, requires_grad=True)
        torch.nn.functional.layer_norm(xt, (D,), gt, bt, 1e-5).backward(torch.tensor(dy))
        assert np.allclose(dx, xt.grad.numpy() + add, atol=1e-9)
        assert np.allclose(dg, gt.grad.numpy(), atol=1e-9) and np.allclose(db, bt.grad.numpy(), atol=1e-9)
