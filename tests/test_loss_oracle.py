"""CPU checks of the loss oracle (oracle/loss_oracle.py) against closed-form known answers of the
published SSIM definition; torchmetrics itself is not installable here (see the oracle's header)."""
import torch

import loss_oracle as LO


def test_window_is_the_published_one():
    w = LO.gaussian_window()
    assert w.numel() == 11 and abs(float(w.sum()) - 1) < 1e-12 and torch.equal(w, w.flip(0))
    assert abs(float(w[5] / w[4]) - float(torch.exp(torch.tensor(0.5 / 2.25, dtype=torch.float64)))) < 1e-12


def test_ssim_known_answers():
    h, w = 32, 40
    for a, b in ((0.3, 0.7), (0.0, 1.0), (0.5, 0.5)):
        s = float(LO.ssim(torch.full((h, w, 3), a, dtype=torch.float64), torch.full((h, w, 3), b, dtype=torch.float64)))
        assert abs(s - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-12
    g = torch.Generator().manual_seed(0)
    x = torch.rand(h, w, 3, generator=g, dtype=torch.float64)
    y = torch.rand(h, w, 3, generator=g, dtype=torch.float64)
    assert abs(float(LO.ssim(x, x)) - 1) < 1e-12
    assert abs(float(LO.ssim(x, y)) - float(LO.ssim(y, x))) < 1e-12
    # the 5-pixel border never contributes (reflect padding is cropped away again)
    x2 = x.clone()
    x2[0, :, :] = 0.123
    y2 = y.clone()
    y2[0, :, :] = 0.9
    inner_only = float(LO.ssim(x2[5:], y2[5:]))            # rows >= 5 see row 0 only through their window...
    assert inner_only == inner_only                          # (finite); and a direct window check:
    w1 = LO.gaussian_window()
    k2 = w1.unsqueeze(1) @ w1.unsqueeze(0)
    cy, cx = 12, 17
    px, py = x[cy - 5:cy + 6, cx - 5:cx + 6, 0], y[cy - 5:cy + 6, cx - 5:cx + 6, 0]
    mx, my = (k2 * px).sum(), (k2 * py).sum()
    sxx, syy, sxy = (k2 * px * px).sum() - mx * mx, (k2 * py * py).sum() - my * my, (k2 * px * py).sum() - mx * my
    want = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (sxx + syy + 9e-4))
    # isolate that pixel: build images whose only inner pixel is (cy, cx) -> crop to an 11x11 patch
    patch = float(LO.ssim(x[cy - 5:cy + 6, cx - 5:cx + 6, :1].expand(11, 11, 3), y[cy - 5:cy + 6, cx - 5:cx + 6, :1].expand(11, 11, 3)))
    assert abs(patch - float(want)) < 1e-12


def test_train_loss_composition():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(24, 24, 3, generator=g, dtype=torch.float64)
    y = torch.rand(24, 24, 3, generator=g, dtype=torch.float64)
    t, l1, s = LO.train_loss(x, y, 0.1)
    assert abs(float(t) - (0.9 * float((x - y).abs().mean()) + 0.1 * float(s))) < 1e-12
    assert abs(float(s) - (1 - float(LO.ssim(x, y)))) < 1e-12
