"""§8 f-3: the fused L1 + SSIM loss (csrc/loss.cu) against the fp64 CPU oracle (oracle/loss_oracle.py,
a restatement of torchmetrics' published SSIM + reference train.py:99-107), values and gradients."""
import pytest
import torch

import loss_oracle as LO
import synthetic as S
from helpers import rel_err, scene

pytestmark = pytest.mark.gpu


def _pair(h, w, seed, noise=0.15):
    g = torch.Generator("cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(6 * xx + 2 * yy), yy * xx, 0.5 + 0.5 * torch.cos(9 * yy)], dim=-1)
    gt = (base + 0.05 * torch.rand(h, w, 3, generator=g)).clamp(0, 1)
    img = (gt + noise * (torch.rand(h, w, 3, generator=g) - 0.5)).clamp(0, 1)
    return img.float().contiguous(), gt.float().contiguous()


@pytest.mark.parametrize("h,w,weight", [(48, 64, 0.1), (67, 45, 0.25), (11, 11, 0.1), (270, 480, 0.1)])
def test_l1_ssim_loss_vs_oracle(gs, cuda, h, w, weight):
    import loss
    img, gt = _pair(h, w, 3)
    x = img.to(cuda).requires_grad_(True)
    total, l1, ssim_loss = loss.l1_ssim_loss(x, gt.to(cuda), weight)
    total.backward()
    xo = img.double().requires_grad_(True)
    ot, ol1, os_ = LO.train_loss(xo, gt.double(), weight)
    ot.backward()
    assert abs(float(total) - float(ot)) < 2e-6 * max(1.0, abs(float(ot)))
    assert abs(float(l1) - float(ol1)) < 2e-6 and abs(float(ssim_loss) - float(os_)) < 5e-6
    assert rel_err(x.grad, xo.grad) < 2e-4
    # upstream scale flows through (loss * 3).backward()
    x2 = img.to(cuda).requires_grad_(True)
    (loss.l1_ssim_loss(x2, gt.to(cuda), weight)[0] * 3.0).backward()
    assert rel_err(x2.grad, 3.0 * xo.grad) < 2e-4


def test_ssim_known_answers_and_half_target(gs, cuda):
    import loss
    h, w = 40, 56
    # constant images: every window has zero variance -> ssim = (2ab + c1) / (a^2 + b^2 + c1)
    a, b = 0.3, 0.7
    s = float(loss.ssim(torch.full((h, w, 3), a, device=cuda), torch.full((h, w, 3), b, device=cuda)))
    # (fp32: E[x^2] - mu^2 cancels to ~1e-8 instead of 0 and c2 = 9e-4 amplifies that to ~5e-5, exactly as an
    #  fp32 evaluation of the torchmetrics formula does)
    assert abs(s - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 2e-4
    img, gt = _pair(h, w, 7)
    assert abs(float(loss.ssim(gt.to(cuda), gt.to(cuda))) - 1.0) < 1e-6          # identical images
    s_xy = float(loss.ssim(img.to(cuda), gt.to(cuda)))
    assert abs(s_xy - float(loss.ssim(gt.to(cuda), img.to(cuda)))) < 1e-6        # symmetric
    assert abs(s_xy - float(LO.ssim(img.double(), gt.double()))) < 5e-6
    # float16 ground truth (reference splatter.py:478 keeps `ground_truth` in half)
    gth = gt.half()
    x = img.to(cuda).requires_grad_(True)
    t, _, _ = loss.l1_ssim_loss(x, gth.to(cuda), 0.1)
    t.backward()
    xo = img.double().requires_grad_(True)
    to, _, _ = LO.train_loss(xo, gth.double(), 0.1)
    to.backward()
    assert abs(float(t) - float(to)) < 2e-6 and rel_err(x.grad, xo.grad) < 2e-4
    # psnr with torchmetrics' data_range=None semantics
    p = float(loss.psnr(img.to(cuda), gt.to(cuda)))
    mse = float(((img - gt) ** 2).mean())
    dr = float(gt.max() - gt.min())
    assert abs(p - 10 * torch.log10(torch.tensor(dr * dr / mse)).item()) < 1e-3
    with pytest.raises(RuntimeError, match="11x11"):
        loss.ssim(torch.zeros(8, 30, 3, device=cuda), torch.zeros(8, 30, 3, device=cuda))


def test_loss_drives_the_fused_backward(gs, cuda):
    """render -> fused loss -> backward: parameter gradients equal those obtained with the loss written
    in torch ops (the oracle's formula evaluated on the GPU through autograd)."""
    import loss
    import splatter
    n, w, h = 4000, 160, 96
    g, v, cam = scene(n, w, h, k=1)
    teacher, _, _ = scene(n, w, h, seed=5, k=1)
    vd = [dict(width=v.width, height=v.height, focal_x=v.fx, focal_y=v.fy, rot=v.rot, tran=v.tran)]
    with torch.no_grad():
        gt = splatter.Splatter.from_tensors(teacher, vd, device=cuda)(0).half()
    res = []
    for fused in (True, False):
        sp = splatter.Splatter.from_tensors(g, vd, device=cuda)
        img = sp(0)
        if fused:
            total, _, _ = loss.l1_ssim_loss(img, gt, 0.1)
        else:
            total, _, _ = LO.train_loss(img, gt.float(), 0.1)
        total.backward()
        res.append((float(total), [p.grad.clone() for p in sp.gaussian_3ds.parameters()]))
    assert abs(res[0][0] - res[1][0]) < 1e-5
    for a, b in zip(res[0][1], res[1][1]):
        assert rel_err(a, b) < 1e-3
