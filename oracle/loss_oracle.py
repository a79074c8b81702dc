"""CPU oracle of the training loss.  TEST INFRASTRUCTURE ONLY.

reference train.py:99-107 computes  (1-w) * mean|img-gt| + w * (1 - SSIM)  with
`torchmetrics.StructuralSimilarityIndexMeasure(data_range=1.0)` (train.py:72).  torchmetrics is a
third-party dependency that is absent from /root/reference and from this image (requirements.txt
lists it unpinned), so its published algorithm is restated here from
torchmetrics/functional/image/ssim.py `_ssim_update` (v1.x; defaults gaussian_kernel=True,
kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
  * 1-D window exp(-(d/sigma)^2/2), d = -5..5, normalised; 2-D = outer product; per-channel (grouped) conv;
  * inputs reflect-padded by 5, `conv2d` of the five stacked maps (x, y, x^2, y^2, xy);
  * ssim map = ((2 mu_x mu_y + c1)(2 s_xy + c2)) / ((mu_x^2 + mu_y^2 + c1)(s_xx + s_yy + c2));
  * the same 5-pixel border is cropped from the map before the mean (so the padding never contributes);
  * mean over pixels and channels (per image, then over the batch of 1).
PARITY UNPINNED against torchmetrics itself (it cannot be installed here); the formula above is
the published one and the tests pin the CUDA kernels to this restatement in fp64 (+ finite
differences of it through autograd).
"""
import torch
import torch.nn.functional as F


def gaussian_window(kernel_size=11, sigma=1.5, dtype=torch.float64, device=None):
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=dtype, device=device)
    g = torch.exp(-((dist / sigma) ** 2) / 2)
    return g / g.sum()


def ssim(pred_hwc, target_hwc, data_range=1.0, k1=0.01, k2=0.03):
    """Mean SSIM of two [H, W, 3] images, torchmetrics semantics (NCHW inside)."""
    dt = pred_hwc.dtype
    x = pred_hwc.permute(2, 0, 1).unsqueeze(0)
    y = target_hwc.to(dt).permute(2, 0, 1).unsqueeze(0)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    w1 = gaussian_window(dtype=dt, device=x.device)
    ch = x.shape[1]
    kernel = (w1.unsqueeze(1) @ w1.unsqueeze(0)).expand(ch, 1, 11, 11).contiguous()
    pad = 5
    xp = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    yp = F.pad(y, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat((xp, yp, xp * xp, yp * yp, xp * yp))
    out = F.conv2d(stack, kernel, groups=ch)
    mu_x, mu_y, exx, eyy, exy = out.split(1)
    sxx, syy, sxy = exx - mu_x ** 2, eyy - mu_y ** 2, exy - mu_x * mu_y
    full = ((2 * mu_x * mu_y + c1) * (2 * sxy + c2)) / ((mu_x ** 2 + mu_y ** 2 + c1) * (sxx + syy + c2))
    inner = full[..., pad:-pad, pad:-pad]
    return inner.reshape(inner.shape[0], -1).mean(-1).mean()


def train_loss(img, gt, ssim_weight=0.1):
    """train.py:99-107 -> (loss, l1_loss, ssim_loss)."""
    l1 = (img - gt.to(img.dtype)).abs().mean()
    s = 1.0 - ssim(img, gt)
    return (1 - ssim_weight) * l1 + ssim_weight * s, l1, s
