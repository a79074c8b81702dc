import torch


def _as_hwc(x):
    """[1, 3, H, W] (usually a permuted view of an [H, W, 3] image, train.py:102-103) -> [H, W, 3]."""
    if x.dim() == 4:
        if x.shape[0] != 1:
            raise NotImplementedError("shim: batch size 1 only (reference train.py:102)")
        x = x[0]
    if x.dim() != 3 or x.shape[0] != 3:
        raise NotImplementedError("shim: expects [1, 3, H, W] or [3, H, W]")
    return x.permute(1, 2, 0)


def structural_similarity_index_measure(preds, target, data_range=1.0, **_kw):
    import loss                                       # 3d-gaussian-splatting_b200/loss.py (CUDA, no fallback)
    return loss.ssim(_as_hwc(preds), _as_hwc(target))


def peak_signal_noise_ratio(preds, target, data_range=None, **_kw):
    t = target.to(preds.dtype)
    mse = torch.mean((preds - t) ** 2)
    dr = (t.max() - t.min()) if data_range is None else torch.as_tensor(float(data_range), device=preds.device)
    return 10.0 * torch.log10(dr * dr / mse)
