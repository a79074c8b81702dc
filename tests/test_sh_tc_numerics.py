"""CPU check of the operand precision the tensor-core SH kernels rely on (3d-gaussian-splatting_b200/csrc/blend_sh_tc.cu,
tc_common.cuh::split_bf16x2): every factor is split into bf16 hi + lo parts and a product is evaluated as
hi*hi + hi*lo + lo*hi with fp32 accumulation.  The emulation below uses torch's bf16 rounding (round-to-nearest-even, the
same conversion `__floats2bfloat162_rn` performs) and bounds the error of the colour logits and of the coefficient
gradients against fp64 on data of the benchmark's scale.  No GPU, no extension."""
import math

import torch


def split(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def contract3(a, b):
    """sum_k a[..., k] * b[..., k] from the three products the kernels issue, accumulated in fp32."""
    ah, al = split(a)
    bh, bl = split(b)
    return (ah * bh).sum(-1) + (ah * bl).sum(-1) + (al * bh).sum(-1)


def sh_basis(d, k):
    x, y, z = d.unbind(-1)
    out = [torch.full_like(x, 0.28209479177387814), -0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    out += [1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.31539156525252005 * (2 * zz - xx - yy),
            -1.0925484305920792 * xz, 0.5462742152960396 * (xx - yy)]
    if k == 16:
        out += [-0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * xy * z, -0.4570457994644658 * y * (4 * zz - xx - yy),
                0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy), -0.4570457994644658 * x * (4 * zz - xx - yy),
                1.445305721320277 * z * (xx - yy), -0.5900435899266435 * x * (xx - 3 * yy)]
    return torch.stack(out, -1)


def test_split_reconstructs_to_2_pow_minus_16():
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(100000, generator=g) * 2 - 1) * torch.logspace(-6, 3, 100000)
    hi, lo = split(x)
    rel = ((hi + lo).double() - x.double()).abs() / x.double().abs()
    assert float(rel.max()) <= 2.0 ** -16


def test_logits_and_colours_within_the_image_budget():
    g = torch.Generator().manual_seed(1)
    for k in (9, 16):
        d = torch.nn.functional.normalize(torch.randn(4096, 3, generator=g), dim=-1)
        sh = sh_basis(d, k)                                             # [P, K]
        coef = torch.randn(512, k, generator=g) * 0.3                   # synthetic.make_gaussians: N(0, 0.3) + DC term
        coef[:, 0] += torch.randn(512, generator=g) * 3.0
        exact = sh.double() @ coef.double().T                           # logits [P, I]
        approx = contract3(sh[:, None, :], coef[None, :, :])
        err = (approx.double() - exact).abs().max()
        bound = 2.0 ** -15 * float((sh.abs()[:, None, :] * coef.abs()[None, :, :]).sum(-1).max())
        assert float(err) <= bound
        # a pixel's colour is a convex combination of sigmoids (weights sum <= 1): sigmoid' <= 1/4
        assert 0.25 * float(err) < 2e-5                                 # budget: 1e-4 absolute per pixel


def test_coefficient_gradient_contraction_within_the_gradient_budget():
    g = torch.Generator().manual_seed(2)
    d = torch.nn.functional.normalize(torch.randn(256, 3, generator=g), dim=-1)
    sh = sh_basis(d, 16)                                                # one tile: 256 pixels
    # logit gradients of one instance: few pixels carry most of the weight, mixed signs, wide dynamic range
    dl = torch.randn(64, 256, generator=g) * torch.logspace(-8, 0, 256)[torch.randperm(256, generator=g)]
    exact = dl.double() @ sh.double()                                   # [I, K]
    approx = contract3(dl[:, None, :], sh.T[None, :, :])
    rel = (approx.double() - exact).abs().max() / exact.abs().max()
    assert float(rel) < 1e-4                                            # budget: 1e-3 of the largest gradient
    assert math.isfinite(float(rel))
