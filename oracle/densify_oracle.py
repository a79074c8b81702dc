"""CPU/torch oracle of the densification step.  TEST INFRASTRUCTURE ONLY.

Restates reference splatter.py:122-228 (`Gaussian3ds.adaptive_control`) with plain torch ops, same
output layout as its torch.cat calls: [kept (split ones moved + shrunk)], [clones], [second split
samples].  The reference draws the two split positions with
`MultivariateNormal(pos, cov).sample()` (utils.py:391-402, cov from the UN-shrunk scale and the raw
quaternion, splatter.py:100-114,204-205); here the same distribution is sampled as
pos + R diag(s) z from explicit standard normals `z[2, n_split, 3]`, so that the CUDA kernel can be
compared element for element.
"""
import math

import torch

EPS = 1e-4


def inverse_sigmoid(y):
    return -math.log(1 / y - 1)


def quat_to_rotmat(q):                     # utils.py:318-333, no normalisation
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
                        2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w,
                        2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y],
                       dim=-1).reshape(q.shape[:-1] + (3, 3))


def adaptive_control(pos, rgb, opa, quat, scale, grad, taus, delete_thresh, scale_activation="abs", grad_thresh=0.0002,
                     grad_aggregation="max", use_clone=True, use_split=True, clone_dt=0.01, z=None):
    norm = scale.norm(dim=-1) if scale_activation == "abs" else scale.exp().norm(dim=-1)            # :129-136
    keep = (opa > inverse_sigmoid(0.02)) & (norm < delete_thresh)                                   # :137-139
    pos, rgb, opa, quat, scale, grad, norm = (t[keep].clone() for t in (pos, rgb, opa, quat, scale, grad, norm))
    agg = grad.abs().max(-1)[0] if grad_aggregation == "max" else grad.abs().mean(-1)               # :152-157
    densify = agg > grad_thresh
    split_mask = (norm > taus) & densify & bool(use_split)
    clone_mask = (norm <= taus) & densify & bool(use_clone)
    new = [[pos], [rgb], [opa], [quat], [scale]]
    if bool(clone_mask.any()):
        for lst, t in zip(new, (pos[clone_mask] - grad[clone_mask] * clone_dt, rgb[clone_mask], opa[clone_mask],
                                quat[clone_mask], scale[clone_mask])):
            lst.append(t.clone())
    n_split = int(split_mask.sum())
    if n_split:
        R = quat_to_rotmat(quat[split_mask])
        s = scale[split_mask].abs() + EPS if scale_activation == "abs" else torch.exp(scale[split_mask])
        L = R * s.unsqueeze(-2)                                                                       # R diag(s)
        p1 = pos[split_mask] + (L @ z[0].unsqueeze(-1)).squeeze(-1)
        p2 = pos[split_mask] + (L @ z[1].unsqueeze(-1)).squeeze(-1)
        if scale_activation == "abs":
            scale[split_mask] = scale[split_mask] / 1.6
        else:
            scale[split_mask] = scale[split_mask] - math.log(1.6)
        pos[split_mask] = p1
        for lst, t in zip(new, (p2, rgb[split_mask], opa[split_mask], quat[split_mask], scale[split_mask])):
            lst.append(t.clone())
    out = [torch.cat(l) for l in new]
    return out, dict(deleted=int((~keep).sum()), cloned=int(clone_mask.sum()), split=n_split)
